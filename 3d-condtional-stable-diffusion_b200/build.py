"""Build libb200dm.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python 3d-condtional-stable-diffusion_b200/build.py [--force]

The .so is git-ignored but travels with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# (library, object directory, extra nvcc flags): the same sources with bf16 and with IEEE fp16 as the 16-bit storage type
VARIANTS = {"bf16": (os.path.join(HERE, "libb200dm.so"), os.path.join(HERE, "build"), []),
            "fp16": (os.path.join(HERE, "libb200dm_f16.so"), os.path.join(HERE, "build_f16"), ["-DB200DM_ACT_FP16"])}
OUT = VARIANTS["bf16"][0]
SOURCES = ["api.cu", "update.cu", "norm.cu", "norm_ex.cu", "vq.cu", "conv.cu", "attention.cu", "program.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-DB200DM_BUILD", "--expt-relaxed-constexpr"]


def _stamp():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode() + fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Build both storage variants; returns the path of the default (bf16) library."""
    with ThreadPoolExecutor(max_workers=len(VARIANTS)) as ex:   # the two variants compile side by side
        list(ex.map(lambda name: _build_variant(name, force, verbose), VARIANTS))
    return OUT


def _build_variant(name: str, force: bool, verbose: bool) -> str:
    OUT, OBJ, extra = VARIANTS[name]
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _stamp() + name
    if not force and os.path.exists(OUT) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return OUT
    if not os.path.exists(NVCC):
        if os.path.exists(OUT):  # GPU box without a toolchain: use the shipped library
            return OUT
        raise RuntimeError(f"nvcc not found at {NVCC} and no prebuilt {OUT}")

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

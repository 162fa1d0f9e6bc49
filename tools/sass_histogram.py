"""Instruction histogram of the shipped library per kernel: the SASS mnemonics that prove tcgen05 / TMEM / TMA are on the path
(B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store, HMMA = mma.sync).

    python tools/sass_histogram.py [lib.so] > profiles/sass_histogram.txt        (runs without a GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "3d-condtional-stable-diffusion_b200", "libb200dm.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "LDSM", "FFMA2", "FFMA", "MUFU",
        "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "ACQBULK", "ELECT"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
kern, hist, total = None, {}, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern], total[kern] = collections.Counter(), 0
        continue
    m = re.search(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                hist[kern][k] += 1
                break
print(f"# {os.path.relpath(lib, ROOT)}: SASS instruction counts per kernel (cuobjdump -sass), {len(hist)} kernels")
print("# columns: total | " + " ".join(KEYS))
for k in sorted(hist, key=lambda k: -hist[k]["UTCHMMA"] - hist[k]["UTCHMMA.2CTA"] - hist[k]["HMMA"]):
    name = re.sub(r"\(.*", "", demangle(k).replace("(anonymous namespace)::", ""))
    name = re.sub(r"^void ", "", name)
    print(f"{name[:110]:110s} {total[k]:6d} | " + " ".join(f"{hist[k][c]:4d}" for c in KEYS))
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print("# library totals: " + ", ".join(f"{c} {tot[c]}" for c in KEYS if tot[c]))

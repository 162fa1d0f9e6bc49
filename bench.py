#!/usr/bin/env python
"""bench.py -- latent-volume denoise steps/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host CPU (oracle port)

Workload (BASELINE.json configs[1], SURVEY cfg-2): conditional_dm3d U-Net (F=32, widths 64/128/256, cross-attention at
8^3), DDPM, batch 8 per GPU, 32^3 x 256 latent, class-id context, random-init weights, synthetic latents.
A "step" = one reverse-diffusion step of the whole batch (U-Net forward + fused posterior update);
value = volumes * steps / second over all GPUs (weak scaling: independent sample batches per GPU, no collective).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "latent-volume denoise steps/sec"
UNIT = "volume-steps/s"
CFG = dict(S=32, C_lat=256, B=8, T=1000)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return dict(hbm=j["hbm_gbs"], tf=j["bf16_tflops"], tf_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return dict(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)),
                    samples=len(sm), reasons=sorted(reasons))


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world)
    return world, rank, local


def max_over_ranks(v, world, device):
    if world == 1:
        return v
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


# ----------------------------------------------------------------------------------------------- CPU (reference) arm
def oracle_cpu_rate(steps, warmup, budget_s=90.0):
    """The reference's algorithm for this path on the host CPU (oracle port: PyTorch-CPU fp32 restatement of
    conditional_dm3d.build_model + DiffusionModel.sample; TensorFlow is not installable here).  Bounded sample:
    ONE volume (B=1) of the cfg-2 workload per step, all host threads."""
    from oracle.unet import UNet as OUNet
    from oracle import init as OI, sampler as OS
    from oracle.schedule import Betas
    torch.set_num_threads(os.cpu_count() or 1)
    S, C = CFG["S"], CFG["C_lat"]
    ou = OUNet(S, C, [64, 128, 256], [False, False, True, True], first_conv_channels=32, conditional=True)
    P = OI.make_params(ou.spec(), 0, "keras")
    b = Betas(CFG["T"])
    x = OI.normal((1, S, S, S, C), 1)
    ctx = torch.tensor([1])
    z = OI.normal((1, S, S, S, C), 2)

    def one(t):
        nonlocal x
        with torch.no_grad():
            eps = ou.forward(P, x, torch.tensor([t]), ctx=ctx)
            x = OS.ddpm_step(b, x, eps, t, z)

    t0 = time.perf_counter()
    one(CFG["T"] - 1)
    first = time.perf_counter() - t0
    w = max(0, min(warmup - 1, int(budget_s * 0.2 / max(first, 1e-3))))
    for i in range(w):
        one(CFG["T"] - 2 - i)
    n = max(1, min(steps, int(budget_s / max(first, 1e-3))))
    t0 = time.perf_counter()
    for i in range(n):
        one(CFG["T"] - 2 - w - i)
    dt = time.perf_counter() - t0
    return dict(rate=n / dt, steps=n, warmup=w + 1, ms_per_step=1e3 * dt / n, cores=torch.get_num_threads(),
                sample=f"1 volume (of the 8-volume batch) x {n} DDPM steps of cfg-2 (32^3x256 latent, conditional U-Net), PyTorch-CPU fp32 oracle port")


def run_reference(args):
    world, rank, local = dist_setup(args.gpus)
    if rank != 0:
        barrier(world)
        return
    r = oracle_cpu_rate(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["rate"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(1, per_step_volumes=1),
            "cpu_baseline": {"value": r["rate"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["rate"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = TensorFlow/Keras (not installable offline): timed its CPU restatement (oracle/) on the host cores; rank 0 only"}
    print(json.dumps(line), flush=True)
    barrier(world)


def workload_config(n_gpus, per_step_volumes=None):
    return {"workload": "cfg-2: conditional_dm3d U-Net, DDPM T=1000, 32^3x256 latent, class-id context, batch 8 per GPU",
            "latent": [CFG["S"]] * 3 + [CFG["C_lat"]], "batch_per_gpu": CFG["B"] if per_step_volumes is None else per_step_volumes,
            "global_batch": (CFG["B"] if per_step_volumes is None else per_step_volumes) * n_gpus, "timesteps": CFG["T"],
            "parallelism": f"independent sample batches x{n_gpus} (no collective on the sampling path)",
            "l2": "not flushed: one step touches >10 GB of activations+weights, far larger than the 126 MB L2"}


# ----------------------------------------------------------------------------------------------- CUDA arm
def run_ours(args):
    world, rank, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import b200dm
    from b200dm import _lib as L, ops

    S, C, B, T = CFG["S"], CFG["C_lat"], CFG["B"], CFG["T"]
    K, W = args.steps, max(args.warmup, 3)
    a = types.SimpleNamespace(timesteps=T, num_gpus=world, kernel_resize=False, bs=B * world)
    dm = b200dm.ConditionalDiffusionModel(S, 1024, C, None, a)
    shape = (B, S, S, S, C)
    sid0 = rank * B
    ctx = (torch.arange(B) + sid0) % 2
    # ---- public-API warm-up: compiles the program, captures the step graph, runs W+ steps
    lat = dm.generate(shape, last_step=T - W, seed=1234, sample_id0=sid0, context=ctx)
    assert torch.isfinite(lat).all(), "non-finite latents"
    assert L.debug_flag() == 0, "tcgen05/TMA watchdog fired"
    st = dm._step
    graph, nets = st["graph"], st["nets"]
    launches_per_step = sum(net.prog.num_launches + 1 for net in nets) + 1   # per chain: U-Net program + update; + t advance

    def reset(t0):
        st["t_dev"].copy_(torch.tensor([t0, t0 - 1], dtype=torch.int32))

    # ---- device-resident timing: K graph replays between CUDA events on the launching stream
    reset(T - 1)
    for _ in range(W):
        graph.replay()
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.3)
    barrier(world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    clk = clocks.stop()
    value = world * B * K / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers: pinned x_T -> H2D -> K steps -> D2H latents
    x_host = torch.randn(shape, generator=torch.Generator().manual_seed(rank)).pin_memory()
    out_host = torch.empty(shape, dtype=torch.float32).pin_memory()
    dm.generate(shape, last_step=T - 2, x_T=x_host, seed=1234, sample_id0=sid0, context=ctx)  # warm the path
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lat = dm.generate(shape, last_step=T - K, x_T=x_host, seed=1234, sample_id0=sid0, context=ctx)
    out_host.copy_(lat, non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0, world, dev)
    barrier(world)
    nbytes = x_host.numel() * 4
    e2e = {"value": world * B * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nbytes / K, "d2h_bytes_per_step": nbytes / K,
           "call": "ConditionalDiffusionModel.generate(shape, last_step=T-K, x_T=<pinned host>, context=ids) -> host latents"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(world), chains=st["chains"]), "clocks": clk, "e2e": e2e, "gpu_launches": launches_per_step * K}

    if rank == 0:
        # ---- roofline of the dominant kernel (conv_igemm): per-launch CUDA events over one more step
        pk = peaks()
        reset(T - 1)
        rows = []
        for net in nets:          # every chain's program, one launch at a time with an event pair around each
            net.prog.run_timed()
            rows += net.prog.run_timed()
        tot_ms = sum(r[3] for r in rows)

        def agg(kinds):
            sel = [r for r in rows if r[0] in kinds]
            ms_, work = sum(r[3] for r in sel), sum(r[2] for r in sel)
            return sel, ms_, work

        # dominant kernel of the step: the persistent halo-reuse 3^3 conv (conv_halo_kernel)
        sel, h_ms, h_fl = agg(("conv_halo",))
        ach = h_fl / (h_ms * 1e-3) / 1e12
        line["roofline"] = {"kernel": "conv_halo_kernel (persistent tcgen05 implicit-GEMM 3^3 Conv3D, TMA halo slabs): all its launches of one step",
                            "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                            "peak_source": f"{pk['src']} MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)",
                            "traffic": None, "launches": len(sel), "avg_launch_ms": h_ms / max(1, len(sel)),
                            "algorithmic_gflop_per_launch_avg": h_fl / 1e9 / max(1, len(sel)), "share_of_step": h_ms / tot_ms,
                            "timing": "CUDA events around every launch on the launching stream (b200dm_program_run_timed)"}
        try:   # ncu dram__bytes_(read+write) of representative launches (one --set full capture each; profiles/)
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1d_traffic.json")) as f:
                line["roofline"]["traffic_samples"] = json.load(f)["samples"]
            line["roofline"]["traffic_note"] = "traffic is null for the 36-launch aggregate; per-launch ncu samples are listed in traffic_samples"
        except OSError:
            pass
        sel, t_ms, t_fl = agg(("conv_halo", "conv", "attn"))
        line["roofline_all_tensor_kernels"] = {"kernels": "conv_halo_kernel + conv_igemm_kernel (1^3 / strided / parity convs, GEMMs) + flash_attn_kernel",
                                               "bound": "tensor", "achieved": t_fl / (t_ms * 1e-3) / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                                               "frac": t_fl / (t_ms * 1e-3) / 1e12 / pk["tf_sustained"], "launches": len(sel),
                                               "algorithmic_gflop_per_step": t_fl / 1e9, "share_of_step": t_ms / tot_ms}
        sel, e_ms, e_by = agg(("norm_act", "layernorm", "gn_stats", "softmax"))
        if sel:
            line["roofline_hbm_kernels"] = {"kernels": "norm_act / layernorm (fused elementwise passes left in the step)", "bound": "hbm",
                                            "achieved": e_by / (e_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                            "frac": e_by / (e_ms * 1e-3) / 1e9 / pk["hbm"], "launches": len(sel), "share_of_step": e_ms / tot_ms}
        # the two HBM-bound kernels that move the bytes: the fused posterior update (largest pass of the step) and the largest
        # BN+swish(+concat) pass -- each timed alone with CUDA events on its stream (tiny launches dominate the aggregate above)
        net0 = nets[0]
        xs = st["x"][:st["chain_batch"]]
        reset(T - 1)

        def upd():
            L.check(L.lib().b200dm_ddpm_update(ctypes.byref(st["descs"][0]), L.ptr(xs), L.ptr(net0.eps), None, L.ptr(xs), L.ptr(net0.x_in), L.stream()))

        for _ in range(3):
            upd()
        torch.cuda.synchronize()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        for _ in range(10):
            upd()
        u1.record()
        torch.cuda.synchronize()
        u_ms = u0.elapsed_time(u1) / 10
        u_bytes = xs.numel() * (4.0 + 4.0 + 4.0 + 2.0)
        line["roofline_update_kernel"] = {"kernel": "update_kernel (fused DDPM posterior + Philox noise): x_t fp32 + eps fp32 -> x_{t-1} fp32 + bf16 copy",
                                          "bound": "hbm", "achieved": u_bytes / (u_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                          "frac": u_bytes / (u_ms * 1e-3) / 1e9 / pk["hbm"], "algorithmic_bytes_per_launch": u_bytes,
                                          "avg_launch_ms": u_ms, "working_set": "940 MB per launch (> 126 MB L2)"}
        big = max((r for r in rows if r[0] == "norm_act"), key=lambda r: r[2], default=None)
        if big is not None:
            line["roofline_largest_norm_pass"] = {"kernel": f"norm_act_kernel ({big[1]})", "bound": "hbm", "achieved": big[2] / (big[3] * 1e-3) / 1e9,
                                                  "peak": pk["hbm"], "unit": "GB/s", "frac": big[2] / (big[3] * 1e-3) / 1e9 / pk["hbm"],
                                                  "algorithmic_bytes_per_launch": big[2], "avg_launch_ms": big[3]}
        if world == 1 and not args.no_cpu:
            r = oracle_cpu_rate(3, 1, budget_s=25.0)
            line["cpu_baseline"] = {"value": r["rate"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        if args.dump_ops:
            with open(args.dump_ops, "w") as f:
                for r in rows:
                    f.write(f"{r[0]},{r[1]},{r[2]:.6g},{r[3]:.6f}\n")
        print(json.dumps(line), flush=True)
    barrier(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--batch", type=int, default=None, help="tuning aid: per-GPU batch other than the cfg-2 value (8); not a bench line")
    ap.add_argument("--dump-ops", default=None, help="write per-op (kind,name,work,ms) CSV of one step")
    args = ap.parse_args()
    if args.batch:
        CFG["B"] = args.batch
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

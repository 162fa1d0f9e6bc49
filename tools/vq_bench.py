"""cfg-3 VQ (N = 16 x 32^3 rows, K = 1024, D = 256): fp32 SIMT kernel vs tensor-core candidate search, CUDA-event timed.
    python tools/vq_bench.py [N] [K] [D]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import b200dm
from b200dm import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16 * 32 ** 3
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
D = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
x = torch.randn(N, D, device=dev, generator=g) * 0.5
cb = (torch.rand(K, D, device=dev, generator=g) - 0.5) * 0.1
sq = ops.vq_prepare(cb)
tc = ops.vq_prepare_tc(cb, sq)
res = {}
for name, ws in (("simt", None), ("tc", tc)):
    stats = torch.zeros(3, dtype=torch.int64, device=dev)
    for _ in range(2):
        idx, q = ops.vq_argmin_gather(x, cb, sq, tc_ws=ws, stats=stats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        idx, q = ops.vq_argmin_gather(x, cb, sq, tc_ws=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[name] = idx
    print(f"{name}: {ms:.3f} ms  {2 * N * K * D / ms / 1e9:.1f} TFLOP/s algorithmic  stats(2 warm calls) {stats.tolist()}")
print("identical indices:", torch.equal(res["simt"], res["tc"]), " flag", b200dm._lib.debug_flag())

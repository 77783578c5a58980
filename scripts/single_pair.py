"""GPU box: s3d_gicp_align on ONE pair, a few calls (the persistent-loop path) — for ncu captures and latency numbers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, slam3d_b200, bench
ctx = slam3d_b200.Context([0])
s, t, _ = bench.make_pairs(1)[0]
hs, ht = torch.from_numpy(slam3d_b200.as_xyzw(s)).pin_memory(), torch.from_numpy(slam3d_b200.as_xyzw(t)).pin_memory()
p = bench.params()
for _ in range(3):
    r = ctx.gicp_align(hs, ht, None, p)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
t0 = time.perf_counter()
for _ in range(n):
    r = ctx.gicp_align(hs, ht, None, p)
print(f"single pair: {1e3 * (time.perf_counter() - t0) / n:.3f} ms per align, status {r.status}, {r.outer_iterations} outer iterations, "
      f"{ctx.counters()['kernel_launches'] / (n + 3):.0f} launches per align")

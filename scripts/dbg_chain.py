import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, slam3d_b200
from slam3d_b200 import synth
from slam3d_b200._abi import RegistrationParameters
ctx = slam3d_b200.Context()
p = RegistrationParameters.defaults(point_cloud_density=0.1)
scans, _ = synth.trajectory(seed=20260117, n_scans=9)
order = list(range(9)) + list(range(7, 0, -1))
B = 64
seq = [order[i % len(order)] for i in range(B + 1)]
pinned = [torch.from_numpy(slam3d_b200.as_xyzw(s)).pin_memory() for s in scans]
host_seq = [pinned[j] for j in seq]
for rep in range(4):
    t0 = time.perf_counter(); hh = ctx.prepare_clouds(host_seq, 0.1, 20); t1 = time.perf_counter()
    rr = ctx.gicp_align_prepared_batch(hh[:-1], hh[1:], None, p); t2 = time.perf_counter()
    for h in hh: h.release()
    t3 = time.perf_counter()
    print(f"prepare {1e3*(t1-t0):.2f} ms  align {1e3*(t2-t1):.2f} ms release {1e3*(t3-t2):.2f} ms  iters {np.mean([r.outer_iterations for r in rr]):.2f} inner {np.mean([r.inner_iterations for r in rr]):.2f} sizes {hh[0].size if False else ''}")
t0 = time.perf_counter(); rr2 = ctx.gicp_align_batch(host_seq[:-1], host_seq[1:], None, p); t1 = time.perf_counter()
rr2 = ctx.gicp_align_batch(host_seq[:-1], host_seq[1:], None, p); t2 = time.perf_counter()
print(f"raw batch on the same pairs {1e3*(t2-t1):.2f} ms iters {np.mean([r.outer_iterations for r in rr2]):.2f}")

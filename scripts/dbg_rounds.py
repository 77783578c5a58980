import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, slam3d_b200
from slam3d_b200 import synth
from slam3d_b200._abi import RegistrationParameters
os.environ["S3D_STREAMS_PER_DEVICE"] = "1"
ctx = slam3d_b200.Context()
p = RegistrationParameters.defaults(point_cloud_density=0.1)
pairs = [synth.scan_pair(seed=20260117 + i) for i in range(4)]
B = 22
srcs = [torch.from_numpy(slam3d_b200.as_xyzw(pairs[i % 4][0])).cuda() for i in range(B)]
tgts = [torch.from_numpy(slam3d_b200.as_xyzw(pairs[i % 4][1])).cuda() for i in range(B)]
for _ in range(3): rr = ctx.gicp_align_batch(srcs, tgts, None, p)
os.environ["S3D_TRACE"] = "1"
t0 = time.perf_counter(); rr = ctx.gicp_align_batch(srcs, tgts, None, p); print("total ms", 1e3 * (time.perf_counter() - t0), [ (r.outer_iterations, r.inner_iterations) for r in rr[:4]])

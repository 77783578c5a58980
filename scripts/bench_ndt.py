"""NDT branch timing on the bench workload (synthetic 131k-point pairs, 0.1 m voxel, slam3d NDT defaults): batch of pairs
through s3d_gicp_align_batch with host scans (e2e) and per-stage device times; oracle on one host thread beside it."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_b200
from slam3d_b200 import _abi, synth
from slam3d_b200._abi import RegistrationParameters

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=32)
ap.add_argument("--distinct", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--oracle", action="store_true")
a = ap.parse_args()
P = RegistrationParameters.defaults(point_cloud_density=0.1, registration_algorithm=_abi.ALG_NDT)
scenes = [synth.scan_pair(seed=100 + i) for i in range(a.distinct)]
srcs = [slam3d_b200.as_xyzw(scenes[i % a.distinct][0]) for i in range(a.pairs)]
tgts = [slam3d_b200.as_xyzw(scenes[i % a.distinct][1]) for i in range(a.pairs)]
ctx = slam3d_b200.Context()
res = ctx.gicp_align_batch(srcs, tgts, None, P)  # warm-up
ctx.set_profiling(True)
ctx.stage_times(reset=True)
t0 = time.perf_counter()
for _ in range(a.steps):
    res = ctx.gicp_align_batch(srcs, tgts, None, P)
dt = (time.perf_counter() - t0) / a.steps
out = {"pairs": a.pairs, "ms_per_step": dt * 1e3, "ndt_registrations_per_s_e2e": a.pairs / dt, "ok": sum(r.status == 0 for r in res),
       "outer_iterations": [r.outer_iterations for r in res[: a.distinct]], "line_iterations": [r.inner_iterations for r in res[: a.distinct]]}
out["stage_ms_per_step"] = {k: round(v["ms"] / a.steps, 3) for k, v in ctx.stage_times(reset=True).items()}  # stage ids of s3d_b200.h (NDT: knn_cov = voxel Gaussians)
if a.oracle:
    import oracle
    t0 = time.perf_counter()
    r = oracle.gicp_align(srcs[0], tgts[0], None, P)
    out["oracle_ms_per_align_1thread"] = (time.perf_counter() - t0) * 1e3
    out["oracle_outer"] = r.outer_iterations
print(json.dumps(out))

"""GPU box: the GICP loop alone (s3d_gicp_align_prepared_batch on clouds prepared once) on the bench workload, host-clock timed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, slam3d_b200, bench
ctx = slam3d_b200.Context([0])
pairs = bench.make_pairs(16)
p = bench.params()
src = ctx.prepare_clouds([slam3d_b200.as_xyzw(pairs[i % 16][0]) for i in range(64)], bench.VOXEL, 20)
tgt = ctx.prepare_clouds([slam3d_b200.as_xyzw(pairs[i % 16][1]) for i in range(64)], bench.VOXEL, 20)
for _ in range(3):
    ctx.gicp_align_prepared_batch(src, tgt, None, p)
torch.cuda.synchronize()
n = 10
t0 = time.perf_counter()
for _ in range(n):
    r = ctx.gicp_align_prepared_batch(src, tgt, None, p)
dt = (time.perf_counter() - t0) / n
print(f"loop only: {1e3 * dt:.2f} ms per 64 pairs  ({sum(x.outer_iterations for x in r) / 64:.2f} outer iterations, ok {sum(1 for x in r if x.status == 0)})  "
      f"env streams={os.environ.get('S3D_STREAMS_PER_DEVICE')} mode={os.environ.get('S3D_LOOP_MODE')} ctas={os.environ.get('S3D_LOOP_CTAS_PER_SM')} lib={os.path.basename(slam3d_b200.LIB_PATH)}")

"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the path once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ.setdefault("S3D_WATCHDOG_MCYCLES", "100000000")  # the loop kernel's watchdog counts cycles; under the sanitizer everything is 100x slower
import slam3d_b200
from slam3d_b200 import synth
from slam3d_b200._abi import RegistrationParameters

src, tgt, _ = synth.scan_pair(seed=3)
src, tgt = src[::8], tgt[::8]
ctx = slam3d_b200.Context()
out, li, ov = ctx.voxel_downsample(src, 0.3)
dense, _, _ = ctx.voxel_downsample(np.tile(src[:3000], (6, 1)), 2.0)                                    # voxels with hundreds of points: the long-voxel kernel
idx, d2, cov = ctx.knn_covariances(out, 20)
nn_i, nn_d = ctx.nearest_neighbors(out, out[::3])
res = ctx.gicp_align_batch([src, tgt, src[:50]], [tgt, src, tgt], None, RegistrationParameters.defaults(point_cloud_density=0.3))
one = ctx.gicp_align(src, tgt, None, RegistrationParameters.defaults(point_cloud_density=0.3))          # the persistent loop kernel (single call)
comb = ctx.combined_measurement([src, tgt], [np.eye(4), np.eye(4)], np.eye(4))
big_k = ctx.knn_covariances(out[::4], 210)                                                              # heap in global memory
print("sanitize smoke:", out.shape, dense.shape, idx.shape, [(r.status, r.outer_iterations) for r in res], (one.status, one.outer_iterations), comb.shape, big_k[0].shape)
ctx.close()

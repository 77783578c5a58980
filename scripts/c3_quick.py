"""VoxelGrid on the 2M-point map cloud (BASELINE configs[2]), voxel stage only: ms per call for each leaf size (CUDA events of the
library, 10 calls after 3 warm-ups).  The quick A/B loop for the streaming kernels; `bench.py --workload c3` is the reported line."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, slam3d_b200
from slam3d_b200 import synth
ctx = slam3d_b200.Context()
cloud = synth.map_cloud(n_scans=16)
dev = torch.from_numpy(slam3d_b200.as_xyzw(cloud)).cuda()
ctx.set_profiling(True)
ref = {}
for leaf in (0.05, 0.1, 0.2):
    for _ in range(3):
        out, _, _ = ctx.voxel_downsample(dev, leaf, want_leaf_index=False)
    ctx.stage_times(reset=True)
    for _ in range(10):
        out, _, _ = ctx.voxel_downsample(dev, leaf, want_leaf_index=False)
    st = ctx.stage_times(reset=True)
    print(f"leaf {leaf}: {st['voxel']['ms'] / 10 * 1e3:.1f} us  voxels {out.shape[0]}  launches {st['voxel']['launches'] // 10}  checksum {float(np.asarray(out, np.float64).sum()):.6f}")

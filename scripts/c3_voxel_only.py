import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, slam3d_b200
from slam3d_b200 import synth
ctx = slam3d_b200.Context()
cloud = synth.map_cloud(n_scans=16)
dev = torch.from_numpy(slam3d_b200.as_xyzw(cloud)).cuda()
for leaf in (0.05, 0.1, 0.2):
    for _ in range(2):
        out, _, _ = ctx.voxel_downsample(dev, leaf, want_leaf_index=False)

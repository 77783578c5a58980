"""BASELINE config 3: VoxelGrid + covariance stress on the 2 097 152-point synthetic map cloud (device-resident input)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, slam3d_b200
from slam3d_b200 import synth
ctx = slam3d_b200.Context()
cloud = synth.map_cloud(n_scans=16)
dev = torch.from_numpy(slam3d_b200.as_xyzw(cloud)).cuda()
ctx.set_profiling(True)
for leaf in (0.05, 0.1, 0.2):
    for _ in range(3):
        out, _, _ = ctx.voxel_downsample(dev, leaf, want_leaf_index=False)
    ctx.stage_times(reset=True)
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        out, _, _ = ctx.voxel_downsample(dev, leaf, want_leaf_index=False)
    wall = (time.perf_counter() - t0) / n
    st = ctx.stage_times(reset=True)
    ms = st["voxel"]["ms"] / n
    M = out.shape[0]
    algo = 16.0 * cloud.shape[0] + 16.0 * M
    print(f"leaf {leaf}: M={M} voxel kernels {ms*1e3:.0f} us ({st['voxel']['launches']//n} launches) wall {wall*1e3:.2f} ms (incl. D2H of the result)  algorithmic {algo/1e6:.1f} MB -> {algo/ms/1e6:.0f} GB/s = {algo/ms/1e6/6444.4*100:.1f}% of measured HBM peak")
    fd = torch.from_numpy(out).cuda()
    for _ in range(2):
        ctx.knn_covariances(fd, 20)
    ctx.stage_times(reset=True)
    for _ in range(3):
        ctx.knn_covariances(fd, 20)
    st = ctx.stage_times(reset=True)
    print(f"          kNN-20+cov on M={M}: grid {st['grid']['ms']/3*1e3:.0f} us, knn_cov {st['knn_cov']['ms']/3*1e3:.0f} us -> {M/(st['knn_cov']['ms']/3)/1e3:.1f} M queries/s")

# ---- map building (buildMap): 16 scans -> accumulate + RadiusOutlierRemoval(0.2, 3) + VoxelGrid(0.1) ----
rng = np.random.default_rng(20260117)
scene = synth.Scene(20260117)
scans, poses = [], []
for i in range(16):
    P = synth.make_pose([12.0 * i / 15 - 6.0, 0.0, 0.0], [0.0, 0.0, 0.02 * i])
    scans.append(synth.scan(scene, P, rng)); poses.append(P)
dev_scans = [torch.from_numpy(slam3d_b200.as_xyzw(s)).cuda() for s in scans]
for _ in range(2):
    m = ctx.build_map(dev_scans, poses, 0.2, 3, 0.1)
t0 = time.perf_counter()
for _ in range(5):
    m = ctx.build_map(dev_scans, poses, 0.2, 3, 0.1)
dt = (time.perf_counter() - t0) / 5
import oracle
t0 = time.perf_counter(); mo = oracle.build_map(scans, poses, 0.2, 3, 0.1); to = time.perf_counter() - t0
print(f"buildMap 16 scans (2 097 152 points) -> {m.shape[0]} map points: {dt*1e3:.2f} ms per map on the GPU (device-resident scans, result to host); oracle {to*1e3:.0f} ms; identical: {np.array_equal(m, mo)}")

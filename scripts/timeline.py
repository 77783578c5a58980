"""GPU box: stage time stamps of every chunk of one batch call (S3D_TIMELINE=1), device-resident inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["S3D_TIMELINE"] = "1"
import torch, slam3d_b200, bench
ctx = slam3d_b200.Context([0])
pairs = bench.make_pairs(16)
src = [torch.from_numpy(slam3d_b200.as_xyzw(pairs[i % 16][0])).cuda() for i in range(64)]
tgt = [torch.from_numpy(slam3d_b200.as_xyzw(pairs[i % 16][1])).cuda() for i in range(64)]
p = bench.params()
for it in range(4):
    sys.stderr.write(f"--- call {it}\n")
    ctx.gicp_align_batch(src, tgt, None, p)

"""GPU box: stage time stamps of every chunk of one batch call (S3D_TIMELINE=1), device-resident inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["S3D_TIMELINE"] = "1"
import torch, slam3d_b200, bench
ctx = slam3d_b200.Context([0])
pairs = bench.make_pairs(16)
host = len(sys.argv) > 1 and sys.argv[1] == "host"   # pinned host scans (the e2e path) instead of device-resident ones
src = [torch.from_numpy(slam3d_b200.as_xyzw(pairs[i % 16][0])).pin_memory() for i in range(64)]
tgt = [torch.from_numpy(slam3d_b200.as_xyzw(pairs[i % 16][1])).pin_memory() for i in range(64)]
if not host:
    src, tgt = [a.cuda() for a in src], [a.cuda() for a in tgt]
p = bench.params()
for it in range(4):
    sys.stderr.write(f"--- call {it}\n")
    ctx.gicp_align_batch(src, tgt, None, p)

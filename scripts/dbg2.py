import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, slam3d_b200, oracle
from slam3d_b200 import synth
from slam3d_b200._abi import RegistrationParameters
from conftest import pose_delta
ctx=slam3d_b200.Context()
coarse = RegistrationParameters.defaults(point_cloud_density=0.5, max_correspondence_distance=5.0, max_translation=5.0)
for i in range(4):
    s,t,truth=synth.scan_pair(seed=100+i, loop=True); s=s[::2]; t=t[::2]
    print('=== pair',i, file=sys.stderr)
    g=ctx.gicp_align(s,t,None,coarse)
    o=oracle.gicp_align(s,t,None,coarse)
    print('PAIR',i,'gpu',g.status,g.outer_iterations,g.inner_iterations,'oracle',o.status,o.outer_iterations,o.inner_iterations, pose_delta(o.pose(),g.pose()), pose_delta(truth,g.pose()), pose_delta(truth,o.pose()), file=sys.stderr)

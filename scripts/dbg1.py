import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, slam3d_b200, oracle
from conftest import load_kitti
from slam3d_b200._abi import RegistrationParameters
k=[load_kitti(i) for i in range(1,5)]
ctx=slam3d_b200.Context()
p=RegistrationParameters.defaults(point_cloud_density=0.2)
for a,b in ((0,1),(1,2),(2,3)):
    r=ctx.gicp_align(k[a],k[b],None,p)
    o=oracle.gicp_align(k[a],k[b],None,p)
    print('PAIR',a,b,'gpu iters',r.outer_iterations,r.inner_iterations,'oracle',o.outer_iterations,o.inner_iterations, r.pose()[:3,3], o.pose()[:3,3])

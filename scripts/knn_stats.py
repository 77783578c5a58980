"""Builds a stats-instrumented copy of the library (-DS3D_KNN_STATS) and prints the per-query work of the kNN kernel."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
csrc = os.path.join(ROOT, "slam3d_b200", "csrc")
out = "/tmp/libs3d_stats.so"
srcs = [os.path.join(csrc, f) for f in ("voxel.cu", "grid.cu", "knn.cu", "gicp.cu", "ndt.cu", "map.cu", "api.cu")]
prebuilt = os.path.join(ROOT, "slam3d_b200", "build", "libs3d_stats.so")  # built on the CPU box, travels with the snapshot
if os.path.exists(prebuilt):
    out = prebuilt
else:
  subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false", "-DS3D_KNN_STATS", "-Xcompiler", "-fPIC",
                       "-shared", "-o", out] + srcs + ["-lcudart"])
import slam3d_b200
slam3d_b200.LIB_PATH = out
from conftest import load_kitti
import oracle
ctx = slam3d_b200.Context()
lib = slam3d_b200.lib()
names = ["queries", "level scans", "cells probed", "cells pruned", "candidates", "heap pushes", "sift-downs", "start-level probes"]
from slam3d_b200 import synth
clouds = {"kitti1@0.1": oracle.voxel_downsample(load_kitti(1), 0.1)[0], "synth@0.1": oracle.voxel_downsample(synth.scan_pair()[0], 0.1)[0]}
for name, f in clouds.items():
    z = (C.c_ulonglong * 8)()
    lib.s3d_debug_knn_stats(z, 1)
    ctx.knn_covariances(f, 20)
    lib.s3d_debug_knn_stats(z, 1)
    q = z[0]
    print(name, f.shape[0], {n: round(z[i] / q, 2) for i, n in enumerate(names)})

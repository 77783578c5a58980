#!/usr/bin/env python
"""Issue-slot model of the search walks, from the work records of the HOST build of the product's search headers
(tests/hostsearch.cpp; no GPU).  For every warp of 32 consecutive threads it prices three schedules of the very same cell scans:

  lockstep   every lane steps through the 27 cell slots together (scan_block<false>, thread_walk): a slot costs the warp as
             many candidate steps as its busiest lane needs;
  two-pass   slots are only noted in lockstep, then every lane scans its own cells back to back (scan_block<true>): the warp
             needs as many candidate steps as its busiest lane has in total;
  balanced   the noted cell ranges of the warp are dealt out evenly over its 32 lanes: total / 32.

Units: candidate steps per warp (one step = one point examined by up to 32 lanes).  The 27-slot overhead is the same for all
three and is reported separately (slot visits per warp)."""
import argparse, ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from slam3d_b200 import synth
import test_hostsearch as T

ap = argparse.ArgumentParser()
ap.add_argument("--density", type=float, default=0.1)
args = ap.parse_args()
import subprocess
src = os.path.join(ROOT, "tests", "hostsearch.cpp"); out = os.path.join(ROOT, "tests", "_hostsearch.so")  # as the test fixture builds it
subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-I" + T.CUDA_INC, "-I" + os.path.join(ROOT, "tests"), "-o", out, src])
lib = C.CDLL(out)
lib.hs_build_grid.restype = C.c_void_p; lib.hs_build_grid.argtypes = [C.c_void_p, C.c_uint64, C.c_float]
lib.hs_free_grid.argtypes = [C.c_void_p]; lib.hs_levels.argtypes = [C.c_void_p]
lib.hs_nn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_int] + [C.c_void_p] * 5
lib.hs_knn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
lib.hs_trace_size.restype = C.c_uint64
lib.hs_trace_end.argtypes = [C.c_void_p] * 4


def traced(fn):
    lib.hs_trace_begin()
    res = fn()
    n = lib.hs_trace_size()
    q, v, s, c = (np.empty(n, np.uint32) for _ in range(4))
    lib.hs_trace_end(T.ptr(q), T.ptr(v), T.ptr(s), T.ptr(c))
    return res, q, v, s, c


def price(name, nq, q, v, s, c, per_warp=32):
    nw = (nq + per_warp - 1) // per_warp
    w = q // per_warp
    vmax = int(v.max()) + 1 if len(v) else 1
    # lockstep: per (warp, visit, slot) the busiest lane
    key = (w.astype(np.int64) * vmax + v) * 27 + s
    lane_work = np.zeros((nw * vmax * 27,), np.int64)
    np.maximum.at(lane_work, key, c)  # one record per (query, visit, slot), so the max over the records of a key is the max over lanes
    lock = lane_work.reshape(nw, -1).sum(1)
    # two-pass: own cell (slot 0) in lockstep, the rest per lane back to back, per visit
    own = np.zeros((nw * vmax,), np.int64); np.maximum.at(own, w.astype(np.int64) * vmax + v, np.where(s == 0, c, 0))
    rest_lane = np.zeros((nq * vmax,), np.int64); np.add.at(rest_lane, q.astype(np.int64) * vmax + v, np.where(s == 0, 0, c))
    rest_lane = np.pad(rest_lane.reshape(nq, vmax), ((0, nw * per_warp - nq), (0, 0))).reshape(nw, per_warp, vmax)
    two = own.reshape(nw, vmax).sum(1) + rest_lane.max(1).sum(1)
    tot = np.zeros(nw, np.int64); np.add.at(tot, w, c)
    bal = own.reshape(nw, vmax).sum(1) + (rest_lane.sum(1).sum(1) + per_warp - 1) // per_warp
    slots = np.zeros(nw, np.int64); np.add.at(slots, w, 1)
    print(f"{name:34s} points/query {c.sum()/nq:7.1f}  cells/query {len(c)/nq:5.2f} | candidate steps per warp: lockstep {lock.mean():7.1f}  two-pass {two.mean():7.1f}  "
          f"balanced {bal.mean():7.1f}  (ideal {tot.mean()/per_warp:6.1f}) | lane use lockstep {tot.sum()/per_warp/lock.sum():.2f} two-pass {tot.sum()/per_warp/two.sum():.2f}")


src, tgt, truth = synth.scan_pair()
fa = oracle.voxel_downsample(tgt, args.density)[0][:, :3]   # moving cloud A (slam3d target)
fb = oracle.voxel_downsample(src, args.density)[0][:, :3]   # fixed cloud B
gb = T.Grid(lib, fb, args.density)
ga = T.Grid(lib, fa, args.density)
# queries in the Morton order of their own cloud, as the kernels see them: the 1-NN of A in A gives every point's sorted position
idx_self, _, _, pos_self = ga.nn(fa)
perm = np.empty(len(fa), np.int64); perm[pos_self] = np.arange(len(fa))  # sorted position -> original index
A = fa[perm]
(res1), q, v, s, c = traced(lambda: gb.nn(A, cutoff2=6.25, gather=True))
price("1-NN, outer iteration 1 (no hint)", len(A), q, v, s, c)
T_true = np.asarray(truth, np.float64)
half = A + np.float32(0.5) * (A @ T_true[:3, :3].T.astype(np.float32) + T_true[:3, 3].astype(np.float32) - A)  # half-way to the solution
(res2), q, v, s, c = traced(lambda: gb.nn(half, cutoff2=6.25, gather=True, hints=res1[3]))
price("1-NN, later iteration (hinted)", len(A), q, v, s, c)
(_), q, v, s, c = traced(lambda: ga.knn(20))
price("kNN-20 (own cloud)", len(fa), q, v, s, c)

#!/usr/bin/env python
"""BASELINE.json configs[4]: a synthetic trajectory of N scans (closed figure-eight, 0.7 m per frame, 131 072 points
per scan) through the unchanged caller logic of slam3d — ScanSensor::addMeasurement(m, odom) (link to previous) followed by
linkLastToNeighbors() (loop closures, coarse + fine align) — as mirrored by slam3d_b200/host/MiniHost.hpp, with the CUDA path
behind PointCloudSensor::createConstraint.  g2o is not installed here: edges go to the recording graph.

The figure-eight lap has 27 poses; a lap is revisited with a sideways offset (two lanes), so only 54 ray casts are needed: they
are cached under .cache_c5/ (noise-free ranges) and every visit draws fresh range noise.

    python scripts/run_c5.py --scans 2000
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=2000)
    ap.add_argument("--build-cache-only", action="store_true")
    ap.add_argument("--patch-range", type=int, default=0, help="Sensor::mPatchBuildingRange: loop closures match patches of scans (SURVEY 8d: 0 then 2)")
    ap.add_argument("--batched", action="store_true", help="candidates of a vertex matched as one device batch (PointCloudSensor::createConstraints)")
    ap.add_argument("--max-links", type=int, default=1)
    ap.add_argument("--radius", type=float, default=1.0)
    args = ap.parse_args()
    from slam3d_b200 import synth
    radius, step = 1.5, 0.7
    lap = int(round(4 * np.pi * radius / step))  # 27 frames; the lap is closed by construction of the pose list below
    step = 4 * np.pi * radius / lap
    cache = os.path.join(ROOT, ".cache_c5", "ranges.npy")
    lanes = [0.0, 0.25]
    lane_poses = [synth.figure_eight_poses(lap, radius, step, lateral=l) for l in lanes]
    if os.path.exists(cache):
        rng_cache = np.load(cache)
    else:
        scene = synth.Scene(20260117)
        rng_cache = np.stack([np.stack([synth.ranges(scene, p) for p in lp]) for lp in lane_poses])
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        np.save(cache, rng_cache)
    if args.build_cache_only:
        print("cache", rng_cache.shape)
        return
    import slam3d_b200
    import test_gpu_host as th
    from slam3d_b200._abi import RegistrationParameters
    host = th.load_host()
    host.s3dhost_run_trajectory2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    n = args.scans
    rng = np.random.default_rng(7)
    truth, scans = [], []
    for i in range(n):
        lane = (i // lap) % len(lanes)
        truth.append(lane_poses[lane][i % lap])
        scans.append(slam3d_b200.as_xyzw(synth.scan_from_ranges(rng_cache[lane, i % lap], rng)))
    # odometry: the true motion with 2 % translation noise and 0.2 deg heading noise per frame
    odoms = [np.eye(4)]
    for i in range(1, n):
        rel = np.linalg.inv(truth[i - 1]) @ truth[i]
        rel = rel @ synth.make_pose(rng.normal(0, 0.014, 3) * [1, 1, 0.2], np.deg2rad(rng.normal(0, 0.2, 3) * [0.2, 0.2, 1]))
        odoms.append(odoms[-1] @ rel)
    sensor = host.s3dhost_sensor_create(b"velodyne")
    fine = RegistrationParameters.defaults(point_cloud_density=0.1)
    coarse = RegistrationParameters.defaults(point_cloud_density=0.5, max_correspondence_distance=5.0)
    host.s3dhost_sensor_set_params(sensor, C.byref(fine), 0)
    host.s3dhost_sensor_set_params(sensor, C.byref(coarse), 1)
    ptrs = (C.c_void_p * n)(*[s.ctypes.data for s in scans])
    sizes = (C.c_uint64 * n)(*[s.shape[0] for s in scans])
    od = np.ascontiguousarray(np.stack([o.T for o in odoms]))
    max_edges = 4 * n
    edges = np.zeros((max_edges, 3), np.int32); T = np.zeros((max_edges, 16)); poses = np.zeros((n, 16)); nw = C.c_int(0)
    t0 = time.perf_counter()
    ne = host.s3dhost_run_trajectory2(sensor, ptrs, sizes, n, od.ctypes.data, args.radius, args.max_links, 10, args.patch_range, int(args.batched), max_edges,
                                      edges.ctypes.data, T.ctypes.data, poses.ctypes.data, C.byref(nw))
    dt = time.perf_counter() - t0
    assert ne >= 0, host.s3dhost_last_message()
    edges = edges[:ne]; T = T[:ne].reshape(ne, 4, 4).transpose(0, 2, 1)
    err_t = []
    for (s, t, loop), rel in zip(edges, T):
        want = np.linalg.inv(truth[s]) @ truth[t]
        err_t.append(float(np.linalg.norm((np.linalg.inv(want) @ rel)[:3, 3])))
    P = poses.reshape(n, 4, 4).transpose(0, 2, 1)
    drift = float(np.linalg.norm((np.linalg.inv(truth[0]) @ truth[-1])[:3, 3] - P[-1][:3, 3]))
    loops = int(edges[:, 2].sum())
    print(json.dumps({"workload": "BASELINE configs[4]: figure-eight trajectory, addMeasurement(m, odom) + linkLastToNeighbors()", "scans": n,
                      "patch_building_range": args.patch_range, "batched_links": bool(args.batched), "max_neighbor_links": args.max_links, "neighbor_radius": args.radius,
                      "seconds": dt, "scans_per_s": n / dt, "edges": int(ne), "odometry_edges": int(ne - loops), "loop_edges": loops,
                      "aligns": int((ne - loops) + 2 * loops), "warnings": int(nw.value), "last_warning": host.s3dhost_last_message().decode(),
                      "edge_translation_error_m": {"median": float(np.median(err_t)), "max": float(np.max(err_t))},
                      "chained_pose_drift_m_after_last_scan": drift, "cache_hits": int(host.s3dhost_cache_hits())}))
    host.s3dhost_sensor_destroy(sensor)


if __name__ == "__main__":
    main()

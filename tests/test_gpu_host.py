"""The C++ host mirror of slam3d::PointCloudSensor (slam3d_b200/host) driven the way slam3d drives the reference:
createConstraint with sensor poses, NoMatch / BadMeasurementType / runtime_error mapping, link-to-previous through the
mini-host, and concurrent entry from two threads (ScanSensor.cpp:209-210)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import pose_delta
from slam3d_b200 import _abi
from slam3d_b200._abi import RegistrationParameters

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_host():
    import slam3d_b200
    slam3d_b200.lib()
    lib = C.CDLL(os.path.join(ROOT, "slam3d_b200", "libs3d_host.so"))
    lib.s3dhost_last_message.restype = C.c_char_p
    lib.s3dhost_sensor_create.restype = C.c_void_p
    lib.s3dhost_sensor_destroy.argtypes = [C.c_void_p]
    lib.s3dhost_sensor_set_params.argtypes = [C.c_void_p, C.POINTER(RegistrationParameters), C.c_int]
    lib.s3dhost_sensor_set_covariance_scale.argtypes = [C.c_void_p, C.c_double]
    lib.s3dhost_create_constraint.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                              C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.s3dhost_downsample.restype = C.c_int64
    lib.s3dhost_downsample.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_void_p]
    lib.s3dhost_build_map.restype = C.c_int64
    lib.s3dhost_build_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_uint, C.c_void_p]
    lib.s3dhost_set_cache_capacity.argtypes = [C.c_uint64]
    lib.s3dhost_cache_hits.restype = C.c_uint64
    lib.s3dhost_run_odometry.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    return lib


@pytest.fixture(scope="module")
def host():
    return load_host()


def cm(T):
    return np.ascontiguousarray(np.asarray(T, np.float64).T)


def rot_z(a):
    T = np.eye(4); T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
    return T


def create_constraint(host, sensor, src, tgt, src_pose, tgt_pose, odom, loop=False, bad=False):
    import slam3d_b200
    s, t = slam3d_b200.as_xyzw(src), slam3d_b200.as_xyzw(tgt)
    T = np.zeros(16); info = np.zeros(36)
    sp, tp, od = cm(src_pose), cm(tgt_pose), cm(odom)
    st = host.s3dhost_create_constraint(sensor, s.ctypes.data, s.shape[0], sp.ctypes.data, t.ctypes.data, t.shape[0], tp.ctypes.data,
                                        od.ctypes.data, int(loop), int(bad), T.ctypes.data, info.ctypes.data)
    return st, T.reshape(4, 4).T.copy(), info.reshape(6, 6).T.copy(), host.s3dhost_last_message().decode()


def test_create_constraint_frames_and_information(host, oracle_mod, kitti):
    sensor = host.s3dhost_sensor_create(b"velodyne")
    fine = RegistrationParameters.defaults(point_cloud_density=0.5)
    host.s3dhost_sensor_set_params(sensor, C.byref(fine), 0)
    host.s3dhost_sensor_set_covariance_scale(sensor, 0.25)
    src, tgt = kitti[0][::2], kitti[1][::2]
    sensor_pose = rot_z(0.3); sensor_pose[:3, 3] = [1.0, 0.2, 1.7]      # lidar mounted on the robot
    odom = np.eye(4); odom[:3, 3] = [0.6, 0.0, 0.0]
    odom = sensor_pose @ odom @ np.linalg.inv(sensor_pose)               # robot-frame odometry
    st, T, info, msg = create_constraint(host, sensor, src, tgt, sensor_pose, sensor_pose, odom)
    assert st == 0, msg
    # expected: PointCloudSensor.cpp:274,292,295 with the oracle's align()
    guess = np.linalg.inv(sensor_pose) @ odom @ sensor_pose
    want = oracle_mod.gicp_align(src, tgt, guess, fine)
    assert want.status == 0
    expect = sensor_pose @ want.pose() @ np.linalg.inv(sensor_pose)
    dt, dr = pose_delta(expect, T)
    assert dt < 1e-4 and dr < 1e-4
    assert np.allclose(info, np.eye(6) / 0.25)                           # :296-298  (I * scale)^-1
    host.s3dhost_sensor_destroy(sensor)


def test_exception_mapping(host, kitti):
    sensor = host.s3dhost_sensor_create(b"velodyne")
    src, tgt = kitti[0][::4], kitti[1][::4]
    I = np.eye(4)
    cases = [
        (RegistrationParameters.defaults(point_cloud_density=20.0), 1, "Too few points after filtering"),
        (RegistrationParameters.defaults(point_cloud_density=0.5, max_fitness_score=1e-6), 1, "ICP failed with Fitness-Score"),
        (RegistrationParameters.defaults(point_cloud_density=0.5, max_translation=0.05), 1, "ICP result is to far away from guess"),
        (RegistrationParameters.defaults(point_cloud_density=0.5, registration_algorithm=_abi.ALG_GICP_OMP), 0, ""),  # runs the GICP branch
        (RegistrationParameters.defaults(point_cloud_density=0.5, registration_algorithm=_abi.ALG_ICP), 3, "Unknown registration algorithm"),
    ]
    for p, code, text in cases:
        host.s3dhost_sensor_set_params(sensor, C.byref(p), 0)
        st, T, info, msg = create_constraint(host, sensor, src, tgt, I, I, I)
        assert st == code and text in msg, (st, msg)
    st, T, info, msg = create_constraint(host, sensor, src, tgt, I, I, I, bad=True)
    assert st == 2  # BadMeasurementType (:279-283)
    host.s3dhost_sensor_destroy(sensor)


def test_loop_closure_runs_coarse_then_fine(host, oracle_mod):
    from slam3d_b200 import synth
    src, tgt, truth = synth.scan_pair(seed=5, loop=True)
    src, tgt = src[::2], tgt[::2]
    sensor = host.s3dhost_sensor_create(b"velodyne")
    coarse = RegistrationParameters.defaults(point_cloud_density=0.5, max_correspondence_distance=5.0, max_translation=5.0)
    fine = RegistrationParameters.defaults(point_cloud_density=0.2, max_translation=5.0)
    host.s3dhost_sensor_set_params(sensor, C.byref(coarse), 1)
    host.s3dhost_sensor_set_params(sensor, C.byref(fine), 0)
    I = np.eye(4)
    st, T, info, msg = create_constraint(host, sensor, src, tgt, I, I, I, loop=True)
    oc = oracle_mod.gicp_align(src, tgt, None, coarse)
    of = oracle_mod.gicp_align(src, tgt, oc.pose(), fine) if oc.status == 0 else oc
    assert (st == 0) == (of.status == 0), msg
    if st == 0:
        dt, dr = pose_delta(of.pose(), T)
        assert dt < 1e-4 and dr < 1e-4
        dt, dr = pose_delta(truth, T)
        assert dt < 0.03 and dr < 5e-3
    host.s3dhost_sensor_destroy(sensor)


def test_minihost_odometry_and_two_threads(host, oracle_mod, kitti):
    import slam3d_b200
    sensor = host.s3dhost_sensor_create(b"velodyne")
    fine = RegistrationParameters.defaults(point_cloud_density=0.5)
    host.s3dhost_sensor_set_params(sensor, C.byref(fine), 0)
    scans = [slam3d_b200.as_xyzw(k[::2]) for k in kitti]
    ptrs = (C.c_void_p * 4)(*[s.ctypes.data for s in scans])
    sizes = (C.c_uint64 * 4)(*[s.shape[0] for s in scans])
    odoms = np.zeros((4, 4, 4))
    for i in range(4):
        o = np.eye(4); o[0, 3] = 0.65 * i
        odoms[i] = o.T  # column-major
    out = np.zeros((3, 16)); nw = C.c_int(0)
    host.s3dhost_set_cache_capacity(0)                  # raw path: every align() preprocesses both clouds
    n0 = host.s3dhost_run_odometry(sensor, ptrs, sizes, 4, odoms.ctypes.data, 1, out.ctypes.data, C.byref(nw))
    assert n0 == 3 and nw.value == 0
    uncached = out.copy()
    host.s3dhost_set_cache_capacity(32)                 # device cache: scans 2 and 3 are preprocessed once, used twice
    hits0 = host.s3dhost_cache_hits()
    n1 = host.s3dhost_run_odometry(sensor, ptrs, sizes, 4, odoms.ctypes.data, 1, out.ctypes.data, C.byref(nw))
    assert n1 == 3 and nw.value == 0
    assert host.s3dhost_cache_hits() - hits0 == 2
    single = out.copy()
    assert np.array_equal(single, uncached)             # the cache never changes a result
    n2 = host.s3dhost_run_odometry(sensor, ptrs, sizes, 4, odoms.ctypes.data, 2, out.ctypes.data, C.byref(nw))
    assert n2 == 3                      # two concurrent threads: identical edges in both (checked inside) ...
    assert np.array_equal(out, single)  # ... and identical to the single-threaded run
    guess = np.eye(4); guess[0, 3] = 0.65
    for i in range(3):
        want = oracle_mod.gicp_align(kitti[i][::2], kitti[i + 1][::2], guess, fine)
        dt, dr = pose_delta(want.pose(), single[i].reshape(4, 4).T)
        assert want.status == 0 and dt < 1e-4 and dr < 1e-4
    host.s3dhost_sensor_destroy(sensor)


def test_host_downsample(host, oracle_mod, kitti):
    import slam3d_b200
    a = slam3d_b200.as_xyzw(kitti[2])
    out = np.zeros_like(a)
    m = host.s3dhost_downsample(a.ctypes.data, a.shape[0], 0.1, out.ctypes.data)
    eo, _, _ = oracle_mod.voxel_downsample(kitti[2], 0.1)
    assert m == eo.shape[0] and np.array_equal(out[:m].view(np.uint32), eo.view(np.uint32))
    assert host.s3dhost_downsample(a.ctypes.data, 0, 0.1, out.ctypes.data) == 0  # empty in, empty out (:193)


def test_host_build_map(host, oracle_mod, kitti):
    """PointCloudSensor::buildMap through the mirror == accumulate + removeOutliers + downsample step by step == oracle."""
    import slam3d_b200
    sensor = host.s3dhost_sensor_create(b"velodyne")
    scans = [slam3d_b200.as_xyzw(k) for k in kitti]
    ptrs = (C.c_void_p * 4)(*[s.ctypes.data for s in scans])
    sizes = (C.c_uint64 * 4)(*[s.shape[0] for s in scans])
    poses = []
    for i in range(4):
        P = np.eye(4); P[0, 3] = 0.69 * i
        poses.append(P)
    pc = np.ascontiguousarray(np.stack([p.T for p in poses]))
    out = np.zeros((sum(s.shape[0] for s in scans), 4), np.float32)
    m = host.s3dhost_build_map(sensor, ptrs, sizes, 4, pc.ctypes.data, 0.1, 0.2, 3, out.ctypes.data)
    want = oracle_mod.build_map(kitti, poses, 0.2, 3, 0.1)
    assert m == want.shape[0] and np.array_equal(out[:m].view(np.uint32), want.view(np.uint32))
    host.s3dhost_sensor_destroy(sensor)


def test_minihost_trajectory_links_to_neighbors(host):
    """BASELINE configs[4] in small: a closed figure-eight through addMeasurement(m, odom) + linkLastToNeighbors()
    (ScanSensor.cpp:94-135, :170-213): consecutive scans are linked through the odometry guess, revisited places get loop
    edges (coarse + fine align, guess = relative pose in the recorded graph), every edge agrees with the true motion."""
    from slam3d_b200 import synth
    import slam3d_b200
    host.s3dhost_run_trajectory.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lap, radius = 20, 0.8
    step = 4 * np.pi * radius / lap
    n = lap + 4
    truth = synth.figure_eight_poses(n, radius, step)
    scene = synth.Scene(3)
    rng = np.random.default_rng(3)
    scans = [slam3d_b200.as_xyzw(synth.scan(scene, p, rng, azimuth_stride=8)) for p in truth]
    odoms = [np.linalg.inv(truth[0]) @ p for p in truth]
    sensor = host.s3dhost_sensor_create(b"velodyne")
    fine = RegistrationParameters.defaults(point_cloud_density=0.2)
    coarse = RegistrationParameters.defaults(point_cloud_density=0.5, max_correspondence_distance=5.0)
    host.s3dhost_sensor_set_params(sensor, C.byref(fine), 0)
    host.s3dhost_sensor_set_params(sensor, C.byref(coarse), 1)
    ptrs = (C.c_void_p * n)(*[s.ctypes.data for s in scans])
    sizes = (C.c_uint64 * n)(*[s.shape[0] for s in scans])
    od = np.ascontiguousarray(np.stack([o.T for o in odoms]))
    edges = np.zeros((4 * n, 3), np.int32); T = np.zeros((4 * n, 16)); poses = np.zeros((n, 16)); nw = C.c_int(0)
    ne = host.s3dhost_run_trajectory(sensor, ptrs, sizes, n, od.ctypes.data, 1.0, 1, 10, 4 * n, edges.ctypes.data, T.ctypes.data,
                                     poses.ctypes.data, C.byref(nw))
    assert ne >= n - 1, host.s3dhost_last_message()
    edges = edges[:ne]; T = T[:ne].reshape(ne, 4, 4).transpose(0, 2, 1)
    odo = edges[edges[:, 2] == 0]
    assert len(odo) == n - 1 and np.array_equal(odo[:, 0] + 1, odo[:, 1])       # every scan linked to its predecessor
    loops = edges[edges[:, 2] == 1]
    assert len(loops) >= 3                                                        # the crossing and the second visit of the start
    assert all(abs(int(t) - int(s)) >= 10 for s, t, _ in loops)                   # mMinLoopLength in graph hops
    for (s, t, _), rel in zip(edges, T):
        dt, dr = pose_delta(np.linalg.inv(truth[s]) @ truth[t], rel)
        assert dt < 0.05 and dr < 0.01, (s, t, dt, dr)
    P = poses.reshape(n, 4, 4).transpose(0, 2, 1)
    dt, _ = pose_delta(np.linalg.inv(truth[0]) @ truth[-1], P[-1])              # chained corrected poses follow the path
    assert dt < 0.2
    host.s3dhost_sensor_destroy(sensor)


def _figure_eight(lap=20, radius=0.8, extra=4, seed=3):
    from slam3d_b200 import synth
    import slam3d_b200
    step = 4 * np.pi * radius / lap
    n = lap + extra
    truth = synth.figure_eight_poses(n, radius, step)
    scene = synth.Scene(seed)
    rng = np.random.default_rng(seed)
    scans = [slam3d_b200.as_xyzw(synth.scan(scene, p, rng, azimuth_stride=8)) for p in truth]
    odoms = [np.linalg.inv(truth[0]) @ p for p in truth]
    return n, truth, scans, odoms


def _run_trajectory2(host, sensor, scans, odoms, n, patch_range, batched, max_links=1, radius=1.0):
    host.s3dhost_run_trajectory2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    ptrs = (C.c_void_p * n)(*[s.ctypes.data for s in scans])
    sizes = (C.c_uint64 * n)(*[s.shape[0] for s in scans])
    od = np.ascontiguousarray(np.stack([o.T for o in odoms]))
    edges = np.zeros((6 * n, 3), np.int32); T = np.zeros((6 * n, 16)); poses = np.zeros((n, 16)); nw = C.c_int(0)
    ne = host.s3dhost_run_trajectory2(sensor, ptrs, sizes, n, od.ctypes.data, radius, max_links, 10, patch_range, int(batched), 6 * n,
                                      edges.ctypes.data, T.ctypes.data, poses.ctypes.data, C.byref(nw))
    assert ne >= n - 1, host.s3dhost_last_message()
    return edges[:ne].copy(), T[:ne].reshape(ne, 4, 4).transpose(0, 2, 1).copy(), nw.value


def test_minihost_patch_range_two_and_batched_links(host):
    """BASELINE configs[4] with Sensor::mPatchBuildingRange = 2: a loop closure matches PATCHES — the scans within two hops of
    each vertex, accumulated by createCombinedMeasurement (ScanSensor.cpp:215-270, PointCloudSensor.cpp:258-266) — and the
    candidates of a vertex go through PointCloudSensor::createConstraints as ONE device batch.  The batched run must record
    the same edges as the one-at-a-time run, bit for bit."""
    n, truth, scans, odoms = _figure_eight()
    sensor = host.s3dhost_sensor_create(b"velodyne")
    fine = RegistrationParameters.defaults(point_cloud_density=0.2)
    coarse = RegistrationParameters.defaults(point_cloud_density=0.5, max_correspondence_distance=5.0)
    host.s3dhost_sensor_set_params(sensor, C.byref(fine), 0)
    host.s3dhost_sensor_set_params(sensor, C.byref(coarse), 1)
    e_seq, T_seq, w_seq = _run_trajectory2(host, sensor, scans, odoms, n, patch_range=2, batched=False)
    e_bat, T_bat, w_bat = _run_trajectory2(host, sensor, scans, odoms, n, patch_range=2, batched=True)
    assert np.array_equal(e_seq, e_bat) and np.array_equal(T_seq, T_bat) and w_seq == w_bat
    loops = e_seq[e_seq[:, 2] == 1]
    assert len(loops) >= 2
    assert all(abs(int(t) - int(s)) >= 10 for s, t, _ in loops)   # mMinLoopLength; also > 2 * mPatchBuildingRange (:196)
    for (s, t, _), rel in zip(e_seq, T_seq):
        dt, dr = pose_delta(np.linalg.inv(truth[s]) @ truth[t], rel)
        assert dt < 0.05 and dr < 0.01, (s, t, dt, dr)
    # several links per vertex in one batch: still the edges of the sequential policy
    e_seq3, T_seq3, _ = _run_trajectory2(host, sensor, scans, odoms, n, patch_range=0, batched=False, max_links=3, radius=1.7)
    e_bat3, T_bat3, _ = _run_trajectory2(host, sensor, scans, odoms, n, patch_range=0, batched=True, max_links=3, radius=1.7)
    assert np.array_equal(e_seq3, e_bat3) and np.array_equal(T_seq3, T_bat3)
    assert len(e_seq3[e_seq3[:, 2] == 1]) >= len(loops)
    host.s3dhost_sensor_destroy(sensor)


def test_host_combined_measurement(host, oracle_mod, kitti):
    import slam3d_b200
    host.s3dhost_combined_measurement.restype = C.c_int64
    host.s3dhost_combined_measurement.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    sensor = host.s3dhost_sensor_create(b"velodyne")
    scans = [slam3d_b200.as_xyzw(k) for k in kitti[:3]]
    ptrs = (C.c_void_p * 3)(*[s.ctypes.data for s in scans])
    sizes = (C.c_uint64 * 3)(*[s.shape[0] for s in scans])
    poses = []
    for i in range(3):
        P = rot_z(0.01 * i); P[0, 3] = 0.69 * i
        poses.append(P)
    patch = poses[1]
    pc = np.ascontiguousarray(np.stack([p.T for p in poses]))
    out = np.zeros((sum(s.shape[0] for s in scans), 4), np.float32)
    m = host.s3dhost_combined_measurement(sensor, ptrs, sizes, 3, pc.ctypes.data, cm(patch).ctypes.data, out.ctypes.data)
    want = oracle_mod.combined_measurement(kitti[:3], poses, patch)
    assert m == want.shape[0] and np.array_equal(out[:m].view(np.uint32), want.view(np.uint32))
    host.s3dhost_sensor_destroy(sensor)


def test_minihost_without_odometry(host, kitti):
    """ScanSensor::addMeasurement(m) (core/ScanSensor.cpp:49-79): no odometry — the guess of a registration is the motion since
    the last vertex chained from the previous results, and Sensor::checkMinDistance decides when a scan becomes a vertex."""
    import slam3d_b200
    host.s3dhost_run_no_odometry.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    sensor = host.s3dhost_sensor_create(b"velodyne")
    fine = RegistrationParameters.defaults(point_cloud_density=0.5)
    host.s3dhost_sensor_set_params(sensor, C.byref(fine), 0)
    scans = [slam3d_b200.as_xyzw(k[::2]) for k in kitti]
    ptrs = (C.c_void_p * 4)(*[s.ctypes.data for s in scans])
    sizes = (C.c_uint64 * 4)(*[s.shape[0] for s in scans])
    added = np.zeros(4, np.int32); T = np.zeros((4, 16)); nw = C.c_int(0)
    # every scan qualifies (the KITTI scans are ~0.7 m apart): three odometry edges, each ~0.7 m forward
    ne = host.s3dhost_run_no_odometry(sensor, ptrs, sizes, 4, 0.0, 0.0, added.ctypes.data, T.ctypes.data, C.byref(nw))
    assert ne == 3 and list(added) == [1, 1, 1, 1] and nw.value == 0, host.s3dhost_last_message()
    steps = [np.linalg.norm(T[e].reshape(4, 4).T[:3, 3]) for e in range(3)]
    assert all(0.4 < s < 1.0 for s in steps), steps
    # with a 1 m threshold only every second scan becomes a vertex; the skipped motion is carried as the next guess (:62, mLastTransform)
    ne2 = host.s3dhost_run_no_odometry(sensor, ptrs, sizes, 4, 1.0, 0.5, added.ctypes.data, T.ctypes.data, C.byref(nw))
    assert list(added) == [1, 0, 1, 0] and ne2 == 1 and nw.value == 0
    assert 1.0 < np.linalg.norm(T[0].reshape(4, 4).T[:3, 3]) < 2.0
    host.s3dhost_sensor_destroy(sensor)

"""The PCL pin.  baseline/doicp_driver.cpp runs slam3d's downsample + doICP<pcl::GeneralizedIterativeClosestPoint>
(PointCloudSensor.cpp:52-82, 119-174, 190-201) with a real PCL on test/cloud1-4.bin and writes tests/golden/pcl_<version>.json.
When such a file is present these tests compare it with the oracle (CPU) and with the CUDA path (GPU):
  * VoxelGrid output (count + hash of the xyz floats), kNN-20 and 1-NN indices / squared distances: bit-exact;
  * align: same status / converged, pose within 1e-4 m / 1e-4 rad, fitness within 1e-4 relative (BASELINE.json north_star).
PCL is not installed in this repository's environment, so no such file is committed and every test here SKIPS with the reason
"parity unpinned" — which is the honest status of the oracle."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import pose_delta
from slam3d_b200._abi import RegistrationParameters

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(os.environ.get("S3D_PCL_GOLDEN_DIR", os.path.join(HERE, "golden")), "pcl_*.json")))
SKIP = "parity unpinned: no PCL golden (tests/golden/pcl_<version>.json); build baseline/doicp_driver where PCL >= 1.8.1 exists"


def cases():
    return [pytest.param(f, id=os.path.basename(f)) for f in FILES] or [pytest.param(None, id="no-pcl-golden")]


def _check_voxel_knn(pcl, impl, kitti, fnv):
    report = {}
    for leaf, rows in pcl["voxel"].items():
        for i, row in enumerate(rows):
            out = impl.voxel_downsample(kitti[i], float(leaf))[0]
            same = out.shape[0] == row["n_out"] and fnv(np.ascontiguousarray(out[:, :3])) == row["xyz_fnv1a"]
            report[f"voxel {leaf} cloud{i + 1}"] = same
    f1 = impl.voxel_downsample(kitti[0], 0.1)[0]
    f2 = impl.voxel_downsample(kitti[1], 0.1)[0]
    idx, d2, _ = impl.knn_covariances(f1, 20)
    k = pcl["knn"]["cloud1@0.1,k=20"]
    report["kNN-20 indices"] = fnv(idx.astype(np.uint32)) == k["index_fnv1a"]
    report["kNN-20 distances"] = fnv(d2.astype(np.float32)) == k["dist2_fnv1a"]
    nn_i, nn_d = impl.nearest_neighbors(f1, f2)
    k = pcl["knn"]["nn cloud2@0.1 -> cloud1@0.1"]
    report["1-NN indices"] = fnv(nn_i.astype(np.uint32)) == k["index_fnv1a"]
    report["1-NN distances"] = fnv(nn_d.astype(np.float32)) == k["dist2_fnv1a"]
    return report


def _check_align(pcl, impl, kitti):
    worst = (0.0, 0.0, 0.0)
    for key, g in pcl["align"].items():
        pair, density = key.split("@")
        a, b = (int(s[-1]) - 1 for s in pair.split("->"))
        r = impl.gicp_align(kitti[a], kitti[b], None, RegistrationParameters.defaults(point_cloud_density=float(density)))
        assert r.status == g["status"] and r.converged == g["converged"], key
        assert (r.n_source, r.n_target) == (g["n_source"], g["n_target"]), key
        dt, dr = pose_delta(np.array(g["T"]), r.pose())
        df = abs(r.fitness - g["fitness"]) / max(abs(g["fitness"]), 1e-12)
        assert dt < 1e-4 and dr < 1e-4 and df < 1e-4, (key, dt, dr, df)
        worst = (max(worst[0], dt), max(worst[1], dr), max(worst[2], df))
    return worst


@pytest.mark.parametrize("path", cases())
def test_oracle_matches_pcl(path, oracle_mod, kitti):
    if path is None:
        pytest.skip(SKIP)
    pcl = json.load(open(path))
    report = _check_voxel_knn(pcl, oracle_mod, kitti, oracle_mod.fnv1a)
    assert all(report.values()), {k: v for k, v in report.items() if not v}
    print("PCL", pcl["pcl_version"], "vs oracle: worst |dt|, |dr|, fitness rel.:", _check_align(pcl, oracle_mod, kitti))


@pytest.mark.gpu
@pytest.mark.parametrize("path", cases())
def test_gpu_matches_pcl(path, oracle_mod, kitti):
    if path is None:
        pytest.skip(SKIP)
    import slam3d_b200
    pcl = json.load(open(path))
    ctx = slam3d_b200.Context()
    report = _check_voxel_knn(pcl, ctx, kitti, oracle_mod.fnv1a)
    assert all(report.values()), {k: v for k, v in report.items() if not v}
    print("PCL", pcl["pcl_version"], "vs GPU: worst |dt|, |dr|, fitness rel.:", _check_align(pcl, ctx, kitti))
    ctx.close()


def test_driver_and_cmake_are_in_place():
    """The ready-to-run pieces exist and name what they must: the swap, the setter sequence, the version floor."""
    root = os.path.dirname(HERE)
    src = open(os.path.join(root, "baseline", "doicp_driver.cpp")).read()
    for needle in ("GeneralizedIterativeClosestPoint", "setInputSource(ft)", "setInputTarget(fs)", "getFitnessScore(cfg.max_correspondence_distance)",
                   "setCorrespondenceRandomness", "pcl::VoxelGrid", "guess.matrix().cast<float>()"):
        assert needle in src, needle
    cm = open(os.path.join(root, "baseline", "CMakeLists.txt")).read()
    assert "find_package(PCL 1.8.1" in cm

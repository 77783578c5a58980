"""The C-ABI library loads and exports every function include/s3d_b200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    __graft_entry__.build()
    import slam3d_b200
    return slam3d_b200


def declared_functions():
    src = open(os.path.join(ROOT, "include", "s3d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s3d_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(built):
    lib = built.lib()
    names = declared_functions()
    assert "s3d_gicp_align" in names and "s3d_voxel_downsample" in names and len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n


def test_struct_layout_matches_header(built):
    from slam3d_b200 import _abi
    assert C.sizeof(_abi.RegistrationParameters) == 112
    assert C.sizeof(_abi.Result) == 128 + 8 + 4 * 4 + 4 * 4
    assert C.sizeof(_abi.Cloud) == 16
    p = _abi.RegistrationParameters()
    built.lib().s3d_default_parameters(C.byref(p))
    d = _abi.RegistrationParameters.defaults()
    for name, _ in _abi.RegistrationParameters._fields_:
        assert getattr(p, name) == getattr(d, name), name
    # RegistrationParameters.hpp:36-97 defaults
    assert (p.registration_algorithm, p.point_cloud_density, p.max_fitness_score, p.max_translation, p.max_rotation) == (1, 0.2, 2.0, 1.0, 1.0)
    assert (p.transformation_epsilon, p.max_correspondence_distance, p.maximum_iterations, p.rotation_epsilon) == (1e-5, 2.5, 50, 2e-3)
    assert (p.correspondence_randomness, p.maximum_optimizer_iterations) == (20, 20)


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.S3DError, match="no CPU fallback"):
        built.Context()


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "slam3d_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "s3d_oracle" not in txt and "oracle/" not in txt.replace("the oracle's", ""), f


def test_host_mirror_library_loads(built):
    lib = C.CDLL(os.path.join(ROOT, "slam3d_b200", "libs3d_host.so"))
    for n in ("s3dhost_sensor_create", "s3dhost_create_constraint", "s3dhost_downsample", "s3dhost_run_odometry"):
        assert hasattr(lib, n)

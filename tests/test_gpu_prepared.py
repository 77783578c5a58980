"""Per-measurement device cache: s3d_prepare_cloud + s3d_gicp_align_prepared(_batch) must give exactly
what s3d_gicp_align gives on the raw clouds."""
import numpy as np
import pytest

from slam3d_b200 import _abi
from slam3d_b200._abi import RegistrationParameters

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import slam3d_b200
    c = slam3d_b200.Context()
    yield c
    c.close()


def same(a, b):
    assert a.status == b.status and a.converged == b.converged
    assert (a.n_source, a.n_target, a.n_correspondences, a.outer_iterations, a.inner_iterations) == \
           (b.n_source, b.n_target, b.n_correspondences, b.outer_iterations, b.inner_iterations)
    assert np.array_equal(a.pose(), b.pose()) and a.fitness == b.fitness


def test_odometry_chain_bit_identical_to_raw(ctx, kitti):
    p = RegistrationParameters.defaults(point_cloud_density=0.2)
    prepared = ctx.prepare_clouds(kitti, 0.2, 20)                      # every scan preprocessed once (one batched pass) ...
    assert [h.size for h in prepared] == [31834, 31481, 30882, 30435]  # SURVEY Appendix B voxel counts
    chain = ctx.gicp_align_prepared_batch(prepared[:-1], prepared[1:], None, p)   # ... and used as target and as source
    for i in range(3):
        same(chain[i], ctx.gicp_align(kitti[i], kitti[i + 1], None, p))
        same(chain[i], ctx.gicp_align_prepared(prepared[i], prepared[i + 1], None, p))
    guess = np.eye(4); guess[0, 3] = 0.6
    same(ctx.gicp_align_prepared(prepared[0], prepared[1], guess, p), ctx.gicp_align(kitti[0], kitti[1], guess, p))
    for h in prepared:
        h.release()


def test_one_target_many_candidates(ctx, kitti):
    """Loop-closure style: the same prepared target is matched against several sources in one batch."""
    p = RegistrationParameters.defaults(point_cloud_density=0.5)
    tgt = ctx.prepare_cloud(kitti[1][::2], 0.5, 20)
    srcs = [ctx.prepare_cloud(kitti[i][::2], 0.5, 20) for i in (0, 2, 0, 2, 0)]
    res = ctx.gicp_align_prepared_batch(srcs, [tgt] * 5, None, p)
    same(res[0], ctx.gicp_align(kitti[0][::2], kitti[1][::2], None, p))
    same(res[1], ctx.gicp_align(kitti[2][::2], kitti[1][::2], None, p))
    for i in (2, 4):
        same(res[i], res[0])
    same(res[3], res[1])


def test_gates_and_mismatch(ctx, kitti):
    import slam3d_b200
    small = ctx.prepare_cloud(kitti[0][:300], 20.0, 20)   # < 100 points after filtering
    big = ctx.prepare_cloud(kitti[1][::4], 20.0, 20)
    r = ctx.gicp_align_prepared(small, big, None, RegistrationParameters.defaults(point_cloud_density=20.0))
    assert r.status == _abi.S3D_TOO_FEW_POINTS
    empty = ctx.prepare_cloud(np.zeros((0, 3), np.float32), 0.5, 20)
    assert empty.size == 0
    a = ctx.prepare_cloud(kitti[0][::4], 0.5, 20)
    b = ctx.prepare_cloud(kitti[1][::4], 0.5, 20)
    r = ctx.gicp_align_prepared(empty, b, None, RegistrationParameters.defaults(point_cloud_density=0.5))
    assert r.status == _abi.S3D_TOO_FEW_POINTS
    with pytest.raises(slam3d_b200.S3DError, match="another point_cloud_density"):
        ctx.gicp_align_prepared(a, b, None, RegistrationParameters.defaults(point_cloud_density=0.2))
    with pytest.raises(slam3d_b200.S3DError, match="another point_cloud_density"):
        ctx.gicp_align_prepared(a, b, None, RegistrationParameters.defaults(point_cloud_density=0.5, correspondence_randomness=10))
    r = ctx.gicp_align_prepared(a, b, None, RegistrationParameters.defaults(point_cloud_density=0.5, registration_algorithm=_abi.ALG_NDT))
    assert r.status == _abi.S3D_UNKNOWN_ALGORITHM
    r = ctx.gicp_align_prepared(a, b, None, RegistrationParameters.defaults(point_cloud_density=0.5, max_translation=0.05))
    assert r.status == _abi.S3D_TOO_FAR_FROM_GUESS
    same(ctx.gicp_align_prepared(a, b, None, RegistrationParameters.defaults(point_cloud_density=0.5)),
         ctx.gicp_align(kitti[0][::4], kitti[1][::4], None, RegistrationParameters.defaults(point_cloud_density=0.5)))
    # released blocks are recycled by later prepares
    a.release(); b.release()
    c = ctx.prepare_cloud(kitti[2][::4], 0.5, 20)
    d = ctx.prepare_cloud(kitti[3][::4], 0.5, 20)
    same(ctx.gicp_align_prepared(c, d, None, RegistrationParameters.defaults(point_cloud_density=0.5)),
         ctx.gicp_align(kitti[2][::4], kitti[3][::4], None, RegistrationParameters.defaults(point_cloud_density=0.5)))

"""The steps either side of the hot path on the GPU against the numpy oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_preprocess_points(oracle):
    import torch
    from objectdetection_3d_b200 import ops_numpy, synth
    pts = synth.forest_tile(n=200_000, seed=41)
    pts = np.concatenate([pts, np.arange(len(pts), dtype=np.float32)[:, None]], axis=1)     # a 5th feature = row id
    pts[::1013, :3] += 300.0
    pts[5::2003, 2] -= 3.0                                           # below the range
    rg = [0, 0, 0, 40.0, 40.0, 30.0]
    ref = oracle.preprocess_points(pts, rg, [0, 1, 2, 3])
    got = ops_numpy.preprocess_points(pts, rg, [0, 1, 2, 3])
    # the statistics follow numpy's float32 order of operations (sequential column sums, pairwise 1-D sums): T0
    assert np.array_equal(ref, got)
    full_ref = oracle.preprocess_points(pts, rg, [4, 0])
    full_got = ops_numpy.preprocess_points(torch.from_numpy(pts).cuda(), rg, [4, 0]).cpu().numpy()
    assert np.array_equal(full_ref, full_got)
    assert 0 < len(got) < len(pts) - 150
    # range filter alone (no statistics): exact
    assert np.array_equal(ops_numpy.preprocess_points(pts, rg, [0, 1, 2, 3, 4], outlier_check=False),
                          oracle.preprocess_points(pts, rg, [0, 1, 2, 3, 4], outlier=False))
    assert np.array_equal(ops_numpy.global_outlier_check(pts), oracle.global_outlier_check(pts))
    assert ops_numpy.preprocess_points(pts[:0], rg, [0, 1]).shape == (0, 2)
    # the fast (float64, fully parallel) statistics: rows within rounding of the threshold may differ
    fast = ops_numpy.preprocess_points(pts, rg, [0, 1, 2, 3], exact=False)
    assert abs(len(fast) - len(ref)) <= 2


@pytest.mark.parametrize("n", [1, 5, 8, 100, 128, 129, 257, 1000, 4097, 65_537, 1_000_003])
def test_global_outlier_check_sizes(oracle, n):
    """Every shape of numpy's pairwise tree (n < 8, one block, uneven halves) and a cloud with a far cluster so that
    the threshold cuts through the data."""
    from objectdetection_3d_b200 import ops_numpy
    rng = np.random.default_rng(n)
    pts = rng.normal(15.0, 6.0, (n, 4)).astype(np.float32)
    pts[::53, :3] += rng.normal(0, 40.0, (len(pts[::53]), 3)).astype(np.float32)
    assert np.array_equal(ops_numpy.global_outlier_check(pts), oracle.global_outlier_check(pts))


def test_custom_voxelizer(oracle):
    from objectdetection_3d_b200 import model_utils
    rng = np.random.default_rng(42)
    n = 150_000
    cloud = np.concatenate([rng.uniform(0, 8, (n, 3)), rng.permutation(n)[:, None] / float(n)], axis=1).astype(np.float32)
    cfg = dict(voxel_size=[0.25, 0.25, 0.25], max_voxel_points=8, reflectance_sampling=True)
    got = model_utils.CustomVoxelizer(cfg).voxelize(cloud)
    ref = oracle.custom_voxelizer_voxelize(cloud, cfg["voxel_size"], 8, True)
    assert got.shape == ref.shape and np.array_equal(got[:, 4], ref[:, 4])          # same voxels, same order, same counts
    assert np.abs(got - ref).max() < 1e-5
    with pytest.raises(UnboundLocalError):
        model_utils.CustomVoxelizer(cfg).voxelize(cloud[:50] * 100)


def test_dense_to_sparse(oracle):
    import torch
    from objectdetection_3d_b200 import pointpillars
    rng = np.random.default_rng(43)
    x = np.zeros((3, 16, 62, 75), dtype=np.float32)
    occ = rng.random((3, 62, 75)) < 0.07
    x[:, :, :, :] = rng.normal(size=x.shape).astype(np.float32) * occ[:, None]
    x[1, 3, 10, 10] = 0.0
    x[2, :, 5, 7] = 0.0
    x[2, 9, 5, 7] = -0.0                                              # negative zero is still zero
    x[0, 15, 61, 74] = 1e-30
    v, c = pointpillars.dense_to_sparse(torch.from_numpy(x).cuda())
    ov, oc = oracle.dense_to_sparse(x)
    assert np.array_equal(c.cpu().numpy(), oc) and np.array_equal(v.cpu().numpy(), ov)
    # round trip through the dense scatter
    z = torch.zeros((1, 4, 8, 8), device="cuda")
    v0, c0 = pointpillars.dense_to_sparse(z)
    assert v0.shape == (0, 4) and c0.shape == (0, 3)

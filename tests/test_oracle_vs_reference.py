"""CPU, build container only: the oracle against the REFERENCE imported live (oracle/ref_shim.py) on
fresh seeded inputs larger than the committed fixtures.  Skipped where /root/reference is absent."""
import numpy as np
import pytest

from oracle import ref_shim

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not mounted")]


@pytest.fixture(scope="module")
def R():
    return ref_shim.load()


@pytest.mark.parametrize("kind,n,ties", [("dense", 200_000, False), ("dense", 50_000, True), ("uniform", 100_000, False)])
def test_voxelize_live(oracle, R, kind, n, ties):
    from objectdetection_3d_b200 import synth
    g = synth.G_KITTI
    pts = (synth.dense_tile(n=n, seed=21, ties=ties) if kind == "dense"
           else synth.uniform_tile(n=n, seed=22, margin=0.05))
    vs = np.array(g["voxel_size"], dtype=np.float32)
    rg = np.array(g["point_cloud_range"], dtype=np.float64)
    rv, rc, rn = R.ops_numba.points_to_voxel(pts.copy(), vs, rg, 32, 12000, True)
    v, c, n_ = oracle.points_to_voxel(pts, vs, rg, 32, 12000, True)
    assert np.array_equal(c, rc) and np.array_equal(n_, rn) and np.array_equal(v, rv)


def test_module_voxelization_live(oracle, R):
    from objectdetection_3d_b200 import synth
    g = synth.G_REF_PILLAR
    pts = synth.forest_tile(n=30_000, seed=23)
    mod = R.pp.PointPillarsVoxelization("cpu", g["voxel_size"], g["point_cloud_range"], 50, 100000)
    rv, rc, rn = mod(pts)
    v, c, n = oracle.pointpillars_voxelization(pts, g["voxel_size"], g["point_cloud_range"], 50, 100000)
    assert np.array_equal(v, rv.numpy()) and np.array_equal(c, rc.numpy()) and np.array_equal(n, rn.numpy())
    assert c.dtype == np.int64 and n.dtype == np.int64


def test_nms_live(oracle, R):
    import torch
    from objectdetection_3d_b200 import synth
    boxes, scores = synth.nms_boxes(n=1500, seed=24, extent=25.0)
    for ithr in (1e-5, 0.3):
        ref = R.utils.multiclass_nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.2, ithr, 2)[0].numpy()
        got = oracle.multiclass_nms(boxes, scores, 0.2, ithr, 2)[0]
        assert set(ref.tolist()) == set(got.tolist())
        assert (np.diff(scores[got, 0]) < 0).all()


def test_outlier_check_and_custom_voxelizer_live(oracle, R):
    """the numpy restatements against the reference's own functions."""
    import importlib
    from objectdetection_3d_b200 import synth
    ops_numpy = importlib.import_module("ops.ops_numpy")
    pts = synth.forest_tile(n=40_000, seed=31)
    pts[::997, :3] += 400.0                                             # outliers
    assert np.array_equal(oracle.global_outlier_check(pts), ops_numpy.global_outlier_check(pts))
    # CustomVoxelizer: a dense cloud (density > 10 points / m^3) so that the downsample branch runs
    rng = np.random.default_rng(32)
    cloud = np.concatenate([rng.uniform(0, 6, (60_000, 3)), rng.permutation(60_000)[:, None] / 60_000.0], axis=1).astype(np.float32)
    cfg = dict(voxel_size=[0.25, 0.25, 0.25], max_voxel_points=8, reflectance_sampling=True)
    ref = R.utils.CustomVoxelizer(cfg).voxelize(cloud.copy())
    got = oracle.custom_voxelizer_voxelize(cloud, cfg["voxel_size"], 8, True)
    assert ref.shape == got.shape and ref.shape[1] == 5 and ref.shape[0] < cloud.shape[0]
    assert np.array_equal(ref[:, 4], got[:, 4]) and np.allclose(ref, got, rtol=1e-6, atol=1e-6)
    with pytest.raises(UnboundLocalError):
        R.utils.CustomVoxelizer(cfg).voxelize(cloud[:50].copy() * 100)   # sparse cloud: the reference's unbound `vp`
    with pytest.raises(UnboundLocalError):
        oracle.custom_voxelizer_voxelize(cloud[:50] * 100, cfg["voxel_size"], 8, True)

"""CPU: the oracle (oracle/pp_oracle.c) against fixtures produced by the REFERENCE itself
(oracle/make_golden.py).  Integer outputs bit-exact (T0); FP32 outputs within 1e-5 relative (T1)."""
import numpy as np
import pytest

from conftest import assert_close_t1, golden

VOX = ["vox_model_clustered", "vox_model_overflow", "vox_model_ties", "vox_f32_boundary", "vox_f64_boundary",
       "vox_shuffle_c5", "vox_gref_forest"]


def vox_args(g):
    vs = g["voxel_size"].tolist() if bool(g["vs_is_list"]) else g["voxel_size"]
    rg = g["coors_range"].tolist() if bool(g["rg_is_list"]) else g["coors_range"]
    return vs, rg, int(g["max_points"]), int(g["max_voxels"]), bool(g["reflectance"])


@pytest.mark.parametrize("name", VOX)
def test_voxelize_bit_exact(oracle, name):
    g = golden(name)
    vs, rg, P, cap, refl = vox_args(g)
    pts = g["points"] if refl else g["points_after"]          # shuffle variant: replay the post-call order
    v, c, n = oracle.points_to_voxel(pts, vs, rg, P, cap, refl)
    assert v.shape == g["voxels"].shape
    assert np.array_equal(c, g["coors"]) and c.dtype == np.int32
    assert np.array_equal(n, g["num"]) and n.dtype == np.int32
    assert np.array_equal(v.view(np.uint32), g["voxels"].view(np.uint32))


def test_regime_kat_differs():
    """the f32 and f64 regimes put boundary points in different cells."""
    a, b = golden("vox_f32_boundary"), golden("vox_f64_boundary")
    assert a["coors"].shape != b["coors"].shape or not np.array_equal(a["coors"], b["coors"])


def test_numba_argsort_tie_order(oracle):
    g = golden("vox_model_ties")
    # the golden voxels can only be reproduced with numba's tie order; a stable order must differ
    pts = g["points"]
    stable = np.argsort(-pts[:, 3], kind="stable")
    vs, rg, P, cap, _ = vox_args(g)
    v, c, n = oracle.points_to_voxel(pts, vs, rg, P, cap, False, perm=stable)
    assert not np.array_equal(v, g["voxels"])
    perm = oracle.numba_argsort_desc(pts[:, 3])
    assert sorted(perm.tolist()) == list(range(len(pts)))
    assert (np.diff(pts[perm, 3]) <= 0).all()


@pytest.mark.parametrize("name", ["pfn_single64", "pfn_two_layer"])
def test_pfn_decorate_scatter(oracle, name):
    g = golden(name)
    vs, rg = g["voxel_size"].tolist(), g["point_cloud_range"].tolist()
    dec = oracle.decorate(g["voxels"], g["num"], g["coors"], vs[0], vs[1], vs[0] / 2 + rg[0], vs[1] / 2 + rg[1])
    # x - mean(x) cancels: the 1e-5 bound is relative to the coordinate magnitude, not to the difference
    scale = float(np.abs(g["voxels"]).max())
    assert_close_t1(dec, g["decorated"], atol=1e-5 * scale, what="decorated")
    layers = [dict(weight=g["w%d" % i], gamma=g["gamma%d" % i], beta=g["beta%d" % i], mean=g["mean%d" % i],
                   var=g["var%d" % i]) for i in range(int(g["n_layers"]))]
    out = oracle.pillar_feature_net(g["voxels"], g["num"], g["coors"], layers, vs, rg)
    assert_close_t1(out, g["out"], atol=1e-5 * scale, what="pfn out")
    assert np.array_equal(out[:, -1], g["num"].astype(np.float32))
    H, W = g["canvas_hw"].tolist()
    canvas = oracle.scatter_dense(g["out"], g["coors"].astype(np.int32), 1, 1, H, W)
    assert np.array_equal(canvas, g["canvas"])


def test_boxes_iou(oracle):
    g = golden("boxes_iou")
    rect = oracle.bbox2rotated_corners2D(g["boxes"])
    assert_close_t1(rect, g["rect"], what="aabb")
    assert_close_t1(oracle.bbox2corners3D(g["boxes"]), g["corners"], atol=2e-6, what="corners")
    # IoU on IDENTICAL rectangles is bit-exact (same op order, no contraction)
    r = g["rect"]
    assert np.array_equal(oracle.bbox_iou2D(r[:200], r[200:500]), g["iou"])
    assert np.array_equal(oracle.bbox_iou2D(r[:50], r[200:300], "iof"), g["iof"])
    assert np.array_equal(oracle.bbox_iou2D(r[:50], r[200:300], "giou"), g["giou"])
    assert np.array_equal(oracle.iou_jit(r[:40], r[300:360], 0.0), g["iou_jit"])
    assert np.array_equal(oracle.iou_jit(r[:40], r[300:360], 1.0), g["iou_jit_eps1"])
    assert (np.diag(oracle.bbox_iou2D(r[:100], r[:100])) == 1.0).all()


def test_multiclass_nms_sets(oracle):
    g = golden("nms_multiclass")
    for si, sthr in enumerate(g["score_thrs"].tolist()):
        for ii, ithr in enumerate(g["iou_thrs"].tolist()):
            keep = oracle.multiclass_nms(g["boxes"], g["scores"], sthr, ithr, 2)
            for c in range(2):
                assert np.array_equal(keep[c], g["keep_s%d_i%d_c%d" % (si, ii, c)]), (sthr, ithr, c)


def test_codec_limit_period(oracle):
    g = golden("codec")
    assert_close_t1(oracle.box_encode(g["anchors"], g["gts"]), g["encoded"], what="encode")
    assert_close_t1(oracle.box_decode(g["anchors"], g["deltas"]), g["decoded"], what="decode")
    rt = oracle.box_decode(g["anchors"], oracle.box_encode(g["anchors"], g["gts"]))
    gts = g["gts"].copy()
    gts[:, 2] += gts[:, 5] / 2          # decode returns centre-z
    assert_close_t1(rt, gts, rtol=1e-5, atol=1e-5, what="roundtrip")
    assert_close_t1(oracle.limit_period(g["val"], 1, np.pi), g["limit_1_pi"], atol=2e-6)
    assert_close_t1(oracle.limit_period(g["val"], 0.5, 2 * np.pi), g["limit_05_2pi"], atol=2e-6)


def test_anchors(oracle):
    from objectdetection_3d_b200 import synth
    g = golden("anchors")
    a = oracle.grid_anchors((5, 7), [0, 0, 0, 40.0, 40.0, 30.0], synth.ANCHOR_SIZES, synth.ANCHOR_ROTATIONS)
    assert a.shape == g["a57"].shape == (1, 5, 7, 3, 4, 9)
    assert_close_t1(a, g["a57"], what="a57")
    assert abs(a.reshape(-1, 9)[12, 0] - 6.6667) < 1e-3          # SURVEY section 4 KAT
    b = oracle.grid_anchors((31, 27), [0, -39.68, -1.78, 69.12, 39.68, -1.78], [[1.6, 3.9, 1.56]],
                            [[0, 0, 0], [0, 0, 1.57]])
    assert_close_t1(b, g["a_kitti"], what="a_kitti")

"""GPU: the CUDA path (through the C ABI of libpp_b200.so) against the golden fixtures produced by the
reference and against the CPU oracle on seeded inputs.  Integer outputs bit-exact (T0); FP32 within
1e-5 relative (T1) -- or bit-exact against the oracle where the op order is identical."""
import numpy as np
import pytest
import torch

from conftest import assert_close_t1, golden
from test_oracle_golden import VOX, vox_args

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pp():
    import objectdetection_3d_b200 as pkg
    from objectdetection_3d_b200 import _lib, model_utils, ops_numba, ops_torch, pointpillars
    _lib.load()
    assert torch.cuda.is_available()
    pkg.ops_numba, pkg.ops_torch, pkg.model_utils, pkg.pointpillars = ops_numba, ops_torch, model_utils, pointpillars
    return pkg


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


# ------------------------------------------------------------------------------------------ sort
@pytest.mark.parametrize("n", [1, 31, 500, 2048, 2049, 4096, 4097, 20_000, 100_003, 1_000_000])
def test_radix_sort_stable(pp, n):
    import ctypes
    from objectdetection_3d_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** 32 if n % 2 else 1000, size=n, dtype=np.uint64).astype(np.uint32)
    k = cu(keys.view(np.int32))
    ko, vo = torch.empty_like(k), torch.empty_like(k)
    ws = torch.empty(int(lib.pp_sort_workspace_bytes(n)), dtype=torch.uint8, device="cuda")
    rc = lib.pp_sort_pairs_u32(k.data_ptr(), None, ko.data_ptr(), vo.data_ptr(), n, ws.data_ptr(), ws.numel(),
                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(vo.cpu().numpy().view(np.uint32), order.astype(np.uint32))
    assert np.array_equal(ko.cpu().numpy().view(np.uint32), keys[order])


# ------------------------------------------------------------------------------------------ voxelize
@pytest.mark.parametrize("name", VOX)
def test_voxelize_golden(pp, oracle, name):
    g = golden(name)
    vs, rg, P, cap, refl = vox_args(g)
    has_ties = name == "vox_model_ties"
    if refl and not has_ties:
        v, c, n = pp.ops_numba.points_to_voxel(g["points"].copy(), vs, rg, P, cap, True)
    elif refl:      # replay the reference's (numba quicksort) tie order
        perm = oracle.numba_argsort_desc(g["points"][:, 3])
        v, c, n = pp.ops_numba.points_to_voxel(g["points"].copy(), vs, rg, P, cap, True, perm=perm)
    else:           # shuffle variant: replay the order the reference left in the caller's array
        v, c, n = pp.ops_numba.points_to_voxel(g["points_after"].copy(), vs, rg, P, cap, False,
                                               perm=np.arange(len(g["points"])))
    assert v.dtype == np.float32 and c.dtype == np.int32 and n.dtype == np.int32
    assert np.array_equal(c, g["coors"])
    assert np.array_equal(n, g["num"])
    assert np.array_equal(v.view(np.uint32), g["voxels"].view(np.uint32))


def test_voxelize_ties_documented_rule(pp, oracle):
    """With ties the GPU order is (reflectance desc, original index asc): equals the oracle fed with
    that permutation."""
    g = golden("vox_model_ties")
    vs, rg, P, cap, _ = vox_args(g)
    pts = g["points"]
    v, c, n = pp.ops_numba.points_to_voxel(pts.copy(), vs, rg, P, cap, True)
    perm = np.argsort(-pts[:, 3], kind="stable")
    ov, oc, on = oracle.points_to_voxel(pts, vs, rg, P, cap, False, perm=perm)
    assert np.array_equal(c, oc) and np.array_equal(n, on) and np.array_equal(v, ov)


def test_voxelize_exact_ties_option(pp):
    """exact_ties=True: the reference's tie order (numba's own argsort on the host) replayed -> bit-exact with the
    REFERENCE's output on quantised reflectance, without the oracle."""
    pytest.importorskip("numba")
    g = golden("vox_model_ties")
    vs, rg, P, cap, _ = vox_args(g)
    v, c, n = pp.ops_numba.points_to_voxel(g["points"].copy(), vs, rg, P, cap, True, exact_ties=True)
    assert np.array_equal(c, g["coors"]) and np.array_equal(n, g["num"])
    assert np.array_equal(v.view(np.uint32), g["voxels"].view(np.uint32))


@pytest.mark.parametrize("kind,n,cap,P", [("dense", 1_000_000, 12000, 32), ("uniform", 1_000_000, 12000, 32),
                                           ("dense", 300_000, 3000, 32), ("forest", 120_000, 7500000, 50)])
def test_voxelize_full_size_vs_oracle(pp, oracle, kind, n, cap, P):
    from objectdetection_3d_b200 import synth
    if kind == "forest":
        g, pts = synth.G_REF, synth.forest_tile(n=n)
    else:
        g = synth.G_KITTI
        pts = synth.dense_tile(n=n) if kind == "dense" else synth.uniform_tile(n=n, margin=0.02)
    vs = np.array(g["voxel_size"], dtype=np.float32)
    rg = np.array(g["point_cloud_range"], dtype=np.float64)
    for refl in (True, False):
        if refl:
            v, c, m = pp.ops_numba.points_to_voxel(pts.copy(), vs, rg, P, cap, True)
        else:
            v, c, m = pp.ops_numba.points_to_voxel(pts.copy(), vs, rg, P, cap, False, perm=np.arange(n))
        ov, oc, om = oracle.points_to_voxel(pts, vs, rg, P, cap, refl)
        assert np.array_equal(c, oc) and np.array_equal(m, om) and np.array_equal(v, ov), (kind, refl)
    # size-independent properties: every kept row is an input point; counts bounded
    assert (m >= 1).all() and (m <= P).all()


@pytest.mark.parametrize("case", ["dense300k", "uniform300k_break", "dense50k_hash", "uniform60k_break_hash", "one_cell",
                                  "two_points_per_cell", "P5", "P50", "P100", "tiny"])
def test_voxelize_addressing_vs_oracle(pp, oracle, case):
    """Both ways a cell is addressed (slot = cell when the grid is no larger than the 2n-slot hash table would be, else
    the hash of the cell id: KITTI grid with n <= 65 536 points), every selection width (max_points <= 32, <= 64, any),
    a cell that holds every point, and the `break` at max_voxels."""
    from objectdetection_3d_b200 import synth
    g = synth.G_KITTI
    vs = np.array(g["voxel_size"], dtype=np.float32)
    rg = np.array(g["point_cloud_range"], dtype=np.float64)
    P, cap = 32, 12000
    rng = np.random.default_rng(11)
    if case == "dense300k":
        pts = synth.dense_tile(n=300_000, seed=21)
    elif case == "uniform300k_break":
        pts = synth.uniform_tile(n=300_000, margin=0.02)          # ~160k occupied cells >> cap: the `break`
    elif case == "dense50k_hash":
        pts = synth.dense_tile(n=50_000, seed=24, n_cells=2500, n_clusters=60)
    elif case == "uniform60k_break_hash":
        pts = synth.uniform_tile(n=60_000, margin=0.02)
        cap = 5000
    elif case == "one_cell":
        pts = np.empty((50_000, 4), np.float32)                  # every point in one pillar
        pts[:, 0] = 10.0 + 0.05 * rng.random(50_000)
        pts[:, 1] = 0.05 * rng.random(50_000)
        pts[:, 2] = -1.0
        pts[:, 3] = rng.permutation(50_000) / 50_000
    elif case == "two_points_per_cell":
        pts = synth.uniform_tile(n=20_000, margin=0.02)
        pts = np.concatenate([pts, pts + np.float32(1e-3)]).astype(np.float32)
        pts[:, 3] = rng.permutation(len(pts)) / len(pts)
        cap = 100_000
    elif case in ("P5", "P50", "P100"):
        pts = synth.dense_tile(n=100_000, seed=22, n_cells=900, n_clusters=40)
        P = int(case[1:])
    else:
        pts = synth.dense_tile(n=700, seed=23, n_cells=30, n_clusters=3)
    n = len(pts)
    for refl in (True, False):
        if refl:
            v, c, m = pp.ops_numba.points_to_voxel(pts.copy(), vs, rg, P, cap, True)
            ov, oc, om = oracle.points_to_voxel(pts, vs, rg, P, cap, True)
        else:
            perm = rng.permutation(n).astype(np.int32)
            v, c, m = pp.ops_numba.points_to_voxel(pts.copy(), vs, rg, P, cap, False, perm=perm)
            ov, oc, om = oracle.points_to_voxel(pts, vs, rg, P, cap, False, perm=perm)
        assert np.array_equal(c, oc) and np.array_equal(m, om) and np.array_equal(v, ov), (case, refl)


def test_voxelize_ties_full_size(pp, oracle):
    """D1M-ties at full size: reflectance quantised to 256 levels (as real LiDAR intensity is).  The reference's order
    among equal reflectances is numba's quicksort order: replayed through PP_ORDER_PERM it is bit-exact; the default
    path follows the documented rule (reflectance desc, index asc; -0.0 == +0.0)."""
    from objectdetection_3d_b200 import synth
    g = synth.G_KITTI
    vs = np.array(g["voxel_size"], dtype=np.float32)
    rg = np.array(g["point_cloud_range"], dtype=np.float64)
    pts = synth.dense_tile(n=1_000_000, seed=2025, ties=True)
    pts[::7919, 3] = -0.0                                          # a negative zero ties with the positive ones
    perm = oracle.numba_argsort_desc(pts[:, 3])
    v, c, m = pp.ops_numba.points_to_voxel(pts.copy(), vs, rg, 32, 12000, True, perm=perm)
    ov, oc, om = oracle.points_to_voxel(pts, vs, rg, 32, 12000, True)
    assert np.array_equal(c, oc) and np.array_equal(m, om) and np.array_equal(v, ov)
    v, c, m = pp.ops_numba.points_to_voxel(pts.copy(), vs, rg, 32, 12000, True)
    stable = np.argsort(-pts[:, 3], kind="stable")
    ov, oc, om = oracle.points_to_voxel(pts, vs, rg, 32, 12000, False, perm=stable)
    assert np.array_equal(c, oc) and np.array_equal(m, om) and np.array_equal(v.view(np.uint32), ov.view(np.uint32))


def test_voxelize_edge_cases(pp):
    from objectdetection_3d_b200 import synth
    g = synth.G_KITTI
    vs, rg = np.array(g["voxel_size"], dtype=np.float32), np.array(g["point_cloud_range"])
    v, c, n = pp.ops_numba.points_to_voxel(np.zeros((0, 4), np.float32), vs, rg, 32, 100, True)
    assert v.shape == (0, 32, 4) and c.shape == (0, 3) and n.shape == (0,)
    far = np.full((100, 4), 1e6, np.float32)
    v, c, n = pp.ops_numba.points_to_voxel(far, vs, rg, 32, 100, True)
    assert v.shape[0] == 0
    one = np.array([[1.0, 1.0, 0.0, 0.5]] * 100, np.float32)
    v, c, n = pp.ops_numba.points_to_voxel(one, vs, rg, 32, 100, True)
    assert v.shape[0] == 1 and n[0] == 32 and (v[0] == one[0]).all()


def test_module_voxelization_device(pp, oracle):
    from objectdetection_3d_b200 import synth
    g = synth.G_KITTI
    pts = synth.dense_tile(n=200_000, seed=5)
    mod = pp.pointpillars.PointPillarsVoxelization("cuda", g["voxel_size"], g["point_cloud_range"], 32, 12000)
    v, c, n = mod(pts)
    ov, oc, on = oracle.pointpillars_voxelization(pts, g["voxel_size"], g["point_cloud_range"], 32, 12000)
    assert c.dtype == torch.int64 and n.dtype == torch.int64 and v.is_cuda
    assert np.array_equal(v.cpu().numpy(), ov) and np.array_equal(c.cpu().numpy(), oc) and np.array_equal(n.cpu().numpy(), on)


# ------------------------------------------------------------------------------------------ PFN / scatter
@pytest.mark.parametrize("name", ["pfn_single64", "pfn_two_layer"])
def test_pfn_golden_and_oracle(pp, oracle, name):
    g = golden(name)
    vs, rg = g["voxel_size"].tolist(), g["point_cloud_range"].tolist()
    nl = int(g["n_layers"])
    feat = [g["w%d" % i].shape[0] * (1 if i == nl - 1 else 2) + (1 if i == nl - 1 else 0) for i in range(nl)]
    net = pp.pointpillars.PillarFeatureNet(4, feat, vs, rg).cuda().eval()
    for i, l in enumerate(net.pfn_layers):
        with torch.no_grad():
            l.linear.weight.copy_(cu(g["w%d" % i])); l.norm.weight.copy_(cu(g["gamma%d" % i]))
            l.norm.bias.copy_(cu(g["beta%d" % i])); l.norm.running_mean.copy_(cu(g["mean%d" % i]))
            l.norm.running_var.copy_(cu(g["var%d" % i]))
    voxels, num, coors = cu(g["voxels"]), cu(g["num"]), cu(g["coors"])
    scale = float(np.abs(g["voxels"]).max())
    dec = net.decorate(voxels, num, coors).cpu().numpy()
    assert_close_t1(dec, g["decorated"], atol=1e-5 * scale, what="decorated vs reference")
    odec = oracle.decorate(g["voxels"], g["num"], g["coors"], vs[0], vs[1], vs[0] / 2 + rg[0], vs[1] / 2 + rg[1])
    assert np.array_equal(dec.view(np.uint32), odec.view(np.uint32)), "decoration is bit-exact vs the oracle"
    with torch.no_grad():
        out = net(voxels, num, coors).cpu().numpy()
    assert_close_t1(out, g["out"], atol=1e-5 * scale, what="pfn vs reference")
    layers = [dict(weight=g["w%d" % i], gamma=g["gamma%d" % i], beta=g["beta%d" % i], mean=g["mean%d" % i],
                   var=g["var%d" % i]) for i in range(nl)]
    oout = oracle.pillar_feature_net(g["voxels"], g["num"], g["coors"], layers, vs, rg)
    assert_close_t1(out, oout, atol=1e-6 * scale, what="pfn vs oracle")
    # int32 inputs (what the voxelizer produces natively) give the same result
    with torch.no_grad():
        out32 = net(voxels, num.int(), coors.int()).cpu().numpy()
    assert np.array_equal(out32, out)
    # scatter
    H, W = g["canvas_hw"].tolist()
    sc = pp.pointpillars.SparseMiddleExtractor([1, H, W])
    canvas = sc(cu(g["out"]), coors, 1).cpu().numpy()
    assert np.array_equal(canvas, g["canvas"])
    canvas = sc(cu(g["out"]), coors.int(), 1).cpu().numpy()
    assert np.array_equal(canvas, g["canvas"])


def test_pfn_eval_mode_keeps_gradients(pp):
    """eval() with autograd recording (frozen-BatchNorm fine-tuning, saliency): the forward-only kernels must not be
    taken -- the output has a grad_fn, matches the kernel path and the weights / points receive gradients."""
    g = golden("pfn_single64")
    vs, rg = g["voxel_size"].tolist(), g["point_cloud_range"].tolist()
    net = pp.pointpillars.PillarFeatureNet(4, [64], vs, rg).cuda().eval()
    l = net.pfn_layers[0]
    with torch.no_grad():
        l.linear.weight.copy_(cu(g["w0"])); l.norm.weight.copy_(cu(g["gamma0"])); l.norm.bias.copy_(cu(g["beta0"]))
        l.norm.running_mean.copy_(cu(g["mean0"])); l.norm.running_var.copy_(cu(g["var0"]))
    voxels, num, coors = cu(g["voxels"]), cu(g["num"]), cu(g["coors"])
    with torch.no_grad():
        ref = net(voxels, num, coors)
    out = net(voxels, num, coors)
    assert out.grad_fn is not None
    scale = float(np.abs(g["voxels"]).max())
    assert_close_t1(out.detach().cpu().numpy(), ref.cpu().numpy(), atol=1e-6 * scale, what="autograd vs kernel path")
    out.sum().backward()
    assert l.linear.weight.grad is not None and float(l.linear.weight.grad.abs().sum()) > 0
    v2 = voxels.clone().requires_grad_(True)
    net(v2, num, coors).sum().backward()
    assert v2.grad is not None and float(v2.grad.abs().sum()) > 0
    for p_ in net.parameters():
        p_.requires_grad_(False)
    assert net(voxels, num, coors).grad_fn is None            # nothing to differentiate: the kernel path again


def test_scatter_3d_batched_and_backward(pp, oracle):
    rng = np.random.default_rng(3)
    B, D, H, W, C, M = 3, 4, 37, 41, 5, 900          # H*W not a multiple of 4 -> scalar store path
    lin = rng.choice(B * D * H * W, size=M, replace=False)
    coors = np.stack(np.unravel_index(lin, (B, D, H, W)), 1).astype(np.int32)
    feat = rng.normal(size=(M, C)).astype(np.float32)
    sc = pp.pointpillars.SparseMiddleExtractor([D, H, W])
    f = cu(feat).requires_grad_(True)
    canvas = sc(f, cu(coors), B)
    assert np.array_equal(canvas.detach().cpu().numpy(), oracle.scatter_dense(feat, coors, B, D, H, W))
    w = torch.randn_like(canvas)
    (canvas * w).sum().backward()
    c = coors.astype(np.int64)
    expect = w.view(B, C, D, H, W).cpu().numpy()[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]]
    assert np.array_equal(f.grad.cpu().numpy(), expect)


# ------------------------------------------------------------------------------------------ boxes
def test_boxes_iou_golden(pp, oracle):
    g = golden("boxes_iou")
    b = cu(g["boxes"])
    rect = pp.ops_torch.bbox2rotated_corners2D(b).cpu().numpy()
    assert_close_t1(rect, g["rect"], what="aabb")
    assert_close_t1(pp.ops_torch.bbox2corners3D(b).cpu().numpy(), g["corners"], atol=2e-6, what="corners")
    r = cu(g["rect"])
    assert np.array_equal(pp.ops_torch.bbox_iou2D(r[:200], r[200:500]).cpu().numpy(), g["iou"])
    assert np.array_equal(pp.ops_torch.bbox_iou2D(r[:50], r[200:300], "iof").cpu().numpy(), g["iof"])
    assert np.array_equal(pp.ops_torch.bbox_iou2D(r[:50], r[200:300], "giou").cpu().numpy(), g["giou"])
    assert np.array_equal(pp.ops_numba.iou_jit(g["rect"][:40], g["rect"][300:360], 0.0), g["iou_jit"])
    assert np.array_equal(pp.ops_numba.iou_jit(g["rect"][:40], g["rect"][300:360], 1.0), g["iou_jit_eps1"])
    with pytest.raises(AssertionError):
        pp.ops_torch.bbox_iou2D(r[:5], r[:5], "bad")
    assert pp.ops_torch.bbox_iou2D(r[:0], r[:5]).shape == (0, 5)


def test_codec_anchors_golden(pp, oracle):
    from objectdetection_3d_b200 import synth
    g = golden("codec")
    enc = pp.model_utils.BBoxCoder.encode(cu(g["anchors"]), cu(g["gts"])).cpu().numpy()
    dec = pp.model_utils.BBoxCoder.decode(cu(g["anchors"]), cu(g["deltas"])).cpu().numpy()
    assert_close_t1(enc, g["encoded"], what="encode")
    assert_close_t1(dec, g["decoded"], what="decode")
    assert_close_t1(pp.model_utils.limit_period(cu(g["val"]), 1, np.pi).cpu().numpy(), g["limit_1_pi"], atol=2e-6)
    assert_close_t1(pp.model_utils.limit_period(cu(g["val"]), 0.5, 2 * np.pi).cpu().numpy(), g["limit_05_2pi"], atol=2e-6)
    a = golden("anchors")
    gen = pp.model_utils.Anchor3DRangeGenerator([[0, 0, 0, 40.0, 40.0, 30.0]], synth.ANCHOR_SIZES,
                                                synth.ANCHOR_ROTATIONS, 9)
    got = gen.grid_anchors((5, 7), device="cuda").cpu().numpy()
    assert got.shape == a["a57"].shape
    assert_close_t1(got, a["a57"], what="anchors")
    gen2 = pp.model_utils.Anchor3DRangeGenerator([[0, -39.68, -1.78, 69.12, 39.68, -1.78]], [[1.6, 3.9, 1.56]],
                                                 [[0, 0, 0], [0, 0, 1.57]], 9)
    assert_close_t1(gen2.grid_anchors((31, 27), device="cuda").cpu().numpy(), a["a_kitti"], what="anchors kitti")


# ------------------------------------------------------------------------------------------ NMS
def test_multiclass_nms_golden(pp):
    g = golden("nms_multiclass")
    b, s = cu(g["boxes"]), cu(g["scores"])
    for si, sthr in enumerate(g["score_thrs"].tolist()):
        for ii, ithr in enumerate(g["iou_thrs"].tolist()):
            keep = pp.model_utils.multiclass_nms(b, s, sthr, ithr, 2)
            for c in range(2):
                assert keep[c].dtype == torch.int64
                assert np.array_equal(keep[c].cpu().numpy(), g["keep_s%d_i%d_c%d" % (si, ii, c)]), (sthr, ithr, c)


@pytest.mark.parametrize("n,extent", [(20_000, 40.0), (20_000, 200.0), (5000, 10.0), (65, 3.0), (1, 1.0)])
def test_nms_full_size_vs_oracle(pp, oracle, n, extent):
    from objectdetection_3d_b200 import synth
    boxes, scores = synth.nms_boxes(n=n, seed=4, extent=extent)
    b, s = cu(boxes), cu(scores)
    for sthr, ithr in ((0.05, 1e-5), (0.3, 0.1), (0.5, 0.5), (0.99999, 0.1)):
        got = pp.model_utils.multiclass_nms(b, s, sthr, ithr, 2)[0].cpu().numpy()
        # the oracle NMS is fed the rectangles the GPU computed, so the keep list must be IDENTICAL
        # (T0 "given identical rectangles"); rectangles themselves are T1-checked above
        cand = np.nonzero(scores[:, 0] > np.float32(sthr))[0]
        order = cand[np.argsort(-scores[cand, 0], kind="stable")]
        rect = pp.ops_torch.bbox2rotated_corners2D(b).cpu().numpy()
        keep = oracle.nms_sorted(rect[order], ithr)
        assert np.array_equal(got, order[keep]), (n, extent, sthr, ithr)
        # properties: descending score; no kept pair overlaps above thr
        assert (np.diff(scores[got, 0]) < 0).all()
        if 0 < len(got) <= 3000:
            iou = oracle.bbox_iou2D(rect[got], rect[got])
            np.fill_diagonal(iou, 0)
            assert not (iou > np.float32(ithr)).any()


def test_nms_empty_and_maximum_size(pp, oracle):
    """N = 0, and the largest N the library accepts (131 072), checked through size-independent properties: descending
    scores, no kept pair above the threshold (sampled), every dropped candidate overlaps a better kept box (sampled)."""
    from objectdetection_3d_b200 import synth
    empty = pp.model_utils.multiclass_nms(torch.zeros((0, 9), device="cuda"), torch.zeros((0, 2), device="cuda"), 0.1, 0.1, 2)
    assert len(empty) == 2 and all(k.numel() == 0 and k.dtype == torch.int64 for k in empty)
    n = 131_072
    boxes, scores = synth.nms_boxes(n=n, seed=77, extent=300.0)
    b, s = cu(boxes), cu(scores)
    keep = pp.model_utils.multiclass_nms(b, s, 0.0, 0.1, 2)[0].cpu().numpy()
    assert len(np.unique(keep)) == len(keep) and (np.diff(scores[keep, 0]) < 0).all() and len(keep) > 5_000
    rect = pp.ops_torch.bbox2rotated_corners2D(b)
    rng = np.random.default_rng(1)
    sub = np.sort(rng.choice(len(keep), 3000, replace=False))
    kk = pp.ops_torch.bbox_iou2D(rect[keep[sub]], rect[keep]).cpu().numpy()
    kk[np.arange(len(sub)), sub] = 0
    assert not (kk > np.float32(0.1)).any()
    dropped = np.setdiff1d(np.arange(n), keep)
    dsub = rng.choice(dropped, 2000, replace=False)
    dk = pp.ops_torch.bbox_iou2D(rect[dsub], rect[keep]).cpu().numpy()
    better = scores[keep, 0][None, :] > scores[dsub, 0][:, None]
    assert ((dk > np.float32(0.1)) & better).any(axis=1).all()
    with pytest.raises(ValueError):
        pp.model_utils.multiclass_nms(torch.zeros((131_073, 9), device="cuda"), torch.zeros((131_073, 1), device="cuda"), 0.1, 0.1, 2)


def test_head_golden(pp):
    """Anchor3DHead.get_bboxes_single / assign_bboxes against the reference's outputs."""
    from objectdetection_3d_b200 import synth
    g = golden("head")
    head = pp.pointpillars.Anchor3DHead(num_classes=1, in_channels=8, nms_dim=2, nms_pre=300, nms_thresh=0.1,
                                        score_thr=0.3, ranges=[[0, 0, 0, 40.0, 40.0, 30.0]], sizes=synth.ANCHOR_SIZES,
                                        rotations=synth.ANCHOR_ROTATIONS, iou_thr=[[0.08, 0.2]]).cuda()
    with torch.no_grad():
        b, s, l = head.get_bboxes_single(cu(g["cls"]), cu(g["reg"]), cu(g["dirs"]))
    assert b.shape == g["bboxes"].shape
    assert_close_t1(s.cpu().numpy(), g["scores"], what="scores")
    assert_close_t1(b.cpu().numpy(), g["bboxes"], atol=1e-5, what="bboxes")
    assert np.array_equal(l.numpy(), g["labels"])
    with torch.no_grad():           # the fused path against the reference's order of operations on the same GPU
        b2, s2, l2 = head.get_bboxes_single_unfused(cu(g["cls"]), cu(g["reg"]), cu(g["dirs"]))
    assert torch.equal(l, l2) and b.shape == b2.shape
    assert_close_t1(s.cpu().numpy(), s2.cpu().numpy(), atol=1e-6, what="fused vs unfused scores")
    assert_close_t1(b.cpu().numpy(), b2.cpu().numpy(), atol=1e-5, what="fused vs unfused boxes")
    head.nms_pre = 10 ** 9          # no top-k: every anchor goes through select/decode (rows == NULL)
    with torch.no_grad():
        b3, s3, l3 = head.get_bboxes_single(cu(g["cls"]), cu(g["reg"]), cu(g["dirs"]))
        b4, s4, l4 = head.get_bboxes_single_unfused(cu(g["cls"]), cu(g["reg"]), cu(g["dirs"]))
    assert torch.equal(l3, l4) and b3.shape == b4.shape and len(b3) >= len(b)
    assert_close_t1(b3.cpu().numpy(), b4.cpu().numpy(), atol=1e-5, what="fused vs unfused boxes (no top-k)")
    head.nms_pre = 300
    with torch.no_grad():
        ab, ti, pi, ni = head.assign_bboxes(cu(g["reg"]).unsqueeze(0), [cu(g["gts"])])
    assert np.array_equal(ti.cpu().numpy(), g["target_idx"])
    assert np.array_equal(pi.cpu().numpy(), g["pos_idx"])
    assert np.array_equal(ni.cpu().numpy(), g["neg_idx"])
    assert_close_t1(ab.cpu().numpy(), g["assigned"], atol=1e-5, what="assigned")


def test_assign_overlaps_fused(pp, oracle):
    """pp_assign_overlaps (IoU + row/column maxima + low-quality flags, no (G, A) matrix) against the reductions of
    model/PointPillars.py:968-978 applied to the oracle's IoU matrix; rectangles: bit-exact."""
    from objectdetection_3d_b200 import synth
    gen = pp.model_utils.Anchor3DRangeGenerator([[0, 0, 0, 40.0, 40.0, 30.0]], synth.ANCHOR_SIZES, synth.ANCHOR_ROTATIONS, 9)
    anchors = gen.grid_anchors((60, 70), device="cuda").reshape(-1, 9)
    gts, _ = synth.nms_boxes(n=37, seed=21, extent=38.0, tilt=0.2)
    gts[3] = anchors[1234].cpu().numpy()                       # an exact match (IoU 1 on several co-located anchors)
    gts[4, :2] = 500.0                                         # a ground truth that overlaps nothing
    tg = cu(gts)
    ar, gr = pp.ops_torch.bbox2rotated_corners2D(anchors), pp.ops_torch.bbox2rotated_corners2D(tg)
    for lo in (0.08, 0.45):
        mo, am, gm, lq = pp.model_utils.assign_overlaps(gr, ar, lo, 2)
        omo, oam, ogm, olq = oracle.assign_overlaps(oracle.bbox_iou2D(gr.cpu().numpy(), ar.cpu().numpy()), lo)
        assert np.array_equal(mo.cpu().numpy(), omo) and np.array_equal(gm.cpu().numpy(), ogm)
        assert np.array_equal(am.cpu().numpy(), oam) and np.array_equal(lq.cpu().numpy(), olq)
        assert olq.sum() >= 10 and gm[4].item() == 0.0 and not lq[(mo == 0)].any()
    # 3-D form against the float64 oracle (T1 on the maxima; flags checked through the kernel's own IoU matrix)
    sub = anchors[::7].contiguous()
    ac, gc = pp.ops_torch.bbox2corners3D(sub), pp.ops_torch.bbox2corners3D(tg)
    mo, am, gm, lq = pp.model_utils.assign_overlaps(gc, ac, 0.08, 3)
    m3 = pp.ops_torch.box3d_overlap(gc, ac).cpu().numpy()
    emo, eam, egm, elq = oracle.assign_overlaps(m3, 0.08)
    assert np.array_equal(mo.cpu().numpy(), emo) and np.array_equal(gm.cpu().numpy(), egm)
    assert np.array_equal(am.cpu().numpy(), eam) and np.array_equal(lq.cpu().numpy(), elq)
    _, o3 = oracle.box3d_overlap(gc.cpu().numpy(), ac.cpu().numpy())
    assert np.abs(mo.cpu().numpy() - o3.max(axis=0)).max() < 5e-5


def test_head_box3d_assign_runs(pp):
    """nms_dim == 3 (what config.yaml:6 selects): assign_bboxes / get_bboxes_single run on the BOX3D kernels and agree
    with the nms_dim == 2 head where the two IoU definitions coincide (every ground truth equal to an anchor)."""
    from objectdetection_3d_b200 import synth
    kw = dict(num_classes=1, in_channels=8, nms_pre=300, nms_thresh=0.1, score_thr=0.3,
              ranges=[[0, 0, 0, 40.0, 40.0, 30.0]], sizes=synth.ANCHOR_SIZES, rotations=synth.ANCHOR_ROTATIONS,
              iou_thr=[[0.3, 0.6]])
    h3 = pp.pointpillars.Anchor3DHead(nms_dim=3, **kw).cuda()
    g = golden("head")
    reg = cu(g["reg"]).unsqueeze(0)
    anchors = h3.anchor_generator.grid_anchors(reg.shape[-2:], device="cuda").reshape(-1, 9)
    na = anchors.shape[0]
    pick = torch.tensor([na // 7, na // 2, na - 5], device="cuda")
    with torch.no_grad():
        ab, ti, pi, ni = h3.assign_bboxes(reg, [anchors[pick].clone()])
        b, s, l = h3.get_bboxes_single(cu(g["cls"]), cu(g["reg"]), cu(g["dirs"]))
    assert set(pick.tolist()) <= set(pi.tolist()) or len(pi) >= 3          # the matched anchors are positives
    assert ab.shape[1] == 9 and len(ti) == len(pi) and len(ni) > 0 and b.shape[1] == 9 and len(s) == len(l) == len(b)
    assert torch.isfinite(ab).all()


def test_extract_feats_sequence_batch3(pp, oracle):
    """PointPillars.extract_feats up to the pseudo-image (model/PointPillars.py:94-134) restated on the mirrors:
    three numpy frames -> voxel_layer each -> concat, batch index padded in front of the coordinates ->
    voxel_encoder -> batch_size = coors[-1, 0] + 1 -> pseudoimage_generator, against the oracle frame by frame."""
    import torch.nn.functional as F
    from objectdetection_3d_b200 import synth
    g, pfn = synth.G_KITTI, synth.pfn_params(9, 63, seed=5)
    frames = [synth.dense_tile(n=80_000, seed=70 + i, n_cells=2000 + 500 * i, n_clusters=40) for i in range(3)]
    voxel_layer = pp.pointpillars.PointPillarsVoxelization("cuda", g["voxel_size"], g["point_cloud_range"], 32, 12000)
    voxel_encoder = pp.pointpillars.PillarFeatureNet(4, [64], g["voxel_size"], g["point_cloud_range"]).cuda().eval()
    l = voxel_encoder.pfn_layers[0]
    with torch.no_grad():
        l.linear.weight.copy_(cu(pfn["weight"])); l.norm.weight.copy_(cu(pfn["gamma"])); l.norm.bias.copy_(cu(pfn["beta"]))
        l.norm.running_mean.copy_(cu(pfn["mean"])); l.norm.running_var.copy_(cu(pfn["var"]))
    pseudoimage_generator = pp.pointpillars.SparseMiddleExtractor([1, 496, 432])
    with torch.no_grad():
        voxels, coors, num_points = [], [], []
        for pc in frames:                                                    # :114-121
            v, c, n = voxel_layer(pc)
            voxels.append(v); coors.append(c); num_points.append(n)
        voxels, num_points = torch.cat(voxels, dim=0), torch.cat(num_points, dim=0)
        coors = torch.cat([F.pad(c, (1, 0), mode="constant", value=i) for i, c in enumerate(coors)], dim=0)   # :129-132
        feats = voxel_encoder(voxels, num_points, coors)                     # :98
        batch_size = coors[-1, 0].item() + 1                                 # :99
        x = pseudoimage_generator(feats, coors, batch_size)                  # :100
    assert batch_size == 3 and tuple(x.shape) == (3, 64, 496, 432) and coors.dtype == torch.int64
    at = 0
    scale = float(max(np.abs(f[:, :3]).max() for f in frames))
    for i, pc in enumerate(frames):
        ov, oc, on = oracle.pointpillars_voxelization(pc, g["voxel_size"], g["point_cloud_range"], 32, 12000)
        m = len(ov)
        assert np.array_equal(voxels[at:at + m].cpu().numpy(), ov) and np.array_equal(num_points[at:at + m].cpu().numpy(), on)
        assert np.array_equal(coors[at:at + m, 1:].cpu().numpy(), oc) and (coors[at:at + m, 0] == i).all()
        c4 = np.concatenate([np.zeros((m, 1), np.int64), oc], 1)
        of = oracle.pillar_feature_net(ov, on, c4, [pfn], g["voxel_size"], g["point_cloud_range"])
        assert_close_t1(feats[at:at + m].cpu().numpy(), of, atol=1e-6 * scale, what="features of frame %d" % i)
        ref = oracle.scatter_dense(of, c4.astype(np.int32), 1, 1, 496, 432)[0]
        assert_close_t1(x[i].cpu().numpy(), ref, atol=1e-6 * scale, what="pseudo-image of frame %d" % i)
        at += m
    assert at == len(voxels)


@pytest.mark.parametrize("order", ["reflectance", "given"])
def test_frame_pipeline_full_size(pp, oracle, order):
    """BASELINE configs[1] through the preallocated pipeline vs the oracle: the one-call frame (pp_voxelize_scatter), the
    gather+PFN kernel with the stand-alone canvas kernel, and the three stand-alone calls give bit-identical frames."""
    from objectdetection_3d_b200 import _lib, pipeline, synth
    g, pfn = synth.G_KITTI, synth.pfn_params(9, 63, seed=5)
    pts = synth.dense_tile()
    pipe = pipeline.FramePipeline(g, pfn, len(pts),
                                  order=_lib.ORDER_REFLECTANCE_DESC if order == "reflectance" else _lib.ORDER_GIVEN)
    assert pipe.fused
    canvas = pipe.new_canvas()
    canvas.fill_(7.0)           # every element of the canvas must be written by the call
    for _ in range(2):          # twice: the workspace must be reusable without re-initialisation by the caller
        pipe.run(cu(pts), canvas)
    torch.cuda.synchronize()
    m = int(pipe.voxel_num.item())
    ov, oc, on = oracle.points_to_voxel(pts, np.array(g["voxel_size"], np.float32), np.array(g["point_cloud_range"]),
                                        32, 12000, order == "reflectance")
    assert m == len(ov)
    assert np.array_equal(pipe.voxels[:m].cpu().numpy(), ov)
    assert np.array_equal(pipe.coors[:m].cpu().numpy(), oc) and np.array_equal(pipe.num[:m].cpu().numpy(), on)
    coors4 = np.concatenate([np.zeros((m, 1), np.int64), oc[:, [2, 1, 0]].astype(np.int64)], 1)
    feat = oracle.pillar_feature_net(ov, on, coors4, [pfn], g["voxel_size"], g["point_cloud_range"])
    ref = oracle.scatter_dense(feat, coors4.astype(np.int32), 1, 1, pipe.H, pipe.W)
    got = canvas.cpu().numpy()
    scale = float(np.abs(pts[:, :3]).max())
    assert_close_t1(pipe.feat[:m].cpu().numpy(), feat, atol=1e-6 * scale, what="features")
    assert_close_t1(got, ref, atol=1e-6 * scale, what="canvas")
    feat_fused, vox_fused = pipe.feat[:m].clone(), pipe.voxels[:m].clone()
    for mode in ("features", False):
        canvas2 = pipe.new_canvas()
        canvas2.fill_(-3.0)
        pipe.run(cu(pts), canvas2, fused=mode)
        torch.cuda.synchronize()
        assert torch.equal(pipe.feat[:m], feat_fused) and torch.equal(pipe.voxels[:m], vox_fused), mode
        assert torch.equal(canvas2, canvas), mode
        pm = pipe.pillar_map.cpu().numpy()
        assert (pm >= 0).sum() == m and np.array_equal(pm[oc[:, 2], oc[:, 1], oc[:, 0]], np.arange(m))
    assert np.array_equal(got == 0, ref == 0) or np.abs(got[(got == 0) != (ref == 0)]).max() < 1e-4


@pytest.mark.parametrize("n,k", [(1_920_000, 500), (1_920_000, 4096), (100_000, 20_000), (5000, 5000), (777, 10), (64, 100)])
def test_head_topk_vs_stable_argsort(pp, n, k):
    """pp_head_topk = top-nms_pre of the per-anchor scores (model/PointPillars.py:1056-1065) at the reference's
    400 x 400 x 12 anchor count: indices in descending score, ties by lower index == np.argsort(-s, kind="stable")[:k].
    Scores are quantised so that ties straddle the k-th place."""
    rng = np.random.default_rng(n + k)
    s = (rng.integers(0, 3000, size=n) / 3000.0).astype(np.float32)
    s[rng.integers(0, n, size=5)] = 1.0
    got = pp.pointpillars.head_topk(cu(s), k).cpu().numpy()
    ref = np.argsort(-s, kind="stable")[:min(k, n)]
    assert got.dtype == np.int64 and np.array_equal(got, ref)
    # tie-free scores: same indices in the same order as torch.topk
    s2 = (rng.permutation(n) / float(n)).astype(np.float32)           # tie-free: distinct in float32 for n <= 2^23
    got2 = pp.pointpillars.head_topk(cu(s2), k)
    assert torch.equal(got2, cu(s2).topk(min(k, n))[1])

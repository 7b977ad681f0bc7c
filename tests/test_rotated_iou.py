"""Rotated BEV IoU (extension): the float64 oracle against an independent half-space intersection (scipy) and
against identities; the CUDA kernel against the oracle (gpu)."""
import numpy as np
import pytest


def _rect_halfspaces(b):
    x, y, dx, dy, rz = float(b[0]), float(b[1]), float(b[3]), float(b[4]), float(b[8])
    c, s = np.cos(rz), np.sin(rz)
    hs = []
    for nx, ny, h in ((c, s, dx / 2), (-c, -s, dx / 2), (-s, c, dy / 2), (s, -c, dy / 2)):
        hs.append([nx, ny, -(nx * x + ny * y) - h])          # n.p - (n.c + h) <= 0
    return np.array(hs)


def _independent_iou(a, b):
    from scipy.optimize import linprog
    from scipy.spatial import ConvexHull, HalfspaceIntersection
    hs = np.vstack([_rect_halfspaces(a), _rect_halfspaces(b)])
    # Chebyshev centre as the interior point
    norm = np.linalg.norm(hs[:, :2], axis=1)
    res = linprog([0, 0, -1], A_ub=np.hstack([hs[:, :2], norm[:, None]]), b_ub=-hs[:, 2], bounds=[(None, None)] * 2 + [(0, None)])
    inter = 0.0
    if res.status == 0 and res.x[2] > 1e-9:
        pts = HalfspaceIntersection(hs, res.x[:2]).intersections
        inter = ConvexHull(pts).volume
    return inter / (a[3] * a[4] + b[3] * b[4] - inter)


def _boxes(n, seed, extent=10.0):
    from objectdetection_3d_b200 import synth
    b, _ = synth.nms_boxes(n=n, seed=seed, extent=extent, tilt=0.0)
    return b


def test_oracle_rotated_iou_vs_independent(oracle):
    a, b = _boxes(25, 1), _boxes(30, 2)
    got = oracle.bbox_iou_rotated_bev(a, b)
    ref = np.array([[_independent_iou(x.astype(np.float64), y.astype(np.float64)) for y in b] for x in a])
    assert np.abs(got - ref).max() < 1e-9
    assert (got > 0.01).sum() > 20            # the sample does contain overlapping pairs


def test_oracle_rotated_iou_identities(oracle):
    a = _boxes(60, 3)
    m = oracle.bbox_iou_rotated_bev(a, a)
    assert np.allclose(np.diag(m), 1.0, atol=1e-12) and np.allclose(m, m.T, atol=1e-12)
    # yaw 0: equals the axis-aligned IoU of the footprints
    z = a.copy(); z[:, 8] = 0
    rect = np.stack([z[:, 0] - z[:, 3] / 2, z[:, 1] - z[:, 4] / 2, z[:, 0] + z[:, 3] / 2, z[:, 1] + z[:, 4] / 2], 1)
    assert np.abs(oracle.bbox_iou_rotated_bev(z, z) - oracle.bbox_iou2D(rect, rect).astype(np.float64)).max() < 1e-5
    # a rotation by pi/2 of a square changes nothing; by pi never does
    s = a.copy(); s[:, 4] = s[:, 3]
    t = s.copy(); t[:, 8] += np.pi / 2
    assert np.abs(oracle.bbox_iou_rotated_bev(s, s) - oracle.bbox_iou_rotated_bev(s, t)).max() < 1e-6
    u = a.copy(); u[:, 8] += np.pi
    assert np.abs(m - oracle.bbox_iou_rotated_bev(a, u)).max() < 1e-6


@pytest.mark.gpu
def test_cuda_rotated_iou_and_nms(oracle):
    import torch
    from objectdetection_3d_b200 import model_utils, ops_torch, synth
    a, b = _boxes(700, 5, 20.0), _boxes(900, 6, 20.0)
    got = ops_torch.bbox_iou_rotated_bev(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    ref = oracle.bbox_iou_rotated_bev(a, b)
    assert np.abs(got - ref).max() < 2e-6          # T1: float64 evaluation (fp32 sin, cos and output) vs the float64 oracle
    assert ops_torch.bbox_iou_rotated_bev(torch.from_numpy(a[:0]).cuda(), torch.from_numpy(b).cuda()).shape == (0, 900)
    # rotated NMS: identical to a CPU greedy loop driven by the GPU's own IoU matrix
    for n, extent in ((3000, 25.0), (20000, 40.0)):
        boxes, scores = synth.nms_boxes(n=n, seed=7, extent=extent, tilt=0.0)
        tb, ts = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
        for sthr, ithr in ((0.3, 0.1), (0.05, 0.5)):
            keep = model_utils.multiclass_nms(tb, ts, sthr, ithr, 2, iou_mode="rot_bev")[0].cpu().numpy()
            cand = np.nonzero(scores[:, 0] > np.float32(sthr))[0]
            order = cand[np.argsort(-scores[cand, 0], kind="stable")]
            assert (np.diff(scores[keep, 0]) < 0).all() and set(keep.tolist()) <= set(order.tolist())
            kept_boxes = torch.from_numpy(boxes[keep]).cuda()
            # (1) no kept pair overlaps above thr; (2) every dropped candidate overlaps a better kept box
            kk = ops_torch.bbox_iou_rotated_bev(kept_boxes, kept_boxes).cpu().numpy()
            np.fill_diagonal(kk, 0)
            assert not (kk > np.float32(ithr)).any()
            dropped = np.setdiff1d(order, keep)
            if len(dropped):
                dk = ops_torch.bbox_iou_rotated_bev(torch.from_numpy(boxes[dropped]).cuda(), kept_boxes).cpu().numpy()
                better = scores[keep, 0][None, :] > scores[dropped, 0][:, None]
                assert ((dk > np.float32(ithr)) & better).any(axis=1).all()


def _greedy_vs_float64(boxes_sorted, pair_iou, thr, gpu_pair_iou, band=1e-4, pad=1e-3):
    """Greedy NMS on score-sorted boxes, decisions from a float64 pair IoU evaluated only for pairs whose (padded) xy
    bounding squares overlap.  With ~10^6 decisive pairs at 20k boxes some IoU always lies within fp32 error of any
    threshold, so the few pairs within `band` of it are decided by the GPU's own pairwise IoU kernel (fp32, `>`), every
    other pair by the oracle.  Returns (kept ranks, number of pairs decided by the oracle, by the GPU)."""
    n = len(boxes_sorted)
    r = np.sqrt((boxes_sorted[:, 3:6].astype(np.float64) ** 2).sum(1)) + pad          # bounds the box from its bottom centre
    x, y = boxes_sorted[:, 0].astype(np.float64), boxes_sorted[:, 1].astype(np.float64)
    alive = np.ones(n, bool)
    kept, n_clear, n_amb = [], 0, 0
    for i in range(n):
        if not alive[i]:
            continue
        kept.append(i)
        later = np.nonzero(alive[i + 1:])[0] + i + 1
        near = later[(np.abs(x[later] - x[i]) < r[later] + r[i]) & (np.abs(y[later] - y[i]) < r[later] + r[i])]
        if len(near):
            iou = pair_iou(boxes_sorted[near], boxes_sorted[i:i + 1])[:, 0]
            hit = iou > thr
            amb = np.abs(iou - thr) <= band
            if amb.any():
                hit[amb] = gpu_pair_iou(boxes_sorted[near[amb]], boxes_sorted[i:i + 1])[:, 0] > np.float32(thr)
            n_clear += int((~amb).sum())
            n_amb += int(amb.sum())
            alive[near[hit]] = False
    return np.array(kept), n_clear, n_amb


@pytest.mark.gpu
@pytest.mark.parametrize("mode,extent,thr", [("rot_bev", 40.0, 0.1), ("rot_bev", 40.0, 0.5), ("rot_bev", 200.0, 0.1),
                                             ("box3d", 200.0, 0.1), ("box3d", 40.0, 0.1)])
def test_nms_20k_clipped_modes_vs_float64_greedy(oracle, mode, extent, thr):
    """NMS20k of SURVEY.md 8(d) with the clipped pair tests against a sequential greedy loop driven by the float64
    oracle IoU (not by the GPU's own IoU matrix): keep lists identical.  Only the pairs whose IoU lies within 1e-4 of
    the threshold (a few hundred of ~10^6) take the GPU's fp32 value, which the other IoU tests bound to 5e-5."""
    import torch
    from objectdetection_3d_b200 import model_utils, ops_torch, synth
    boxes, scores = synth.nms_boxes(n=20_000, seed=13, extent=extent, tilt=0.0 if mode == "rot_bev" else 0.3)
    order = np.argsort(-scores[:, 0], kind="stable")
    bs = boxes[order]
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if mode == "rot_bev":
        pair = oracle.bbox_iou_rotated_bev
        gpu_pair = lambda a, b: ops_torch.bbox_iou_rotated_bev(cu(a), cu(b)).cpu().numpy()
    else:
        pair = lambda a, b: oracle.box3d_overlap(oracle.bbox2corners3D(a), oracle.bbox2corners3D(b))[1]
        gpu_pair = lambda a, b: ops_torch.box3d_overlap(ops_torch.bbox2corners3D(cu(a)), ops_torch.bbox2corners3D(cu(b))).cpu().numpy()
    kept, n_clear, n_amb = _greedy_vs_float64(bs, pair, thr, gpu_pair)
    assert n_clear > 1000 * max(n_amb, 1) or n_amb < 2000, (n_clear, n_amb)
    kw = dict(iou_mode="rot_bev") if mode == "rot_bev" else {}
    got = model_utils.multiclass_nms(cu(boxes), cu(scores), 0.0, thr, 2 if mode == "rot_bev" else 3, **kw)[0].cpu().numpy()
    assert np.array_equal(got, order[kept]), (mode, extent, thr, len(got), len(kept), n_clear, n_amb)

"""Oriented 3-D box IoU (ops/ops_torch.py:692-755 -> pytorch3d _C.iou_box3d, absent: PARITY UNPINNED against the
reference).  The float64 oracle is pinned against an independent half-space intersection + convex hull (scipy) and
against identities; the CUDA kernels against the oracle (gpu)."""
import numpy as np
import pytest


def _boxes(n, seed, extent=6.0, tilt=0.3):
    from objectdetection_3d_b200 import synth
    b, _ = synth.nms_boxes(n=n, seed=seed, extent=extent, tilt=tilt)
    b[:, 2] = np.random.default_rng(seed).uniform(0, 2, n).astype(np.float32)
    return b


def _halfspaces(c):
    """12 -> 6 half-spaces n.x + d <= 0 of the box with corners c (8,3) in the reference's order."""
    c = c.astype(np.float64)
    o, e = c[0], [c[1] - c[0], c[3] - c[0], c[4] - c[0]]
    hs = []
    for k in range(3):
        nrm = np.cross(e[(k + 1) % 3], e[(k + 2) % 3])
        if nrm @ e[k] < 0:
            nrm = -nrm
        nrm /= np.linalg.norm(nrm)
        hs.append(np.r_[-nrm, nrm @ o])                 # -n.x + n.o <= 0
        hs.append(np.r_[nrm, -(nrm @ (o + e[k]))])      #  n.x - n.(o+e) <= 0
    return np.array(hs)


def _independent_volume(c1, c2):
    from scipy.optimize import linprog
    from scipy.spatial import ConvexHull, HalfspaceIntersection
    hs = np.vstack([_halfspaces(c1), _halfspaces(c2)])
    norm = np.linalg.norm(hs[:, :3], axis=1)
    res = linprog([0, 0, 0, -1], A_ub=np.hstack([hs[:, :3], norm[:, None]]), b_ub=-hs[:, 3],
                  bounds=[(None, None)] * 3 + [(0, None)])
    if res.status != 0 or res.x[3] < 1e-7:
        return 0.0
    return ConvexHull(HalfspaceIntersection(hs, res.x[:3]).intersections).volume


def test_oracle_box3d_vs_independent(oracle):
    a, b = _boxes(14, 1), _boxes(16, 2)
    ca, cb = oracle.bbox2corners3D(a), oracle.bbox2corners3D(b)
    vol, iou = oracle.box3d_overlap(ca, cb)
    ref = np.array([[_independent_volume(x, y) for y in cb] for x in ca])
    assert np.abs(vol - ref).max() < 1e-6 * max(1.0, ref.max())
    assert (ref > 1e-3).sum() > 15                      # the sample does contain intersecting pairs
    va = np.prod(a[:, 3:6].astype(np.float64), axis=1)
    vb = np.prod(b[:, 3:6].astype(np.float64), axis=1)
    assert np.allclose(iou, ref / (va[:, None] + vb[None, :] - ref), atol=1e-6)


def test_oracle_box3d_identities(oracle):
    a = _boxes(40, 3)
    ca = oracle.bbox2corners3D(a)
    vol, iou = oracle.box3d_overlap(ca, ca)
    assert np.allclose(np.diag(iou), 1.0, atol=1e-6) and np.allclose(iou, iou.T, atol=1e-6)
    # upright boxes: volume = rotated BEV footprint intersection x z overlap
    u = a.copy(); u[:, 6:8] = 0
    cu = oracle.bbox2corners3D(u)
    vol_u, _ = oracle.box3d_overlap(cu, cu)
    bev = oracle.bbox_iou_rotated_bev(u, u)
    area = (u[:, 3] * u[:, 4]).astype(np.float64)
    inter = bev * (area[:, None] + area[None, :]) / (1 + bev)
    zlo = np.maximum(u[:, None, 2], u[None, :, 2]); zhi = np.minimum((u[:, 2] + u[:, 5])[:, None], (u[:, 2] + u[:, 5])[None, :])
    assert np.abs(vol_u - inter * np.clip(zhi - zlo, 0, None)).max() < 2e-4
    # a box inside another; boxes sharing a face; disjoint boxes
    big = np.array([[0, 0, 0, 4, 4, 4, 0.2, -0.1, 0.7]], np.float32)
    small = np.array([[0.2, -0.1, 1.0, 1, 1, 1, 0.5, 0.3, -0.4]], np.float32)
    v, i = oracle.box3d_overlap(oracle.bbox2corners3D(big), oracle.bbox2corners3D(small))
    assert abs(v[0, 0] - 1.0) < 1e-5 and abs(i[0, 0] - 1.0 / 64.0) < 1e-6
    p = np.array([[0, 0, 0, 2, 2, 2, 0, 0, 0], [2, 0, 0, 2, 2, 2, 0, 0, 0], [1, 0, 0, 2, 2, 2, 0, 0, 0], [9, 9, 0, 1, 1, 1, 0, 0, 0]], np.float32)
    v, i = oracle.box3d_overlap(oracle.bbox2corners3D(p), oracle.bbox2corners3D(p))
    assert abs(v[0, 1]) < 1e-6 and abs(v[0, 2] - 4.0) < 1e-6 and v[0, 3] == 0 and abs(i[0, 2] - 4.0 / 12.0) < 1e-6


# two boxes of different height standing on the same tilted plane (an anchor and a ground truth with equal rx): their
# bottom faces are coplanar up to the fp32 rounding of the corners.  Exact volume: x overlap x 1.3 x 17.
_COPLANAR = np.array([[18.550724, 0.6779661, 0., 1., 1.75, 20., 0.3142, 0., 0.],
                      [17.971014, 0.6779661, 0., 1.3, 1.3, 17., 0.3142, 0., 0.]], np.float32)
_COPLANAR_VOL = (float(_COPLANAR[1, 0]) + 0.65 - (float(_COPLANAR[0, 0]) - 0.5)) * 1.3 * 17.0


def test_oracle_box3d_nearly_coplanar_faces(oracle):
    c = oracle.bbox2corners3D(_COPLANAR)
    vol, iou = oracle.box3d_overlap(c, c)
    assert abs(vol[0, 1] - _COPLANAR_VOL) < 2e-5 and abs(vol[1, 0] - _COPLANAR_VOL) < 2e-5     # either argument order
    assert abs(vol[0, 1] - vol[1, 0]) < 1e-6 and abs(iou[0, 1] - iou[1, 0]) < 1e-7


@pytest.mark.gpu
def test_cuda_box3d_overlap_checks_and_nms(oracle):
    import torch
    from objectdetection_3d_b200 import model_utils, ops_torch, synth
    a, b = _boxes(300, 5, 12.0), _boxes(400, 6, 12.0)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    ca, cb = ops_torch.bbox2corners3D(ta), ops_torch.bbox2corners3D(tb)
    vol, iou = ops_torch.box3d_overlap(ca, cb, return_vol=True)
    ovol, oiou = oracle.box3d_overlap(ca.cpu().numpy(), cb.cpu().numpy())
    # the kernel defines the IoU as 0 when the xy bounding rectangles of the corners do not overlap (exact for boxes)
    # the kernel evaluates the volume in float64 (a different algorithm from the oracle's polygon clipping): what is left
    # is the rounding of the float32 outputs
    assert np.abs(vol.cpu().numpy() - ovol).max() < 4e-7 * max(1.0, ovol.max())
    assert np.abs(iou.cpu().numpy() - oiou).max() < 1e-6
    assert (oiou > 0.01).sum() > 200
    assert torch.equal(ops_torch.box3d_overlap(ca, cb), iou)                          # default return: iou only
    d = ops_torch.box3d_overlap(ca, ca).cpu().numpy()
    assert np.abs(np.diag(d) - 1).max() < 1e-6 and np.array_equal(d, d.T)             # symmetric bit for bit
    assert ops_torch.box3d_overlap(ca[:0], cb).shape == (0, 400)
    cc = ops_torch.bbox2corners3D(torch.from_numpy(_COPLANAR).cuda())
    vc, _ = ops_torch.box3d_overlap(cc, cc, return_vol=True)
    assert abs(vc[0, 1].item() - _COPLANAR_VOL) < 2e-5 and vc[0, 1].item() == vc[1, 0].item()
    # validity checks: same exceptions as ops/ops_torch.py:743-748
    with pytest.raises(ValueError, match="shape"):
        ops_torch.box3d_overlap(ca[:, :4], cb)
    flat = a[:3].copy(); flat[1, 5] = 0.0
    with pytest.raises(ValueError, match="zero areas"):
        ops_torch.box3d_overlap(ops_torch.bbox2corners3D(torch.from_numpy(flat).cuda()), cb)
    bent = ca[:3].clone(); bent[2, 6, 2] += 0.5
    with pytest.raises(ValueError, match="not coplanar"):
        ops_torch.box3d_overlap(bent, cb)
    # nms_dim == 3: greedy NMS consistent with the kernel's own 3-D IoU matrix
    for n, extent in ((2500, 20.0), (20000, 40.0)):
        boxes, scores = synth.nms_boxes(n=n, seed=9, extent=extent, tilt=0.3)
        tbx, tsc = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
        for sthr, ithr in ((0.3, 0.1), (0.05, 0.5)):
            keep = model_utils.multiclass_nms(tbx, tsc, sthr, ithr, 3)[0].cpu().numpy()
            cand = np.nonzero(scores[:, 0] > np.float32(sthr))[0]
            order = cand[np.argsort(-scores[cand, 0], kind="stable")]
            assert (np.diff(scores[keep, 0]) < 0).all() and set(keep.tolist()) <= set(order.tolist())
            kc = ops_torch.bbox2corners3D(tbx[torch.from_numpy(keep).cuda()])
            kk = ops_torch.box3d_overlap(kc, kc).cpu().numpy()
            np.fill_diagonal(kk, 0)
            assert not (kk > np.float32(ithr)).any()
            dropped = np.setdiff1d(order, keep)
            if len(dropped):
                dk = ops_torch.box3d_overlap(ops_torch.bbox2corners3D(tbx[torch.from_numpy(dropped).cuda()]), kc).cpu().numpy()
                better = scores[keep, 0][None, :] > scores[dropped, 0][:, None]
                assert ((dk > np.float32(ithr)) & better).any(axis=1).all()
    # small case against a CPU greedy loop on the float64 oracle IoU (thresholds away from any pair's IoU)
    boxes, scores = synth.nms_boxes(n=400, seed=11, extent=8.0, tilt=0.3)
    oc = oracle.bbox2corners3D(boxes)
    _, om = oracle.box3d_overlap(oc, oc)
    order = np.argsort(-scores[:, 0], kind="stable")
    vals = np.unique(om)
    for target in (0.1, 0.3):
        # a threshold near the target that no pair's IoU comes within 1e-4 of (fp32 kernel vs float64 oracle)
        near = vals[(vals > target - 0.02) & (vals < target + 0.02)]
        edges = np.r_[target - 0.02, near, target + 0.02]
        k = int(np.argmax(np.diff(edges)))
        ithr = float(np.float32(0.5 * (edges[k] + edges[k + 1])))
        assert np.abs(om - ithr).min() > 1e-4
        alive, kept = np.ones(len(order), bool), []
        for r, i in enumerate(order):
            if alive[r]:
                kept.append(i)
                alive[r + 1:] &= ~(om[order[r + 1:], i] > ithr)
        got = model_utils.multiclass_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), 0.0, ithr, 3)[0]
        assert np.array_equal(got.cpu().numpy(), np.array(kept))

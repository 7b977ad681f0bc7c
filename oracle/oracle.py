"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (oracle/pp_oracle.c).

Function names and argument order follow the reference (michalp0lak/ObjectDetection_3D)
so that parity tests read like calls into the reference itself.  Nothing under
``objectdetection_3d_b200/`` imports this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.

Parity status: pinned against the reference (see the header of pp_oracle.c).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpp_oracle.so")
_lib = None

c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    src = os.path.join(_HERE, "pp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libpp_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.ppo_points_to_voxel.restype = ctypes.c_int64
        _lib.ppo_nms_sorted.restype = ctypes.c_int64
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(ty)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# --------------------------------------------------------------------------- voxelize
def regime(points, voxel_size, coors_range):
    """ops/ops_numba.py:139-142: lists are cast to points.dtype, ndarrays keep their dtype.

    Returns (vsize f64[3], range f64[6], range_is_f64, vsize_is_f64, grid int32[3]).
    """
    if not isinstance(voxel_size, np.ndarray):
        voxel_size = np.array(voxel_size, dtype=points.dtype)
    if not isinstance(coors_range, np.ndarray):
        coors_range = np.array(coors_range, dtype=points.dtype)
    grid = (coors_range[3:] - coors_range[:3]) / voxel_size            # :144
    grid = np.round(grid).astype(np.int32)                              # :145
    r64 = coors_range.dtype != np.float32 and coors_range.dtype != np.float16
    v64 = voxel_size.dtype != np.float32 and voxel_size.dtype != np.float16
    return (voxel_size.astype(np.float64), coors_range.astype(np.float64), int(r64), int(v64), grid)


def numba_argsort_desc(keys):
    """points[:, 3].argsort()[::-1] with numba's quicksort tie order (ops_numba.py:262)."""
    keys = np.asarray(keys)
    assert keys.dtype == np.float32
    n = keys.shape[0]
    perm = np.empty(n, dtype=np.int64)
    stride = keys.strides[0] // 4 if n else 1
    rc = lib().ppo_numba_argsort_desc(
        ctypes.cast(keys.ctypes.data, c_f32p), ctypes.c_int64(stride), ctypes.c_int64(n), _p(perm, c_i64p))
    assert rc == 0
    return perm


def points_to_voxel(points, voxel_size, coors_range, max_points, max_voxels, reflectance_sampling,
                    perm=None):
    """ops/ops_numba.py:109-168.  reflectance_sampling=False processes the GIVEN order (the
    reference shuffles the caller's array in place first, :190; replay that order)."""
    points = _f32(points)
    N, C = points.shape
    vs, rg, r64, v64, grid = regime(points, voxel_size, coors_range)
    if reflectance_sampling and perm is None:
        perm = numba_argsort_desc(points[:, 3])
    voxels = np.zeros((max_voxels, max_points, C), dtype=np.float32)
    coors = np.zeros((max_voxels, 3), dtype=np.int32)
    num = np.zeros((max_voxels,), dtype=np.int32)
    ws = np.empty(int(np.prod(grid.astype(np.int64))), dtype=np.int32)
    permp = _p(np.ascontiguousarray(perm, dtype=np.int64), c_i64p) if perm is not None else None
    vn = lib().ppo_points_to_voxel(
        _p(points, c_f32p), ctypes.c_int64(N), C, _p(vs, c_f64p), _p(rg, c_f64p), r64, v64,
        _p(grid, c_i32p), int(max_points), int(max_voxels), permp,
        _p(voxels, c_f32p), _p(coors, c_i32p), _p(num, c_i32p), _p(ws, c_i32p))
    return voxels[:vn], coors[:vn], num[:vn]


def pointpillars_voxelization(points, voxel_size, point_cloud_range, max_voxel_points, max_voxels):
    """model/PointPillars.py:330-354: voxel_size -> f32 (VoxelGenerator, ops_numba.py:48),
    range stays f64 (:324); outputs f32 / int64 zyx / int64."""
    vs = np.array(np.array(voxel_size), dtype=np.float32)
    rg = np.array(point_cloud_range)
    if rg.dtype.kind != "f":
        rg = rg.astype(np.float64)
    v, c, n = points_to_voxel(points, vs, rg, max_voxel_points, max_voxels, True)
    return v, c[:, [2, 1, 0]].astype(np.int64), n.astype(np.int64)


# --------------------------------------------------------------------------- decorate / PFN / scatter
def decorate(voxels, num_points, coors, vx, vy, x_offset, y_offset):
    voxels = _f32(voxels)
    M, P, C = voxels.shape
    out = np.empty((M, P, C + 5), dtype=np.float32)
    num_points = np.ascontiguousarray(num_points, dtype=np.int64)
    coors = np.ascontiguousarray(coors, dtype=np.int64)
    lib().ppo_decorate(_p(voxels, c_f32p), _p(num_points, c_i64p), _p(coors, c_i64p), ctypes.c_int64(M),
                       P, C, ctypes.c_double(vx), ctypes.c_double(vy), ctypes.c_double(x_offset),
                       ctypes.c_double(y_offset), _p(out, c_f32p))
    return out


def pfn_layer(x, weight, gamma, beta, rmean, rvar, eps=1e-3, last_layer=True):
    x = _f32(x)
    M, P, Cin = x.shape
    U = weight.shape[0]
    out = np.empty((M, U) if last_layer else (M, P, 2 * U), dtype=np.float32)
    w, g, b, mu, var = (_f32(a) for a in (weight, gamma, beta, rmean, rvar))
    lib().ppo_pfn_layer(_p(x, c_f32p), ctypes.c_int64(M), P, Cin, _p(w, c_f32p), _p(g, c_f32p),
                        _p(b, c_f32p), _p(mu, c_f32p), _p(var, c_f32p), ctypes.c_double(eps), U,
                        int(last_layer), _p(out, c_f32p))
    return out


def pillar_feature_net(voxels, num_points, coors, layers, voxel_size, point_cloud_range, eps=1e-3):
    """model/PointPillars.py:480-526.  layers = list of dicts(weight,gamma,beta,mean,var)."""
    vx, vy = voxel_size[0], voxel_size[1]
    x = decorate(voxels, num_points, coors, vx, vy, vx / 2 + point_cloud_range[0],
                 vy / 2 + point_cloud_range[1])
    for i, L in enumerate(layers):
        x = pfn_layer(x, L["weight"], L["gamma"], L["beta"], L["mean"], L["var"], eps,
                      last_layer=(i == len(layers) - 1))
    return np.concatenate([x, np.asarray(num_points, dtype=np.float32).reshape(-1, 1)], axis=1)  # :526


def scatter_dense(feat, coors, batch_size, D, H, W):
    feat = _f32(feat)
    M, C = feat.shape
    coors = np.ascontiguousarray(coors, dtype=np.int32)
    canvas = np.zeros((batch_size, C * D, H, W), dtype=np.float32)
    lib().ppo_scatter_dense(_p(feat, c_f32p), _p(coors, c_i32p), ctypes.c_int64(M), C, batch_size, D, H, W,
                            _p(canvas, c_f32p))
    return canvas


# --------------------------------------------------------------------------- boxes
def bbox2corners3D(boxes):
    boxes = _f32(boxes)
    out = np.empty((boxes.shape[0], 8, 3), dtype=np.float32)
    lib().ppo_box_corners3d(_p(boxes, c_f32p), ctypes.c_int64(boxes.shape[0]), _p(out, c_f32p))
    return out


def bbox2rotated_corners2D(boxes):
    boxes = _f32(boxes)
    out = np.empty((boxes.shape[0], 4), dtype=np.float32)
    lib().ppo_box_aabb2d(_p(boxes, c_f32p), ctypes.c_int64(boxes.shape[0]), _p(out, c_f32p))
    return out


_MODES = {"iou": 0, "iof": 1, "giou": 2}


def bbox_iou2D(b1, b2, mode="iou", eps=1e-6):
    b1, b2 = _f32(b1), _f32(b2)
    out = np.empty((b1.shape[0], b2.shape[0]), dtype=np.float32)
    lib().ppo_bbox_iou2d(_p(b1, c_f32p), ctypes.c_int64(b1.shape[0]), _p(b2, c_f32p),
                         ctypes.c_int64(b2.shape[0]), _MODES[mode], ctypes.c_double(eps), _p(out, c_f32p))
    return out


def iou_jit(boxes, query_boxes, eps=0.0):
    b, q = _f32(boxes), _f32(query_boxes)
    out = np.zeros((b.shape[0], q.shape[0]), dtype=np.float32)
    lib().ppo_iou_jit(_p(b, c_f32p), ctypes.c_int64(b.shape[0]), _p(q, c_f32p), ctypes.c_int64(q.shape[0]),
                      ctypes.c_double(eps), _p(out, c_f32p))
    return out


def nms_sorted(rect, iou_thr):
    rect = _f32(rect)
    n = rect.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = lib().ppo_nms_sorted(_p(rect, c_f32p), ctypes.c_int64(n), ctypes.c_double(iou_thr), _p(keep, c_i64p))
    return keep[:k]


def multiclass_nms(boxes, scores, score_thr, iou_thr, nms_dim=2):
    """model/utils.py:353-426 (nms_dim == 2).  Returns, per class, the kept original indices in
    descending-score order (the reference returns the same SET in CPython set order, :423)."""
    assert nms_dim == 2, "oracle pins the AABB2D form only (SURVEY.md 8c)"
    boxes, scores = _f32(boxes), _f32(scores)
    out = []
    for c in range(scores.shape[1]):
        cand = np.nonzero(scores[:, c] > np.float32(score_thr))[0]          # :381
        if cand.size == 0:
            out.append(np.empty((0,), dtype=np.int64))
            continue
        order = np.argsort(-scores[cand, c], kind="stable")                  # :398 (unique scores)
        orig = cand[order]
        rect = bbox2rotated_corners2D(boxes[orig])                           # :397
        keep = nms_sorted(rect, iou_thr)
        out.append(orig[keep].astype(np.int64))
    return out


def box_encode(src, dst):
    src, dst = _f32(src), _f32(dst)
    out = np.empty_like(src)
    lib().ppo_box_encode(_p(src, c_f32p), _p(dst, c_f32p), ctypes.c_int64(src.shape[0]), _p(out, c_f32p))
    return out


def box_decode(anchors, deltas):
    anchors, deltas = _f32(anchors), _f32(deltas)
    out = np.empty_like(anchors)
    lib().ppo_box_decode(_p(anchors, c_f32p), _p(deltas, c_f32p), ctypes.c_int64(anchors.shape[0]),
                         _p(out, c_f32p))
    return out


def limit_period(val, offset=0.5, period=np.pi):
    val = _f32(val)
    out = np.empty_like(val)
    lib().ppo_limit_period(_p(val, c_f32p), ctypes.c_int64(val.size), ctypes.c_double(offset),
                           ctypes.c_double(period), _p(out, c_f32p))
    return out


def grid_anchors(featmap_size, anchor_range, sizes, rotations):
    """model/utils.py:168-264 for a single range.  Returns (D,H,W,S,R,9) f32."""
    if len(featmap_size) == 2:
        featmap_size = [1, featmap_size[0], featmap_size[1]]
    D, H, W = (int(v) for v in featmap_size)
    rg = _f32(anchor_range)
    sz = _f32(sizes).reshape(-1, 3)
    rt = _f32(rotations).reshape(-1, 3)
    out = np.empty((D, H, W, sz.shape[0], rt.shape[0], 9), dtype=np.float32)
    lib().ppo_grid_anchors(_p(rg, c_f32p), _p(sz, c_f32p), sz.shape[0], _p(rt, c_f32p), rt.shape[0],
                           D, H, W, _p(out, c_f32p))
    return out


def bbox_iou_rotated_bev(b1, b2):
    """Rotated BEV IoU of 9-parameter boxes (footprint x, y, dx, dy, rz), float64.  Extension, see pp_oracle.c."""
    b1, b2 = _f32(b1), _f32(b2)
    out = np.empty((b1.shape[0], b2.shape[0]), dtype=np.float64)
    lib().ppo_iou_rotated_bev(_p(b1, c_f32p), ctypes.c_int64(b1.shape[0]), _p(b2, c_f32p), ctypes.c_int64(b2.shape[0]),
                              _p(out, c_f64p))
    return out


def box3d_overlap(c1, c2):
    """Oriented 3-D box intersection volume and IoU from corners (n,8,3),(m,8,3), float64.  PARITY UNPINNED against
    the reference (pytorch3d _C.iou_box3d is absent); pinned against scipy in tests/test_box3d.py."""
    c1, c2 = _f32(c1).reshape(-1, 24), _f32(c2).reshape(-1, 24)
    vol = np.empty((c1.shape[0], c2.shape[0]), dtype=np.float64)
    iou = np.empty_like(vol)
    lib().ppo_box3d_overlap(_p(c1, c_f32p), ctypes.c_int64(c1.shape[0]), _p(c2, c_f32p), ctypes.c_int64(c2.shape[0]),
                            _p(vol, c_f64p), _p(iou, c_f64p))
    return vol, iou


def assign_overlaps(overlaps, lo_thr):
    """The reductions of Anchor3DHead.assign_bboxes on a (G, A) IoU matrix, model/PointPillars.py:968-978:
    max / first argmax over the ground truths, max over the anchors, low-quality-match flags."""
    ov = np.asarray(overlaps)
    gt_max = ov.max(axis=1)
    lowq = ((ov == gt_max[:, None]) & (gt_max >= ov.dtype.type(lo_thr))[:, None]).any(axis=0)
    return ov.max(axis=0), ov.argmax(axis=0), gt_max, lowq


# ---- either side of the path -------------------------------------------------------------------------
def global_outlier_check(point_cloud):
    """ops/ops_numpy.py:111-115, restated op for op in numpy (the reference is numpy)."""
    norm = np.sum((point_cloud[:, :3] - np.mean(point_cloud[:, :3], axis=0)) ** 2, axis=1) ** 0.5
    return point_cloud[norm < np.mean(norm) + 5 * np.std(norm), :]


def preprocess_points(points, point_cloud_range, input_features, outlier=True):
    """model/PointPillars.py:241-266 without the bbox filter and the augmentation: outlier check, range filter,
    feature selection."""
    if outlier:
        points = global_outlier_check(points)
    points = np.array(points, dtype=np.float32)
    lo, hi = np.array(point_cloud_range[:3]), np.array(point_cloud_range[3:])
    points = points[np.where(np.all(np.logical_and(points[:, :3] >= lo, points[:, :3] < hi), axis=-1))]
    return points[:, input_features]


def custom_voxelizer_voxelize(point_cloud, voxel_size, max_voxel_points, reflectance_sampling, perm=None):
    """model/utils.py:15-43 (CustomVoxelizer.voxelize) on the oracle voxelizer: the cloud's own min/max as a Python list
    (-> all-f32 cell arithmetic), density-dependent pillar cap, centroids with the point count appended.
    Raises UnboundLocalError like the reference when neither branch voxelizes (:43 reads an unbound `vp`)."""
    rng = point_cloud[:, :3].min(axis=0).tolist() + point_cloud[:, :3].max(axis=0).tolist()
    dims = point_cloud[:, :3].max(axis=0) - point_cloud[:, :3].min(axis=0)
    density = point_cloud.shape[0] / np.prod(dims)
    a, b, c, voxel_limit = 20000, 0.01, 70000, 3000000
    vp = None
    if density > 10:
        max_voxels = np.min([int(a * np.exp(b * density) + c), point_cloud.shape[0]])
        if max_voxels <= point_cloud.shape[0]:
            max_voxels = np.min([max_voxels, voxel_limit])
            vox, _, vp = points_to_voxel(point_cloud, np.array(voxel_size, dtype=np.float32), rng, max_voxel_points,
                                         int(max_voxels), reflectance_sampling, perm=perm)
            point_cloud = np.sum(vox, axis=1) / vp.reshape(-1, 1)
    elif point_cloud.shape[0] > voxel_limit:
        vox, _, vp = points_to_voxel(point_cloud, np.array(voxel_size, dtype=np.float32), rng, max_voxel_points,
                                     voxel_limit, reflectance_sampling, perm=perm)
        point_cloud = np.sum(vox, axis=1) / vp.reshape(-1, 1)
    if vp is None:
        raise UnboundLocalError("cannot access local variable 'vp' where it is not associated with a value")
    return np.concatenate((point_cloud, vp.reshape(-1, 1)), axis=1)


def dense_to_sparse(x):
    """model/PointPillars.py:766-789: non-empty cells of a dense (B,C,H,W) map in row-major order."""
    coords, values = [], []
    for i in range(x.shape[0]):
        ys, xs = np.where((x[i] != 0).any(axis=0))
        coords.append(np.stack([np.full_like(ys, i), ys, xs], axis=1).astype(np.int32))
        values.append(x[i][:, ys, xs].T)
    return np.concatenate(values, axis=0), np.concatenate(coords, axis=0)

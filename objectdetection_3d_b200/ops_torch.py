"""Drop-in for the hot-path functions of the reference's ``ops/ops_torch.py`` (same names and
signatures), backed by csrc/pp_boxes.cu.  Inputs are CUDA float32 tensors; outputs are CUDA tensors.
"""
import torch

from . import _lib
from .ops_numba import _dev, _ptr, _stream


def _f32c(t):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t, dtype=torch.float32)
    if not t.is_cuda:
        t = t.to(_dev())
    return t.to(torch.float32).contiguous()


def bbox2rotated_corners2D(bbxs):
    """ops/ops_torch.py:13-114: (N,9) boxes -> (N,4) [xmin, ymin, xmax, ymax] of the rotated corners."""
    b = _f32c(bbxs)
    out = torch.empty((b.shape[0], 4), dtype=torch.float32, device=b.device)
    _lib.check(_lib.load().pp_box_aabb2d(_ptr(b), b.shape[0], _ptr(out), _stream()))
    return out


def bbox2corners3D(bbxs):
    """ops/ops_torch.py:160-256: (N,9) boxes -> (N,8,3) rotated corners."""
    b = _f32c(bbxs)
    out = torch.empty((b.shape[0], 8, 3), dtype=torch.float32, device=b.device)
    _lib.check(_lib.load().pp_box_corners3d(_ptr(b), b.shape[0], _ptr(out), _stream()))
    return out


def bbox_iou2D(bboxes1, bboxes2, mode='iou', eps=1e-6):
    """ops/ops_torch.py:538-607: (m,4),(n,4) -> (m,n).  Same assertions as the reference."""
    assert mode in ['iou', 'iof', 'giou'], f'Unsupported mode {mode}'
    assert (bboxes1.size(-1) == 4 or bboxes1.size(0) == 0)
    assert (bboxes2.size(-1) == 4 or bboxes2.size(0) == 0)
    assert bboxes1.shape[:-2] == bboxes2.shape[:-2]
    assert bboxes1.dim() == 2, "batched (B, m, 4) input is not on the hot path"
    b1, b2 = _f32c(bboxes1), _f32c(bboxes2)
    m, n = b1.shape[0], b2.shape[0]
    out = torch.empty((m, n), dtype=torch.float32, device=b1.device)
    if m * n == 0:
        return out
    _lib.check(_lib.load().pp_bbox_iou2d(_ptr(b1), m, _ptr(b2), n, _lib.IOU_MODES[mode], float(eps), _ptr(out),
                                          _stream()), AssertionError)
    return out


def bbox_iou_rotated_bev(bboxes1, bboxes2):
    """Extension (north star "rotated BEV IoU"): IoU of the rotated BEV footprints (x, y, dx, dy, rz) of two sets of
    9-parameter boxes, (m,9),(n,9) -> (m,n).  Not part of the reference (its nms_dim == 2 path is the AABB form)."""
    assert bboxes1.size(-1) == 9 and bboxes2.size(-1) == 9
    b1, b2 = _f32c(bboxes1), _f32c(bboxes2)
    m, n = b1.shape[0], b2.shape[0]
    out = torch.empty((m, n), dtype=torch.float32, device=b1.device)
    if m * n == 0:
        return out
    _lib.check(_lib.load().pp_iou_rotated_bev(_ptr(b1), m, _ptr(b2), n, _ptr(out), _stream()))
    return out


def _box3d_flags(boxes, eps):
    c = _f32c(boxes)
    flags = torch.empty((c.shape[0],), dtype=torch.int32, device=c.device)
    if c.shape[0]:
        _lib.check(_lib.load().pp_box3d_check(_ptr(c), c.shape[0], float(eps), _ptr(flags), _stream()))
    return flags


def check_coplanar(boxes, eps=1e-4):
    """ops/ops_torch.py:610-648: raises ValueError when the plane residual test fails for any box."""
    bad = (_box3d_flags(boxes, eps) & 1) != 0
    if bad.any().item():
        raise ValueError("Plane vertices are not coplanar. This applies for bboxes in positions: {}".format(
            torch.arange(0, boxes.shape[0])[bad.cpu()]))


def check_nonzero(boxes, eps=1e-4):
    """ops/ops_torch.py:651-690: raises ValueError when a face triangle of any box has area < eps."""
    bad = (_box3d_flags(boxes, eps) & 2) != 0
    if bad.any().item():
        raise ValueError("Planes have zero areas. This applies for bboxes in positions: {}".format(
            torch.arange(0, boxes.shape[0])[bad.cpu()]))


def box3d_overlap(boxes1, boxes2, eps=1e-2, return_vol=False):
    """ops/ops_torch.py:711-755: oriented 3-D IoU of boxes given by their 8 corners, (N,8,3),(M,8,3) -> iou (N,M)
    (the reference returns only the IoU, :753-755).  The reference delegates to pytorch3d 0.7.4 `_C.iou_box3d`, which is
    not part of its checkout: this kernel computes the exact convex intersection volume of the two boxes (taken as
    the parallelepipeds v0; v1-v0, v3-v0, v4-v0) and is validated against an independent float64 computation
    (tests/test_box3d.py), not against pytorch3d.  Same ValueError behaviour as the reference's validity checks."""
    if not all((8, 3) == tuple(box.shape[1:]) for box in [boxes1, boxes2]):
        raise ValueError("Each box in the batch must be of shape (8, 3)")
    c1, c2 = _f32c(boxes1), _f32c(boxes2)
    for c in (c1, c2):                       # one flag kernel + one sync per operand (the reference: two each)
        fl = _box3d_flags(c, eps)
        nc, nz = (fl & 1) != 0, (fl & 2) != 0
        if nc.any().item():
            raise ValueError("Plane vertices are not coplanar. This applies for bboxes in positions: {}".format(
                torch.arange(0, c.shape[0])[nc.cpu()]))
        if nz.any().item():
            raise ValueError("Planes have zero areas. This applies for bboxes in positions: {}".format(
                torch.arange(0, c.shape[0])[nz.cpu()]))
    n, m = c1.shape[0], c2.shape[0]
    iou = torch.empty((n, m), dtype=torch.float32, device=c1.device)
    vol = torch.empty((n, m), dtype=torch.float32, device=c1.device) if return_vol else None
    if n * m:
        _lib.check(_lib.load().pp_box3d_overlap(_ptr(c1), n, _ptr(c2), m, _ptr(vol) if return_vol else None, _ptr(iou),
                                                _stream()))
    return (vol, iou) if return_vol else iou

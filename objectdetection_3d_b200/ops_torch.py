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


def box3d_overlap(boxes1, boxes2, eps=1e-2):
    """ops/ops_torch.py:711-755 (pytorch3d oriented 3-D IoU).  Out of the pinned scope: SURVEY.md
    section 8(f) rank 1 ("next"); the reference's own implementation lives in an absent third-party
    library, so there is no oracle to pin it against yet."""
    raise NotImplementedError("BOX3D oriented IoU is not built yet (SURVEY.md 8f); use nms_dim=2")

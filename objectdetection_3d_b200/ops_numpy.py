"""Drop-in for the one function of the reference's ops/ops_numpy.py that sits directly before the hot path:
global_outlier_check (:111-115), plus the device-side form of PointPillars.preprocess's point filtering
(model/PointPillars.py:241-266) that it belongs to.  CUDA only; no CPU fallback."""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops_numba import _dev, _ptr, _stream


def preprocess_points(points, point_cloud_range=None, input_features=None, outlier_check=True, exact=True):
    """Outlier check (ops/ops_numpy.py:111-115) + range filter (model/PointPillars.py:251-252) + feature selection
    (:266) in one pass over a tile that is uploaded once.  points: numpy (N,C) or CUDA tensor; returns the same kind,
    rows in their original order.  point_cloud_range None: no range filter (the bare global_outlier_check).
    exact=True (default): the 5-sigma statistics follow numpy's own float32 order of operations, so the kept rows
    are the reference's bit for bit (float32 input); exact=False: float64 statistics, fully parallel and ~10x
    faster on 2e5 points, rows within rounding of the threshold can differ."""
    lib = _lib.load()
    is_numpy = isinstance(points, np.ndarray)
    p = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).to(_dev()) if is_numpy else points.float().contiguous()
    n, C = p.shape
    feats = list(range(C)) if input_features is None else [int(f) for f in input_features]
    big = np.float32(3.0e38)
    rg = np.asarray([-big] * 3 + [big] * 3 if point_cloud_range is None else point_cloud_range, dtype=np.float32)
    fa = np.asarray(feats, dtype=np.int32)
    out = torch.empty((max(n, 1), len(feats)), dtype=torch.float32, device=p.device)
    count = torch.zeros((1,), dtype=torch.int32, device=p.device)
    ws_bytes = int(lib.pp_preprocess_workspace_bytes(n))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=p.device)
    _lib.check(lib.pp_preprocess_points(_ptr(p), n, C, (1 if exact else 2) if outlier_check else 0,
                                        rg.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                        fa.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(feats), _ptr(out),
                                        _ptr(count), _ptr(ws), ws_bytes, _stream()))
    out = out[:int(count.item())]
    return out.cpu().numpy() if is_numpy else out


def global_outlier_check(point_cloud):
    """ops/ops_numpy.py:111-115: drop the points farther than mean + 5 sigma from the centroid."""
    return preprocess_points(point_cloud, None, None, True)

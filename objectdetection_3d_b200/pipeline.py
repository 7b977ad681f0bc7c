"""Device-resident frame pipeline: voxelize -> decorate + PFN -> dense scatter, and the NMS stage,
with every buffer preallocated so that a frame is a fixed sequence of kernel launches on one stream
(no allocation, no host synchronisation).  This is the path bench.py measures; it is the same C ABI
the drop-in modules call, minus the per-call allocations and the int64 / zyx conversions that
PointPillars.voxelize (model/PointPillars.py:106-134) performs between the stages.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops_numba import _ptr, voxel_cfg


def _sp(stream):
    return ctypes.c_void_p(stream.cuda_stream)


class FramePipeline:
    """One frame: points (N,C) on device -> BEV canvas (1, U+1, H, W) on device."""

    def __init__(self, geom, pfn, n_points, num_feats=4, order=_lib.ORDER_REFLECTANCE_DESC, device=None):
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        vs = np.array(geom["voxel_size"], dtype=np.float32)                     # ops_numba.py:48
        rg = np.array(geom["point_cloud_range"], dtype=np.float64)              # PointPillars.py:324
        self.cfg = voxel_cfg(np.float32, vs, rg, geom["max_voxel_points"], geom["max_voxels"], num_feats)
        self.order = order
        self.n_points = int(n_points)
        self.P, self.C = int(self.cfg.max_points), int(num_feats)
        self.W, self.H, self.D = (int(self.cfg.grid[i]) for i in range(3))
        d = self.device
        self.rows = int(self.lib.pp_voxelize_max_rows(self.n_points, ctypes.byref(self.cfg)))
        self.voxels = torch.empty((self.rows, self.P, self.C), dtype=torch.float32, device=d)
        self.coors = torch.empty((self.rows, 3), dtype=torch.int32, device=d)
        self.num = torch.empty((self.rows,), dtype=torch.int32, device=d)
        self.voxel_num = torch.zeros((1,), dtype=torch.int32, device=d)
        self.vox_ws_bytes = int(self.lib.pp_voxelize_workspace_bytes(self.n_points, ctypes.byref(self.cfg), order))
        self.vox_ws = torch.empty((self.vox_ws_bytes,), dtype=torch.uint8, device=d)
        # PFN (single layer, eval): BatchNorm folded like PFNLayer.folded()
        w = torch.as_tensor(pfn["weight"], dtype=torch.float32, device=d).contiguous()
        scale = torch.as_tensor(pfn["gamma"] / np.sqrt(pfn["var"] + np.float32(1e-3)), dtype=torch.float32, device=d)
        shift = torch.as_tensor(pfn["beta"], dtype=torch.float32, device=d) - \
            torch.as_tensor(pfn["mean"], dtype=torch.float32, device=d) * scale
        self.w, self.scale, self.shift = w, scale.contiguous(), shift.contiguous()
        self.U = int(w.shape[0])
        assert w.shape[1] == self.C + 5
        self.feat = torch.empty((self.rows, self.U + 1), dtype=torch.float32, device=d)
        # cell -> pillar map written by the voxelizer, consumed by the canvas kernel (no separate map build)
        self.pillar_map = torch.empty((self.D, self.H, self.W), dtype=torch.int32, device=d)
        vx, vy = float(geom["voxel_size"][0]), float(geom["voxel_size"][1])
        self.vx, self.vy = vx, vy
        self.x_off = vx / 2 + geom["point_cloud_range"][0]
        self.y_off = vy / 2 + geom["point_cloud_range"][1]
        # gather + PFN in one kernel (pp_voxelize_features) when the shapes allow it
        self.fused = self.C == 4 and self.P <= 32 and self.U <= 64
        self.pfn_args = _lib.PfnFused(self.w.data_ptr(), self.scale.data_ptr(), self.shift.data_ptr(), self.U,
                                      self.vx, self.vy, self.x_off, self.y_off, self.feat.data_ptr())

    def new_canvas(self):
        return torch.empty((1, (self.U + 1) * self.D, self.H, self.W), dtype=torch.float32, device=self.device)

    def voxelize(self, points, stream):
        n = points.shape[0]
        assert n <= self.n_points and points.shape[1] == self.C
        _lib.check(self.lib.pp_voxelize(_ptr(points), n, ctypes.byref(self.cfg), self.order, None, _ptr(self.voxels),
                                        _ptr(self.coors), _ptr(self.num), _ptr(self.voxel_num), _ptr(self.pillar_map),
                                        _ptr(self.vox_ws), self.vox_ws_bytes, _sp(stream)))

    def encode_scatter(self, canvas, stream):
        _lib.check(self.lib.pp_pillar_features(
            _ptr(self.voxels), _ptr(self.num), _lib.NUM_I32, _ptr(self.coors), _lib.COORS_XYZ_I32, self.rows,
            _ptr(self.voxel_num), self.P, self.C, self.vx, self.vy, self.x_off, self.y_off, _ptr(self.w),
            _ptr(self.scale), _ptr(self.shift), self.U, _ptr(self.feat), _sp(stream)))
        _lib.check(self.lib.pp_scatter_mapped(_ptr(self.feat), _ptr(self.pillar_map), self.U + 1, 1, self.D, self.H,
                                              self.W, _ptr(canvas), _sp(stream)))

    def voxelize_features(self, points, stream):
        """voxelize + PillarFeatureNet: self.voxels / coors / num / voxel_num / pillar_map AND self.feat."""
        n = points.shape[0]
        assert self.fused and n <= self.n_points and points.shape[1] == self.C
        _lib.check(self.lib.pp_voxelize_features(
            _ptr(points), n, ctypes.byref(self.cfg), self.order, None, _ptr(self.voxels), _ptr(self.coors),
            _ptr(self.num), _ptr(self.voxel_num), _ptr(self.pillar_map), ctypes.byref(self.pfn_args), _ptr(self.vox_ws),
            self.vox_ws_bytes, _sp(stream)))

    def voxelize_scatter(self, points, canvas, stream):
        """The frame in one call: voxelize + PillarFeatureNet + dense scatter (pp_voxelize_scatter)."""
        n = points.shape[0]
        assert self.fused and n <= self.n_points and points.shape[1] == self.C
        assert canvas.is_contiguous() and canvas.numel() == (self.U + 1) * self.D * self.H * self.W
        _lib.check(self.lib.pp_voxelize_scatter(
            _ptr(points), n, ctypes.byref(self.cfg), self.order, None, _ptr(self.voxels), _ptr(self.coors),
            _ptr(self.num), _ptr(self.voxel_num), None, ctypes.byref(self.pfn_args), _ptr(canvas), _ptr(self.vox_ws),
            self.vox_ws_bytes, _sp(stream)))

    def scatter(self, canvas, stream):
        _lib.check(self.lib.pp_scatter_mapped(_ptr(self.feat), _ptr(self.pillar_map), self.U + 1, 1, self.D, self.H,
                                              self.W, _ptr(canvas), _sp(stream)))

    def run(self, points, canvas, stream=None, fused=None):
        """fused: None / True = one call (pp_voxelize_scatter) when the shapes allow it; "features" = gather fused with the
        PFN + the stand-alone canvas kernel; False = the three stand-alone calls."""
        stream = stream or torch.cuda.current_stream()
        mode = (True if self.fused else False) if fused is None else fused
        if mode is True and canvas.data_ptr() % 32 == 0 and canvas.numel() % 8 == 0:
            self.voxelize_scatter(points, canvas, stream)
        elif mode:
            self.voxelize_features(points, stream)
            self.scatter(canvas, stream)
        else:
            self.voxelize(points, stream)
            self.encode_scatter(canvas, stream)
        return canvas

    # algorithmic bytes per frame: each boundary input read once, each output written once
    def algorithmic_bytes(self, n_points, m_pillars):
        vox = n_points * self.C * 4 + m_pillars * self.P * self.C * 4 + m_pillars * 3 * 4 + m_pillars * 4
        enc = m_pillars * self.P * self.C * 4 + m_pillars * (4 + 1) * 8 + (self.U + 1) * self.D * self.H * self.W * 4
        return vox, enc


class NmsStage:
    """One class of multiclass_nms with preallocated buffers: boxes (N,9), scores (N,ncls) on device."""

    def __init__(self, n_boxes, device=None, iou_mode=_lib.NMS_AABB2D):
        self.lib = _lib.load()
        self.iou_mode = int(iou_mode)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n = int(n_boxes)
        self.keep = torch.empty((max(self.n, 1),), dtype=torch.int64, device=self.device)
        self.count = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self.ws_bytes = int(self.lib.pp_nms_workspace_bytes_mode(self.n, self.iou_mode))
        self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=self.device)

    def run(self, boxes, scores, score_thr, iou_thr, cls_index=0, stream=None):
        stream = stream or torch.cuda.current_stream()
        sc = ctypes.c_void_p(scores.data_ptr() + 4 * cls_index)
        _lib.check(self.lib.pp_nms_mode(_ptr(boxes), sc, scores.shape[1], boxes.shape[0], float(np.float32(score_thr)),
                                        float(np.float32(iou_thr)), self.iou_mode, _ptr(self.keep), _ptr(self.count),
                                        _ptr(self.ws), self.ws_bytes, _sp(stream)))
        return self.keep, self.count

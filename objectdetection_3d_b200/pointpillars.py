"""Drop-in modules for the pre/post-processing parts of the reference's ``model/PointPillars.py``:
PointPillarsVoxelization, PFNLayer, PillarFeatureNet, the dense scatter of SparseMiddleExtractor and the
box post-processing / target assignment of Anchor3DHead.  Same constructor arguments, parameter names
(state_dict compatible) and return conventions; the arithmetic runs in libpp_b200.

The CNN backbone and the 1x1 conv heads stay stock PyTorch (BASELINE.json north_star).
"""
import ctypes

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import _lib
from .model_utils import Anchor3DRangeGenerator, BBoxCoder, assign_overlaps, limit_period, multiclass_nms
from .ops_numba import VoxelGenerator, _ptr, _stream, voxel_cfg, voxelize_device
from .ops_torch import (bbox2corners3D, bbox2rotated_corners2D, bbox_iou2D, box3d_overlap, check_coplanar,
                        check_nonzero)


class PointPillarsVoxelization(nn.Module):
    """model/PointPillars.py:304-354.  forward(points) -> (voxels f32 [M,P,C], coords int64 [M,3] zyx,
    num_points int64 [M]) on ``device``.  points: numpy (N,C) float32 (the reference's contract) or a
    CUDA tensor."""

    def __init__(self, device, voxel_size, point_cloud_range, max_voxel_points, max_voxels, exact_ties=False):
        super().__init__()
        self.point_cloud_range = np.array(point_cloud_range)
        self.voxel_size = np.array(voxel_size)
        self.max_voxel_points = max_voxel_points
        self.max_voxels = max_voxels
        self.device = device
        self.exact_ties = exact_ties          # reproduce numba's order among equal reflectances (host argsort per call)

    def forward_device(self, points):
        """Device-resident variant: (voxels, coors xyz int32, num int32, voxel_num scalar) without the
        host synchronisation or the int64 / zyx conversions; what the fused pipeline consumes."""
        vs = np.array(self.voxel_size, dtype=np.float32)                 # VoxelGenerator.__init__ (ops_numba.py:48)
        cfg = voxel_cfg(np.float32, vs, self.point_cloud_range, self.max_voxel_points, self.max_voxels, points.shape[1])
        return voxelize_device(points, cfg, _lib.ORDER_REFLECTANCE_DESC)

    def forward(self, points):
        vg = VoxelGenerator(self.voxel_size, self.point_cloud_range, self.max_voxel_points, self.max_voxels)
        if isinstance(points, np.ndarray):
            points = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).to(self.device)
        voxels, coords, num_points = vg.generate(points, self.max_voxels, self.point_cloud_range, True,
                                                 exact_ties=self.exact_ties)
        out_coords = coords[:, [2, 1, 0]].to(torch.int64)
        return voxels, out_coords, num_points.to(torch.int64)


class PFNLayer(nn.Module):
    """model/PointPillars.py:357-423 (parameters: linear.weight, norm.*)."""

    def __init__(self, in_channels, out_channels, last_layer=False, mode='avg'):
        super().__init__()
        self.name = 'PFNLayer'
        self.last_vfe = last_layer
        if not self.last_vfe:
            out_channels = out_channels // 2
        self.units = out_channels
        self.norm = nn.BatchNorm1d(self.units, eps=1e-3, momentum=0.01)
        self.linear = nn.Linear(in_channels, self.units, bias=False)
        assert mode in ['max', 'avg']
        self.mode = mode

    def folded(self):
        """BatchNorm running statistics folded to (scale, shift): y = x * scale + shift."""
        n = self.norm
        scale = n.weight / torch.sqrt(n.running_var + n.eps)
        shift = n.bias - n.running_mean * scale
        return self.linear.weight.contiguous().float(), scale.contiguous().float(), shift.contiguous().float()

    def _wants_grad(self, *tensors):
        """The CUDA kernels are forward-only: with autograd recording and something to differentiate (frozen-BatchNorm
        fine-tuning, saliency, a validation pass that backpropagates) the torch ops run instead, like the reference's."""
        return torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or
                                            any(t is not None and t.requires_grad for t in tensors))

    def forward(self, inputs, num_voxel_points=None, aligned_distance=None):
        if (self.training or self.mode != 'max' or aligned_distance is not None or not inputs.is_cuda
                or self._wants_grad(inputs)):
            return self._forward_autograd(inputs, num_voxel_points, aligned_distance)
        M, P, Cin = inputs.shape
        w, scale, shift = self.folded()
        x = inputs.contiguous().float()
        out = torch.empty((M, 1, self.units) if self.last_vfe else (M, P, 2 * self.units), dtype=torch.float32,
                          device=x.device)
        _lib.check(_lib.load().pp_pfn_layer(_ptr(x), M, P, Cin, _ptr(w), _ptr(scale), _ptr(shift), self.units,
                                            int(self.last_vfe), _ptr(out), _stream()))
        return out

    def _forward_autograd(self, inputs, num_voxel_points, aligned_distance):
        # training path: batch statistics and gradients need autograd; plain torch ops
        x = self.linear(inputs)
        x = self.norm(x.transpose(1, 2)).transpose(1, 2)
        x = F.relu(x)
        if aligned_distance is not None:
            x = x * aligned_distance.unsqueeze(-1)
        if self.mode == 'max':
            pooled = x.max(dim=1, keepdim=True)[0]
        else:
            pooled = x.sum(dim=1, keepdim=True) / num_voxel_points.type_as(inputs).view(-1, 1, 1)
        if self.last_vfe:
            return pooled
        return torch.cat([x, pooled.expand(-1, inputs.shape[1], -1)], dim=2)


def _coors_kind(coors):
    if coors.dtype == torch.int64:
        return _lib.COORS_BZYX_I64
    if coors.dtype == torch.int32:
        return _lib.COORS_BZYX_I32
    raise ValueError("coors must be int32 or int64 (b, z, y, x)")


def _num_kind(num):
    if num.dtype == torch.int64:
        return _lib.NUM_I64
    if num.dtype == torch.int32:
        return _lib.NUM_I32
    raise ValueError("num_points must be int32 or int64")


class PillarFeatureNet(nn.Module):
    """model/PointPillars.py:426-526.  forward(features (M,P,C), num_points (M,), coors (M,4) b,z,y,x)
    -> (M, feat_channels[-1]); the last channel is num_points (:526)."""

    def __init__(self, in_channels, feat_channels, voxel_size, point_cloud_range):
        super().__init__()
        assert len(feat_channels) > 0
        self.raw_channels = in_channels
        in_channels += 5
        self.in_channels = in_channels
        chans = [in_channels] + list(feat_channels)
        layers = []
        for i in range(len(chans) - 1):
            last = i == len(chans) - 2
            layers.append(PFNLayer(chans[i], chans[i + 1] - 1 if last else chans[i + 1], last_layer=last, mode='max'))
        self.pfn_layers = nn.ModuleList(layers)
        self.point_cloud_range = point_cloud_range
        self.vx = voxel_size[0]
        self.vy = voxel_size[1]
        self.x_offset = self.vx / 2 + self.point_cloud_range[0]
        self.y_offset = self.vy / 2 + self.point_cloud_range[1]

    def decorate(self, features, num_points, coors, m_dev=None):
        """(M,P,C) -> (M,P,C+5): :490-521 in one kernel."""
        M, P, C = features.shape
        x = features.contiguous().float()
        out = torch.empty((M, P, C + 5), dtype=torch.float32, device=x.device)
        num_points, coors = num_points.contiguous(), coors.contiguous()
        _lib.check(_lib.load().pp_decorate(_ptr(x), _ptr(num_points), _num_kind(num_points), _ptr(coors),
                                           _coors_kind(coors), M, _ptr(m_dev), P, C, float(self.vx), float(self.vy),
                                           float(self.x_offset), float(self.y_offset), _ptr(out), _stream()))
        return out

    def _decorate_autograd(self, features, num_points, coors):
        """:490-521 as torch ops (only when the points themselves need gradients)."""
        P = features.shape[1]
        mean = features[:, :, :3].sum(dim=1, keepdim=True) / num_points.type_as(features).view(-1, 1, 1)
        f_cluster = features[:, :, :3] - mean
        cx = coors[:, 3].to(features.dtype).unsqueeze(1) * self.vx + self.x_offset
        cy = coors[:, 2].to(features.dtype).unsqueeze(1) * self.vy + self.y_offset
        f_center = torch.stack([features[:, :, 0] - cx, features[:, :, 1] - cy], dim=-1)
        x = torch.cat([features, f_cluster, f_center], dim=-1)
        mask = (num_points.view(-1, 1) > torch.arange(P, device=features.device).view(1, -1)).unsqueeze(-1)
        return x * mask.type_as(x)

    def forward(self, features, num_points, coors):
        if not features.is_cuda:
            raise _lib.PPError("PillarFeatureNet: CUDA tensors required (no CPU fallback)")
        M, P, C = features.shape
        single = len(self.pfn_layers) == 1
        wants_grad = torch.is_grad_enabled() and (features.requires_grad or any(p.requires_grad for p in self.parameters()))
        if features.requires_grad and torch.is_grad_enabled():
            x = self._decorate_autograd(features, num_points, coors)       # gradients w.r.t. the points: torch ops
            for pfn in self.pfn_layers:
                x = pfn(x, num_points)
            return torch.cat((x.squeeze(1), num_points.view(-1, 1).to(x.dtype)), dim=-1)
        if single and not self.training and not wants_grad:
            layer = self.pfn_layers[0]
            w, scale, shift = layer.folded()
            x = features.contiguous().float()
            num_points, coors = num_points.contiguous(), coors.contiguous()
            out = torch.empty((M, layer.units + 1), dtype=torch.float32, device=x.device)
            _lib.check(_lib.load().pp_pillar_features(
                _ptr(x), _ptr(num_points), _num_kind(num_points), _ptr(coors), _coors_kind(coors), M, None, P, C,
                float(self.vx), float(self.vy), float(self.x_offset), float(self.y_offset), _ptr(w), _ptr(scale),
                _ptr(shift), layer.units, _ptr(out), _stream()))
            return out
        x = self.decorate(features, num_points, coors)
        for pfn in self.pfn_layers:
            x = pfn(x, num_points)
        return torch.cat((x.squeeze(1), num_points.view(-1, 1).to(x.dtype)), dim=-1)


def head_topk(scores, k):
    """Indices of the k largest of a 1-D CUDA float tensor, descending score, equal scores by lower index
    (`scores.topk(k)[1]` of model/PointPillars.py:1058-1059 with a deterministic tie rule), on pp_head_topk."""
    lib = _lib.load()
    scores = scores.contiguous().float()
    n = scores.numel()
    k = min(int(k), n)
    rows = torch.empty((k,), dtype=torch.int64, device=scores.device)
    ws_bytes = int(lib.pp_head_topk_workspace_bytes(n, k))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=scores.device)
    _lib.check(lib.pp_head_topk(_ptr(scores), n, k, _ptr(rows), _ptr(ws), ws_bytes, _stream()))
    return rows


class _DenseScatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, coors, batch_size, D, H, W):
        lib = _lib.load()
        feat = feat.contiguous().float()
        coors = coors.contiguous()
        M, C = feat.shape
        canvas = torch.empty((batch_size, C * D, H, W), dtype=torch.float32, device=feat.device)
        ws_bytes = int(lib.pp_scatter_workspace_bytes(batch_size, D, H, W))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=feat.device)
        _lib.check(lib.pp_scatter_dense(_ptr(feat), _ptr(coors), _coors_kind(coors), M, None, C, 0, batch_size, D, H, W,
                                        _ptr(canvas), _ptr(ws), ws_bytes, _stream()))
        ctx.save_for_backward(coors)
        ctx.dims = (C, D, H, W)
        return canvas

    @staticmethod
    def backward(ctx, grad):
        (coors,) = ctx.saved_tensors
        C, D, H, W = ctx.dims
        c = coors.long()
        g = grad.reshape(grad.shape[0], C, D, H, W)[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]]
        return g, None, None, None, None, None


class SparseMiddleExtractor(nn.Module):
    """The dense-scatter semantics of model/PointPillars.py:529-573:
    SparseConvTensor(features, coors, sparse_shape, batch).dense().view(N, C*D, H, W).
    The reference's sparse 3-D convolutions (spconv, absent) belong to the CNN backbone and are not part
    of this path; ``output_shape`` is (D, H, W)."""

    def __init__(self, output_shape, in_channels=None, out_channels=None):
        super().__init__()
        self.sparse_shape = [int(round(float(v))) for v in output_shape]
        self.in_channels = in_channels

    def forward(self, voxel_features, coors, batch_size):
        D, H, W = self.sparse_shape
        return _DenseScatter.apply(voxel_features, coors, int(batch_size), D, H, W)


def dense_to_sparse(x):
    """The dense -> sparse step of SubmanifoldSparseRPN.forward, model/PointPillars.py:766-789: the cells of x
    (B,C,H,W) with any non-zero channel, in (b, y, x) row-major order.  Returns (values (nnz,C) f32, coords (nnz,3)
    int32 = [b, y, x]), i.e. the arguments of spconv.SparseConvTensor(values, coords, x.shape[-2:], x.shape[0])."""
    lib = _lib.load()
    x = x.detach().float().contiguous()
    B, C, H, W = (int(v) for v in x.shape)
    cells = B * H * W
    coords = torch.empty((cells, 3), dtype=torch.int32, device=x.device)
    values = torch.empty((cells, C), dtype=torch.float32, device=x.device)
    nnz = torch.zeros((1,), dtype=torch.int32, device=x.device)
    ws_bytes = int(lib.pp_compact_workspace_bytes(cells))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
    _lib.check(lib.pp_dense_to_sparse(_ptr(x), B, C, H, W, _ptr(coords), _ptr(values), _ptr(nnz), _ptr(ws), ws_bytes,
                                      _stream()))
    k = int(nnz.item())
    return values[:k], coords[:k]


class Anchor3DHead(nn.Module):
    """model/PointPillars.py:795-1094.  The 1x1 conv heads are stock torch; get_bboxes / assign_bboxes
    run the box decode, BEV IoU, NMS and encode kernels of libpp_b200 (nms_dim == 2 form)."""

    def __init__(self, num_classes=1, in_channels=384, nms_dim=2, nms_pre=100, nms_thresh=0.7, score_thr=0.1,
                 box_params_num=9, dir_offset=0, ranges=[], sizes=[], rotations=[], iou_thr=[]):
        super().__init__()
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.nms_pre = nms_pre
        self.nms_thresh = nms_thresh
        self.score_thr = score_thr
        self.dir_offset = dir_offset
        self.iou_thr = list(iou_thr)
        self.sizes, self.ranges, self.rotations = sizes, ranges, rotations
        self.box_params_num = box_params_num
        self.nms_dim = nms_dim
        if len(self.iou_thr) != num_classes:
            assert len(self.iou_thr) == 1
            self.iou_thr = self.iou_thr * num_classes
        self.anchor_generator = Anchor3DRangeGenerator(ranges=ranges, sizes=sizes, rotations=rotations,
                                                       box_params_num=box_params_num)
        self.num_anchors = self.anchor_generator.num_base_anchors
        self.bbox_coder = BBoxCoder()
        self.cls_out_channels = self.num_anchors * self.num_classes
        self.conv_cls = nn.Conv2d(in_channels, self.cls_out_channels, 1)
        self.conv_reg = nn.Conv2d(in_channels, self.num_anchors * box_params_num, 1)
        self.conv_dir_cls_x = nn.Conv2d(in_channels, self.num_anchors * 2, 1)
        self.conv_dir_cls_y = nn.Conv2d(in_channels, self.num_anchors * 2, 1)
        self.conv_dir_cls_z = nn.Conv2d(in_channels, self.num_anchors * 2, 1)
        nn.init.normal_(self.conv_cls.weight, 0, 0.01)
        nn.init.constant_(self.conv_cls.bias, float(-np.log((1 - 0.01) / 0.01)))
        nn.init.normal_(self.conv_reg.weight, 0, 0.01)
        nn.init.constant_(self.conv_reg.bias, 0)

    def forward(self, x):
        dirs = [self.conv_dir_cls_x(x), self.conv_dir_cls_y(x), self.conv_dir_cls_z(x)]
        return self.conv_cls(x), self.conv_reg(x), torch.cat(dirs, dim=1)

    # ---- inference: :1002-1094 ------------------------------------------------------------------
    def get_bboxes(self, cls_scores, bbox_preds, dir_preds):
        out = [self.get_bboxes_single(c, b, d) for c, b, d in zip(cls_scores, bbox_preds, dir_preds)]
        return [o[0] for o in out], [o[1] for o in out], [o[2] for o in out]

    def get_bboxes_single(self, cls_scores, bbox_preds, dir_preds):
        """:1025-1094 on the fused kernels: per-anchor score straight from the conv layout, top-k, then anchors
        generated on the fly + gather + decode for the nms_pre survivors only, NMS, fused direction fix-up."""
        assert cls_scores.size()[-2:] == bbox_preds.size()[-2:]
        assert cls_scores.size()[-2:] == dir_preds.size()[-2:]
        lib = _lib.load()
        H, W = (int(v) for v in cls_scores.shape[-2:])
        A, ncls = self.num_anchors, self.num_classes
        dev = cls_scores.device
        cls = cls_scores.detach().to(torch.float32).contiguous()
        reg = bbox_preds.detach().to(torch.float32).contiguous()
        dirs = dir_preds.detach().to(torch.float32).contiguous()
        assert cls.shape[0] == A * ncls and reg.shape[0] == A * self.box_params_num and dirs.shape[0] == A * 6
        total = H * W * A
        rows, K = None, total
        if total > self.nms_pre:
            max_scores = torch.empty((total,), dtype=torch.float32, device=dev)
            _lib.check(lib.pp_head_max_scores(_ptr(cls), A, ncls, H, W, _ptr(max_scores), _stream()))
            rows = head_topk(max_scores, self.nms_pre)                   # :1058-1059
            K = self.nms_pre
        gen = self.anchor_generator
        assert len(gen.ranges) == 1, "one anchor range (as in config.yaml:64)"
        fp = ctypes.POINTER(ctypes.c_float)
        rg = np.asarray(gen.ranges[0], dtype=np.float32)
        sizes = np.ascontiguousarray(np.asarray(gen.sizes, dtype=np.float32).reshape(-1, 3))
        rots = np.ascontiguousarray(np.asarray(gen.rotations, dtype=np.float32).reshape(-1, 3))
        bboxes = torch.empty((K, 9), dtype=torch.float32, device=dev)
        scores = torch.empty((K, ncls), dtype=torch.float32, device=dev)
        dir_bits = torch.empty((K, 3), dtype=torch.int32, device=dev)
        _lib.check(lib.pp_head_select_decode(_ptr(cls), _ptr(reg), _ptr(dirs), _ptr(rows) if rows is not None else None, K,
                                             rg.ctypes.data_as(fp), sizes.ctypes.data_as(fp), sizes.shape[0],
                                             rots.ctypes.data_as(fp), rots.shape[0], ncls, H, W, _ptr(bboxes), _ptr(scores),
                                             _ptr(dir_bits), _stream()))
        idxs = multiclass_nms(bboxes, scores, self.score_thr, self.nms_thresh, self.nms_dim)
        labels = torch.cat([torch.full((len(idxs[i]),), i, dtype=torch.long) for i in range(ncls)])
        out_scores = torch.cat([scores[idxs[i], i] for i in range(ncls)])
        idxs = torch.cat(idxs)
        bboxes = bboxes[idxs].contiguous()
        dir_bits = dir_bits[idxs].contiguous()
        _lib.check(lib.pp_head_direction_fixup(_ptr(bboxes), _ptr(dir_bits), bboxes.shape[0], float(self.dir_offset),
                                               _stream()))
        return bboxes, out_scores, labels

    def get_bboxes_single_unfused(self, cls_scores, bbox_preds, dir_preds):
        """The same function in the reference's own order of operations (materialised anchors, permuted views,
        per-op kernels): kept as the in-repo cross-check of the fused path (tests/test_gpu_parity.py)."""
        assert cls_scores.size()[-2:] == bbox_preds.size()[-2:]
        assert cls_scores.size()[-2:] == dir_preds.size()[-2:]
        anchors = self.anchor_generator.grid_anchors(cls_scores.shape[-2:], device=cls_scores.device)
        anchors = anchors.reshape(-1, self.box_params_num)
        dir_preds = dir_preds.permute(1, 2, 0).reshape(-1, 6)
        dir_bits = torch.stack([dir_preds[:, 0:2].max(dim=-1)[1], dir_preds[:, 2:4].max(dim=-1)[1],
                                dir_preds[:, 4:6].max(dim=-1)[1]], dim=1)
        scores = cls_scores.permute(1, 2, 0).reshape(-1, self.num_classes).sigmoid()
        bbox_preds = bbox_preds.permute(1, 2, 0).reshape(-1, self.box_params_num)
        if scores.shape[0] > self.nms_pre:
            # decode is element-wise, so selecting first and decoding the nms_pre survivors gives the
            # same boxes as the reference's decode-all-then-select (:1053-1065)
            _, topk = scores.max(dim=1)[0].topk(self.nms_pre)
            anchors, bbox_preds, scores, dir_bits = anchors[topk], bbox_preds[topk], scores[topk], dir_bits[topk]
        bboxes = self.bbox_coder.decode(anchors, bbox_preds.contiguous())
        scores = scores.contiguous()
        idxs = multiclass_nms(bboxes, scores, self.score_thr, self.nms_thresh, self.nms_dim)
        labels = torch.cat([torch.full((len(idxs[i]),), i, dtype=torch.long) for i in range(self.num_classes)])
        out_scores = torch.cat([scores[idxs[i], i] for i in range(self.num_classes)])
        idxs = torch.cat(idxs)
        bboxes = bboxes[idxs]
        dir_bits = dir_bits[idxs]
        if bboxes.shape[0] > 0:
            for k in range(3):
                col = bboxes[:, 6 + k].contiguous()
                rot = limit_period(col - self.dir_offset, 1, np.pi)
                bboxes[:, 6 + k] = rot + self.dir_offset + np.pi * dir_bits[:, k].to(bboxes.dtype)
        return bboxes, out_scores, labels

    # ---- training targets: :886-1000 ------------------------------------------------------------
    def assign_bboxes(self, pred_bboxes, target_bboxes):
        dev = pred_bboxes.device
        anchors = self.anchor_generator.grid_anchors(pred_bboxes.shape[-2:], device=dev)
        anchors_cnt = int(np.prod(anchors.shape[:-1]))
        rot_angles = anchors.shape[-2]
        flat = anchors.reshape(-1, self.box_params_num)
        box2vertices = bbox2corners3D if self.nms_dim == 3 else bbox2rotated_corners2D      # :899-905
        anchor_rect = box2vertices(flat)
        if self.nms_dim == 3:
            check_coplanar(anchor_rect, 1e-2)
            check_nonzero(anchor_rect, 1e-2)
        assigned, target_idxs, pos_idxs, neg_idxs = [], [], [], []

        def flatten_idx(idx, j):
            z = torch.div(idx, rot_angles, rounding_mode='trunc')
            return z * self.num_classes * rot_angles + j * rot_angles + idx % rot_angles

        idx_off = 0
        for i, gts in enumerate(target_bboxes):
            for j in range(self.num_classes):
                if gts.shape[0] == 0:
                    assigned.append(torch.zeros((0, self.box_params_num), device=dev))
                    for lst in (target_idxs, pos_idxs, neg_idxs):
                        lst.append(torch.zeros((0,), dtype=torch.long, device=dev))
                    continue
                lo, hi = self.iou_thr[j]
                gt_vertices = box2vertices(gts)
                if self.nms_dim == 3:                      # box3d_overlap's validity checks (ops_torch.py:743-748)
                    check_coplanar(gt_vertices, 1e-2)
                    check_nonzero(gt_vertices, 1e-2)
                # IoU (:964-965) fused with both maxima (:968-971) and the low-quality matching (:976-978)
                max_ov, argmax_ov, _, lowq = assign_overlaps(gt_vertices, anchor_rect, lo, self.nms_dim)
                pos = (max_ov >= hi) | lowq
                neg = (max_ov >= 0) & (max_ov < lo)
                assigned.append(self.bbox_coder.encode(flat[pos], gts[argmax_ov[pos]]))
                target_idxs.append(argmax_ov[pos] + idx_off)
                pos_idxs.append(flatten_idx(pos.nonzero(as_tuple=False).squeeze(-1), j) + i * anchors_cnt)
                neg_idxs.append(flatten_idx(neg.nonzero(as_tuple=False).squeeze(-1), j) + i * anchors_cnt)
            idx_off += len(gts)
        return (torch.cat(assigned, dim=0), torch.cat(target_idxs, dim=0), torch.cat(pos_idxs, dim=0),
                torch.cat(neg_idxs, dim=0))

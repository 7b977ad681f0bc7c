"""The training step of BASELINE.json configs[4] around the hot path: what
``ObjectDetection.run_training`` does per batch (pipeline/pipeline.py:485-499) with ``PointPillars.loss``
(model/PointPillars.py:147-226), restated so that it can be benchmarked without the reference's control plane.

On the hot path (this library's kernels): per-frame voxelization, pillar decoration, dense scatter (+ its backward),
target assignment (IoU + both maxima + low-quality matches, no G x A matrix) and box encoding.  Stock PyTorch, by
design: the PFN layers in training mode (batch statistics, autograd), a dense 2-D backbone STAND-IN (the reference's
sparse-conv RPN needs spconv, which is out of scope), the 1x1 conv heads, the three losses (plain torch restatements
of losses/focal_loss.py, losses/smooth_L1.py, losses/cross_entropy.py), AdamW and DistributedDataParallel's NCCL
gradient all-reduce -- the one exchange step of the path.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .model_utils import limit_period
from .pointpillars import Anchor3DHead, PillarFeatureNet, PointPillarsVoxelization, SparseMiddleExtractor


def focal_loss(pred, target, avg_factor, gamma=2.0, alpha=0.25, loss_weight=1.0):
    """losses/focal_loss.py:33-52 (sigmoid focal loss on one-hot targets; label == num_classes is background)."""
    pred_sigmoid = pred.sigmoid()
    target = (target.unsqueeze(-1) == torch.arange(pred.shape[-1], device=pred.device).unsqueeze(0)).type_as(pred)
    pt = (1 - pred_sigmoid) * target + pred_sigmoid * (1 - target)
    focal_weight = (alpha * target + (1 - alpha) * (1 - target)) * pt.pow(gamma)
    loss = F.binary_cross_entropy_with_logits(pred, target, reduction="none") * focal_weight * loss_weight
    if avg_factor is None:
        return loss.mean()
    return loss.sum() / avg_factor if avg_factor > 0 else loss


def smooth_l1_loss(pred, target, avg_factor, beta=0.11, loss_weight=2.0):
    """losses/smooth_L1.py:36-47 with config.yaml:26-28."""
    diff = torch.abs(pred - target)
    loss = torch.where(diff < beta, 0.5 * diff * diff / beta, diff - 0.5 * beta) * loss_weight
    return loss.sum() / avg_factor if avg_factor else loss.mean()


def cross_entropy_loss(cls_score, label, avg_factor, loss_weight=0.2):
    """losses/cross_entropy.py:37-45 with config.yaml:29-30."""
    loss = F.cross_entropy(cls_score, label, reduction="none") * loss_weight
    return loss.sum() / avg_factor if avg_factor else loss.mean()


def pointpillars_loss(head, results, gt_bboxes, gt_labels):
    """PointPillars.loss, model/PointPillars.py:147-226: assign_bboxes (kernels), then the three losses."""
    scores, bboxes, dirs = results
    target_bboxes, target_idx, pos_idx, neg_idx = head.assign_bboxes(bboxes, gt_bboxes)
    avg_factor = pos_idx.size(0)
    scores = scores.permute((0, 2, 3, 1)).reshape(-1, head.num_classes)
    target_labels = torch.full((scores.size(0),), head.num_classes, device=scores.device, dtype=gt_labels[0].dtype)
    target_labels[pos_idx] = torch.cat(gt_labels, dim=0)[target_idx]
    sel = torch.cat([pos_idx, neg_idx], dim=0)
    loss_cls = focal_loss(scores[sel], target_labels[sel], avg_factor)
    cond = (target_labels[pos_idx] >= 0) & (target_labels[pos_idx] < head.num_classes)
    pos_idx, target_idx, target_bboxes = pos_idx[cond], target_idx[cond], target_bboxes[cond]
    bboxes = bboxes.permute((0, 2, 3, 1)).reshape(-1, head.box_params_num)[pos_idx]
    dirs = dirs.permute((0, 2, 3, 1)).reshape(-1, 6)[pos_idx]
    if len(pos_idx) > 0:
        gt = torch.cat(gt_bboxes, dim=0)[target_idx]
        loss_dirs = []
        for k in range(3):                                   # direction bins of rx, ry, rz (:181-196)
            t = limit_period(gt[:, -3 + k].contiguous(), 0, 2 * np.pi)
            t = (t / np.pi).long() % 2
            loss_dirs.append(cross_entropy_loss(dirs[:, 2 * k:2 * k + 2], t, avg_factor))
        r0 = torch.sin(bboxes[:, -3:]) * torch.cos(target_bboxes[:, -3:])      # sine-difference transform (:200-206)
        r1 = torch.cos(bboxes[:, -3:]) * torch.sin(target_bboxes[:, -3:])
        loss_bbox = smooth_l1_loss(torch.cat([bboxes[:, :-3], r0], dim=-1), torch.cat([target_bboxes[:, :-3], r1], dim=-1),
                                   avg_factor)
    else:
        loss_cls, loss_bbox = loss_cls.sum(), bboxes.sum()
        loss_dirs = [dirs[:, 2 * k:2 * k + 2].sum() for k in range(3)]
    return {"loss_cls": loss_cls, "loss_bbox": loss_bbox, "loss_dir_x": loss_dirs[0], "loss_dir_y": loss_dirs[1],
            "loss_dir_z": loss_dirs[2]}


class DenseBackboneStandIn(nn.Module):
    """Stock dense 2-D backbone + upsampling neck of the PointPillars paper (three stride-2 stages of 3x3 conv-BN-ReLU,
    each upsampled to the first stage's resolution and concatenated).  A stand-in for the CNN the north star leaves on
    stock PyTorch: it is NOT part of the product and NOT the reference's sparse-conv RPN."""

    def __init__(self, in_channels=64, channels=(64, 128, 256), layers=(3, 5, 5), up_channels=128):
        super().__init__()
        self.blocks, self.ups = nn.ModuleList(), nn.ModuleList()
        c_in = in_channels
        for i, (c, n) in enumerate(zip(channels, layers)):
            seq = [nn.Conv2d(c_in, c, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(c, eps=1e-3, momentum=0.01), nn.ReLU(inplace=True)]
            for _ in range(n):
                seq += [nn.Conv2d(c, c, 3, padding=1, bias=False), nn.BatchNorm2d(c, eps=1e-3, momentum=0.01), nn.ReLU(inplace=True)]
            self.blocks.append(nn.Sequential(*seq))
            s = 2 ** i
            self.ups.append(nn.Sequential(nn.ConvTranspose2d(c, up_channels, s, stride=s, bias=False),
                                          nn.BatchNorm2d(up_channels, eps=1e-3, momentum=0.01), nn.ReLU(inplace=True)))
            c_in = c
        self.out_channels = up_channels * len(channels)

    def forward(self, x):
        outs = []
        for blk, up in zip(self.blocks, self.ups):
            x = blk(x)
            outs.append(up(x))
        return torch.cat(outs, dim=1)


class TrainableNet(nn.Module):
    """Everything with parameters (what DistributedDataParallel wraps): PFN -> scatter -> backbone -> head convs."""

    def __init__(self, geom, sizes, rotations, iou_thr, pfn_out=64, grid_hw=(496, 432)):
        super().__init__()
        vs, rg = geom["voxel_size"], geom["point_cloud_range"]
        self.voxel_encoder = PillarFeatureNet(4, [pfn_out], vs, rg)
        self.pseudoimage_generator = SparseMiddleExtractor([1, grid_hw[0], grid_hw[1]])
        self.backbone = DenseBackboneStandIn(pfn_out)
        self.bbox_head = Anchor3DHead(num_classes=1, in_channels=self.backbone.out_channels, nms_dim=2, nms_pre=500,
                                      nms_thresh=1e-5, score_thr=0.3, ranges=[list(rg)], sizes=sizes, rotations=rotations,
                                      iou_thr=iou_thr)

    def forward(self, voxels, num_points, coors, batch_size):
        feats = self.voxel_encoder(voxels, num_points, coors)
        x = self.pseudoimage_generator(feats, coors, batch_size)
        return self.bbox_head(self.backbone(x))


class TrainStep:
    """One optimisation step as pipeline/pipeline.py:485-499 runs it: forward, loss, zero_grad, backward,
    clip_grad_value_(2) (config.yaml:108), AdamW(lr 1e-4, betas (0.95, 0.99), weight decay 0.01; config.yaml:113-116)."""

    def __init__(self, net, geom, device, ddp=None):
        self.net = net                                   # the bare module (losses need its head)
        self.model = ddp if ddp is not None else net     # what is called (DDP hooks the gradient all-reduce into backward)
        self.voxel_layer = PointPillarsVoxelization(device, geom["voxel_size"], geom["point_cloud_range"],
                                                    geom["max_voxel_points"], geom["max_voxels"])
        self.optimizer = torch.optim.AdamW(net.parameters(), lr=1e-4, betas=(0.95, 0.99), weight_decay=0.01)
        self.grad_clip = 2

    @torch.no_grad()
    def voxelize(self, points):
        """PointPillars.voxelize, model/PointPillars.py:106-134."""
        voxels, coors, num_points = [], [], []
        for pc in points:
            v, c, n = self.voxel_layer(pc)
            voxels.append(v); coors.append(c); num_points.append(n)
        coors = torch.cat([F.pad(c, (1, 0), mode="constant", value=i) for i, c in enumerate(coors)], dim=0)
        return torch.cat(voxels, dim=0), torch.cat(num_points, dim=0), coors

    def __call__(self, points, gt_bboxes, gt_labels, marks=None):
        mark = marks.append if marks is not None else (lambda _: None)
        mark("start")
        voxels, num_points, coors = self.voxelize(points)
        mark("voxelize")
        results = self.model(voxels, num_points, coors, len(points))
        mark("forward")
        loss = pointpillars_loss(self.net.bbox_head, results, gt_bboxes, gt_labels)
        loss_sum = sum(loss.values())
        mark("assign+loss")
        self.optimizer.zero_grad()
        loss_sum.backward()
        mark("backward(+allreduce)")
        torch.nn.utils.clip_grad_value_(self.net.parameters(), self.grad_clip)
        self.optimizer.step()
        mark("clip+adamw")
        return loss_sum.detach()

"""Launch every kernel family of libpp_b200 once at benchmark sizes (for ncu --set full / compute-sanitizer).
usage: run_all_kernels.py [small]     (small: sizes for compute-sanitizer, which slows kernels ~50x)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from objectdetection_3d_b200 import _lib, model_utils, ops_numpy, ops_torch, pipeline, pointpillars, synth

small = "small" in sys.argv
dev = torch.device("cuda", 0)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
g, pfn = synth.G_KITTI, synth.pfn_params(9, 63, seed=5)
n_pts = 60_000 if small else 1_000_000
n_box = 3000 if small else 20_000
st = torch.cuda.current_stream()

# ---- stage 1 + 2: the one-call frame, the unfused calls, both key widths, hash addressing (G_ref), any max_points
pts = cu(synth.dense_tile(n=n_pts, seed=99, n_cells=3000 if small else 11_500, n_clusters=80 if small else 600))
for order in (_lib.ORDER_REFLECTANCE_DESC, _lib.ORDER_GIVEN):
    pipe = pipeline.FramePipeline(g, pfn, n_pts, order=order)
    canvas = pipe.new_canvas()
    pipe.run(pts, canvas, st)                       # pp_voxelize_scatter
    pipe.run(pts, canvas, st, fused="features")     # pp_voxelize_features + canvas kernel
    pipe.run(pts, canvas, st, fused=False)          # pp_voxelize + pp_pillar_features + pp_scatter_mapped
f120 = synth.forest_tile(n=30_000 if small else 120_000)
vs, rg = np.array(synth.G_REF["voxel_size"], np.float32), np.array(synth.G_REF["point_cloud_range"], np.float64)
from objectdetection_3d_b200 import ops_numba
ops_numba.points_to_voxel(cu(f120), vs, rg, 50, 7_500_000, True)           # hash addressing, max_points 50
ops_numba.points_to_voxel(cu(f120), np.array(g["voxel_size"], np.float32), np.array(g["point_cloud_range"]), 100, 12000, True)   # any max_points
# decoration alone, generic (two-layer) PFN, scatter with its own map
m = int(pipe.voxel_num.item())
net = pointpillars.PillarFeatureNet(4, [32, 64], g["voxel_size"], g["point_cloud_range"]).to(dev).eval()
c4 = torch.nn.functional.pad(pipe.coors[:m][:, [2, 1, 0]].long(), (1, 0), value=0)
with torch.no_grad():
    feat = net(pipe.voxels[:m], pipe.num[:m].long(), c4)
    pointpillars.SparseMiddleExtractor([1, pipe.H, pipe.W])(feat, c4, 1)
    pointpillars.dense_to_sparse(canvas)
# ---- either side of the path
raw = synth.forest_tile(n=50_000 if small else 200_000, seed=41)
ops_numpy.preprocess_points(cu(raw), [0, 0, 0, 40.0, 40.0, 30.0], [0, 1, 2, 3])
cloud = np.concatenate([np.random.default_rng(1).uniform(0, 8, (50_000, 3)), np.random.default_rng(2).random((50_000, 1))], 1).astype(np.float32)
model_utils.CustomVoxelizer(dict(voxel_size=[0.25, 0.25, 0.25], max_voxel_points=8, reflectance_sampling=True)).voxelize(cloud)
# ---- stage 3: codec, anchors, IoU forms, assignment, head post-processing, NMS in the three pair-test modes
boxes, scores = synth.nms_boxes(n=n_box, seed=4, extent=40.0)
tb, ts = cu(boxes), cu(scores)
gen = model_utils.Anchor3DRangeGenerator([[0, 0, 0, 40.0, 40.0, 30.0]], synth.ANCHOR_SIZES, synth.ANCHOR_ROTATIONS, 9)
anchors = gen.grid_anchors((50, 50) if small else (200, 200), device="cuda").reshape(-1, 9)
gts = tb[:40].contiguous()
ne = min(4000, n_box, anchors.shape[0])
enc = model_utils.BBoxCoder.encode(anchors[:ne].contiguous(), tb[:ne].contiguous())
model_utils.BBoxCoder.decode(anchors[:ne].contiguous(), enc)
model_utils.limit_period(tb[:, 8].contiguous(), 0.5, np.pi)
rect = ops_torch.bbox2rotated_corners2D(tb)
corners = ops_torch.bbox2corners3D(tb[:2000].contiguous())
ops_torch.bbox_iou2D(rect[:2000].contiguous(), rect[2000:3000].contiguous())
ops_numba.iou_jit(rect[:500].contiguous(), rect[500:900].contiguous(), 0.0)
flat = tb.clone(); flat[:, 6:8] = 0
ops_torch.bbox_iou_rotated_bev(flat[:2000].contiguous(), flat[2000:3000].contiguous())
ops_torch.box3d_overlap(corners[:1000].contiguous(), corners[1000:2000].contiguous())
model_utils.assign_overlaps(ops_torch.bbox2rotated_corners2D(gts), ops_torch.bbox2rotated_corners2D(anchors), 0.08, 2)
model_utils.assign_overlaps(ops_torch.bbox2corners3D(gts), ops_torch.bbox2corners3D(anchors[::8].contiguous()), 0.08, 3)
H = W = 50 if small else 200
head = pointpillars.Anchor3DHead(num_classes=1, in_channels=8, nms_pre=500, nms_thresh=0.1, score_thr=0.3,
                                 ranges=[[0, 0, 0, 40.0, 40.0, 30.0]], sizes=synth.ANCHOR_SIZES,
                                 rotations=synth.ANCHOR_ROTATIONS, iou_thr=[[0.08, 0.2]]).to(dev)
rng = np.random.default_rng(5)
with torch.no_grad():
    head.get_bboxes_single(cu(rng.normal(-2, 2, (12, H, W)).astype(np.float32)), cu(rng.normal(0, 0.1, (108, H, W)).astype(np.float32)),
                           cu(rng.normal(0, 1, (72, H, W)).astype(np.float32)))
model_utils.multiclass_nms(tb, ts, 0.0, 0.1, 2)
model_utils.multiclass_nms(flat, ts, 0.0, 0.1, 2, iou_mode="rot_bev")
model_utils.multiclass_nms(tb, ts, 0.0, 0.1, 3)
torch.cuda.synchronize()
print("ok: %d kernel launches" % _lib.launch_count())

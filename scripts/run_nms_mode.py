"""A few NMS calls on 20k boxes with one pair test (for ncu captures).  usage: run_nms_mode.py {aabb2d|rot_bev|box3d} [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_3d_b200 import pipeline, synth, _lib
mode = {"aabb2d": _lib.NMS_AABB2D, "rot_bev": _lib.NMS_ROT_BEV, "box3d": _lib.NMS_BOX3D}[sys.argv[1]]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
b, s = synth.nms_boxes(n=20000, seed=4, extent=40.0)
b, s = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
nms = pipeline.NmsStage(20000, iou_mode=mode)
for _ in range(reps):
    nms.run(b, s, 0.0, 0.1, 0, torch.cuda.current_stream())
torch.cuda.synchronize()
print("ok", sys.argv[1], int(nms.count.item()))

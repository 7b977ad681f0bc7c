import sys, os
sys.path.insert(0, os.getcwd())
import torch
from objectdetection_3d_b200 import pipeline, synth, _lib
b, s = synth.nms_boxes(n=20000, seed=4, extent=40.0)
b, s = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
st = torch.cuda.current_stream()
nms = pipeline.NmsStage(20000, iou_mode=_lib.NMS_AABB2D)
for _ in range(3):
    nms.run(b, s, 0.0, 0.1, 0, st); torch.cuda.synchronize(); print("---")

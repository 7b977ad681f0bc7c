"""Sweep timing prints of the development build (PP_NVCC_EXTRA=-DPP_TIMING, PP_B200_LIB=...libpp_b200_dev.so).  usage: sweep_dbg.py [mode]"""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from objectdetection_3d_b200 import pipeline, synth, _lib
b, s = synth.nms_boxes(n=20000, seed=4, extent=40.0)
b, s = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
st = torch.cuda.current_stream()
mode = {"aabb2d": _lib.NMS_AABB2D, "rot_bev": _lib.NMS_ROT_BEV, "box3d": _lib.NMS_BOX3D}[sys.argv[1] if len(sys.argv) > 1 else "aabb2d"]
nms = pipeline.NmsStage(20000, iou_mode=mode)
for _ in range(3):
    nms.run(b, s, 0.0, 0.1, 0, st); torch.cuda.synchronize(); print("---", flush=True)

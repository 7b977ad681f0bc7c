"""NMS on 20k boxes with the three pair tests: time per call and per-kernel split.  usage: nms_modes.py [extent]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_3d_b200 import pipeline, synth, _lib
ext = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
b, s = synth.nms_boxes(n=20000, seed=4, extent=ext)
b, s = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
st = torch.cuda.current_stream()
for name, mode in (("aabb2d", _lib.NMS_AABB2D), ("rot_bev", _lib.NMS_ROT_BEV), ("box3d", _lib.NMS_BOX3D)):
    nms = pipeline.NmsStage(20000, iou_mode=mode)
    for _ in range(3): nms.run(b, s, 0.0, 0.1, 0, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): nms.run(b, s, 0.0, 0.1, 0, st)
    e1.record(); torch.cuda.synchronize()
    print("%-8s %9.1f us  kept %d" % (name, 100 * e0.elapsed_time(e1), int(nms.count.item())))
    _lib.profile(True)
    for _ in range(5): nms.run(b, s, 0.0, 0.1, 0, st)
    torch.cuda.synchronize(); _lib.profile(False)
    for k, (c, ms) in sorted(_lib.profile_report().items(), key=lambda kv: -kv[1][1])[:5]:
        print("     %-24s x%.0f %9.1f us" % (k, c / 5, 200 * ms))

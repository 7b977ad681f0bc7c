"""Run one stage of the pipeline a few times (for ncu captures).  usage: run_stage.py {nms|vox|enc|frame} [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_3d_b200 import pipeline, synth, _lib

what = sys.argv[1] if len(sys.argv) > 1 else "frame"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
order = _lib.ORDER_GIVEN if "given" in sys.argv else _lib.ORDER_REFLECTANCE_DESC
g, pfn = synth.G_KITTI, synth.pfn_params(9, 63, seed=5)
if what in ("vox", "enc", "frame"):
    pts = torch.from_numpy(synth.dense_tile()).cuda()
    pipe = pipeline.FramePipeline(g, pfn, pts.shape[0], order=order)
    canvas = pipe.new_canvas()
if what in ("nms", "frame"):
    b, s = synth.nms_boxes(n=20000, seed=4, extent=float(os.environ.get("NMS_EXTENT", "40")))
    b, s = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
    nms = pipeline.NmsStage(20000)
st = torch.cuda.current_stream()
for _ in range(reps):
    if what == "vox":
        pipe.voxelize(pts, st)
    if what in ("enc", "frame"):
        pipe.run(pts, canvas, st)
    if what in ("nms", "frame"):
        nms.run(b, s, 0.0, 0.1, 0, st)
torch.cuda.synchronize()
print("ok", what, reps)

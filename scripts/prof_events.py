"""Per-kernel CUDA-event times (library profiler) for one stage.  usage: prof_events.py {vox|enc|nms|frame} [given] [uniform]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_3d_b200 import pipeline, synth, _lib

what = sys.argv[1] if len(sys.argv) > 1 else "frame"
order = _lib.ORDER_GIVEN if "given" in sys.argv else _lib.ORDER_REFLECTANCE_DESC
g, pfn = synth.G_KITTI, synth.pfn_params(9, 63, seed=5)
reps = 20
pts = torch.from_numpy(synth.uniform_tile() if "uniform" in sys.argv else synth.dense_tile()).cuda()
pipe = pipeline.FramePipeline(g, pfn, pts.shape[0], order=order)
canvas = pipe.new_canvas()
b, s = synth.nms_boxes(n=20000, seed=4, extent=float(os.environ.get("NMS_EXTENT", "40")))
b, s = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
nms = pipeline.NmsStage(20000)
st = torch.cuda.current_stream()

def once():
    if what == "vox":
        pipe.voxelize(pts, st)
    if what in ("enc", "frame"):
        pipe.run(pts, canvas, st, fused=(False if "unfused" in sys.argv else ("features" if "features" in sys.argv else True)))
    if what in ("nms", "frame"):
        nms.run(b, s, 0.0, 0.1, 0, st)

for _ in range(3):
    once()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    once()
e1.record(); torch.cuda.synchronize()
print("%s %s: %.1f us per call (no profiler)" % (what, sys.argv[2:], 1e3 * e0.elapsed_time(e1) / reps))
_lib.profile(True)
for _ in range(reps):
    once()
torch.cuda.synchronize()
_lib.profile(False)
for k, (c, ms) in sorted(_lib.profile_report().items(), key=lambda kv: -kv[1][1]):
    print("  %-28s x%.0f  %8.2f us" % (k, c / reps, 1e3 * ms / reps))
print("  pillars", int(pipe.voxel_num.item()))

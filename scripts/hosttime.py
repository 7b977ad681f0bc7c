"""Host-side enqueue cost of one frame (no synchronisation inside the loop)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_3d_b200 import pipeline, synth
g, pfn = synth.G_KITTI, synth.pfn_params(9, 63, seed=5)
pts = torch.from_numpy(synth.dense_tile(n=1000)).cuda()      # tiny inputs: the GPU keeps up, we time the host
b, s = synth.nms_boxes(n=256, seed=4, extent=40.0)
b, s = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
pipe = pipeline.FramePipeline(g, pfn, 1000); nms = pipeline.NmsStage(20000); canvas = pipe.new_canvas()
st = torch.cuda.current_stream()
for _ in range(20):
    pipe.run(pts, canvas, st); nms.run(b, s, 0.0, 0.1, 0, st)
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 500
for _ in range(K):
    pipe.run(pts, canvas, st); nms.run(b, s, 0.0, 0.1, 0, st)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue per frame: %.1f us ; incl. drain %.1f us" % (1e6 * (t1 - t0) / K, 1e6 * (t2 - t0) / K))

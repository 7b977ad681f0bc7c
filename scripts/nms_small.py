import sys, os
sys.path.insert(0, "/root/repo")
import torch
from objectdetection_3d_b200 import pipeline, synth, _lib
for n in (500, 4000):
    b, s = synth.nms_boxes(n=n, seed=4, extent=12.0)
    b, s = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
    st = torch.cuda.current_stream()
    nms = pipeline.NmsStage(n)
    for _ in range(3): nms.run(b, s, 0.0, 0.1, 0, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): nms.run(b, s, 0.0, 0.1, 0, st)
    e1.record(); torch.cuda.synchronize()
    print("n=%d  %.1f us  kept %d" % (n, 50 * e0.elapsed_time(e1), int(nms.count.item())))

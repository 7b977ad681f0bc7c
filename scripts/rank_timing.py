"""Development: timeline of one frame's kernels (library built with PP_NVCC_EXTRA=-DPP_TIMING): first CTA past its
dependency wait and last CTA done, per kernel, relative to the count kernel's start."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_3d_b200 import pipeline, synth, _lib
g, pfn = synth.G_KITTI, synth.pfn_params(9, 63, seed=5)
pts = torch.from_numpy(synth.uniform_tile() if "uniform" in sys.argv else synth.dense_tile()).cuda()
pipe = pipeline.FramePipeline(g, pfn, pts.shape[0], order=_lib.ORDER_GIVEN if "given" in sys.argv else _lib.ORDER_REFLECTANCE_DESC)
canvas = pipe.new_canvas()
st = torch.cuda.current_stream()
names = {1: "count start", 2: "count end", 4: "cells start", 5: "cells end", 7: "place start", 8: "place end",
         10: "rank start", 11: "rank all CTAs", 12: "rank last CTA end", 13: "bucket start", 14: "bucket end",
         16: "gather start", 17: "gather end"}
starts = {1, 4, 7, 10, 13, 16}
for it in range(4):
    pipe.run(pts, canvas, st, fused=("features" if "features" in sys.argv else True))
    torch.cuda.synchronize()
    ts = pipe.vox_ws[512 + 64: 512 + 64 + 8 * 18].view(torch.int64).cpu().tolist()
    t = {k: ((~ts[k]) & ((1 << 64) - 1) if k in starts else ts[k]) for k in names}
    t0 = t[1]
    print("frame %d: " % it + " | ".join("%s %.1f" % (names[k], (t[k] - t0) / 1e3) for k in sorted(names)))

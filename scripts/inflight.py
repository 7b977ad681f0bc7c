"""Throughput of stage subsets with S frames in flight.  usage: inflight.py [slots]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_3d_b200 import pipeline, synth, _lib
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g, pfn = synth.G_KITTI, synth.pfn_params(9, 63, seed=5)
pts = [torch.from_numpy(synth.dense_tile(seed=3000 + i)).cuda() for i in range(4)]
bs = [synth.nms_boxes(n=20000, seed=4 + i, extent=40.0) for i in range(4)]
bs = [(torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()) for b, s in bs]
slots = [dict(st=torch.cuda.Stream(), pipe=pipeline.FramePipeline(g, pfn, 1_000_000), nms=pipeline.NmsStage(20000)) for _ in range(S)]
for sl in slots: sl["canvas"] = sl["pipe"].new_canvas()
def enq(sl, i, what):
    if "v" in what and "e" in what: sl["pipe"].run(pts[i % 4], sl["canvas"], sl["st"])      # gather + PFN fused
    elif "v" in what: sl["pipe"].voxelize(pts[i % 4], sl["st"])
    elif "e" in what: sl["pipe"].encode_scatter(sl["canvas"], sl["st"])
    if "n" in what: sl["nms"].run(bs[i % 4][0], bs[i % 4][1], 0.0, 0.1, 0, sl["st"])
def run(K, what):
    graphs = []
    for k, sl in enumerate(slots):                # one CUDA graph per slot: host launch cost out of the picture
        with torch.cuda.stream(sl["st"]): enq(sl, k, what)
        sl["st"].synchronize()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_, stream=sl["st"]): enq(sl, k, what)
        graphs.append(g_)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    main = torch.cuda.current_stream()
    for sl in slots: sl["st"].wait_stream(main)
    for i in range(K):
        sl = slots[i % S]
        if i >= S: sl["st"].synchronize()
        with torch.cuda.stream(sl["st"]):
            graphs[i % S].replay()
    for sl in slots: sl["st"].synchronize(); main.wait_stream(sl["st"])
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / K
for what in ("ven", "ve", "v", "e", "n"):
    run(20, what)
    print("slots=%d stages=%-4s %8.1f us/frame" % (S, what, run(400, what)))

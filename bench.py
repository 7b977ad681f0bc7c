#!/usr/bin/env python
"""Benchmark of the PointPillars pre/post-processing hot path (BASELINE.json metric:
voxelize+scatter+NMS frames/s).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One step = one batch of FRAMES_PER_STEP frames per GPU (BASELINE.json configs[2]: 64 tiles over 8 GPUs = 8 per GPU)
of configs[1] (a 1M-point dense tile, G_kitti geometry: 12k pillars x 32 points, 432x496 canvas, reflectance
pre-order as on the model path), each through voxelize -> decorate+PFN -> dense scatter (ONE C-ABI call,
pp_voxelize_scatter) followed by one class of NMS on 20 000 boxes.  For N > 1 every rank runs its own frames
(per-frame data parallelism, no collective on the data path, weak scaling); timing is CUDA events per rank, max over
ranks.  See DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_POINTS = 1_000_000
N_BOXES = 20_000
FRAMES_PER_STEP = 8      # per GPU
NMS_SCORE_THR = 0.0      # every one of the 20k boxes is a candidate (scores are (perm + 0.5) / N > 0)
NMS_IOU_THR = 0.1
NMS_EXTENT = 40.0        # dense case of SURVEY.md 8(d) NMS20k
RING_TILES = 24          # distinct input tiles per GPU  (24 x 16 MB), one per frame slot
RING_CANVAS = 24         # distinct output canvases      (24 x 54.9 MB) -> working set > 126 MB L2
WORKLOAD = "D1M tile (1e6 pts, 0.16 m pillars, 12000x32, 432x496 canvas, reflectance order) + NMS20k dense"
NCU_TABLE = os.path.join(ROOT, "profiles", "r02_ncu_frame_full.json")     # written by scripts/ncu_table.py
CPU_CALIB = os.path.join(ROOT, "profiles", "ref_vs_port_cpu.json")        # written by scripts/calibrate_cpu_baseline.py


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--slots", type=int, default=RING_TILES, help="frames in flight per GPU (<= %d)" % RING_TILES)
    ap.add_argument("--quick", action="store_true", help="headline + e2e only (no workloads / stages / baselines)")
    return ap.parse_args()


def base_config(n_gpus):
    """Identical for both arms (the driver compares the two lines' config)."""
    return {"workload": WORKLOAD, "n_points": N_POINTS, "n_boxes": N_BOXES, "frames_per_step_per_gpu": FRAMES_PER_STEP,
            "global_frames_per_step": FRAMES_PER_STEP * n_gpus,
            "nms": {"score_thr": NMS_SCORE_THR, "iou_thr": NMS_IOU_THR, "extent_m": NMS_EXTENT},
            "parallelism": "frames sharded over %d GPU(s), no collective" % n_gpus,
            "l2": "inputs larger than L2: ring of %d tiles + %d canvases per GPU (%.0f MB)" %
                  (RING_TILES, RING_CANVAS, RING_TILES * 16.0 + RING_CANVAS * 54.85)}


def load_json(path):
    try:
        return json.load(open(path))
    except (OSError, ValueError):
        return None


# ------------------------------------------------------------------------------------------ CPU arm
def make_frame(seed):
    from objectdetection_3d_b200 import synth
    pts = synth.dense_tile(n=N_POINTS, seed=seed)
    boxes, scores = synth.nms_boxes(n=N_BOXES, seed=seed + 7, extent=NMS_EXTENT)
    return pts, boxes, scores


def cpu_frame(O, geom, pfn, pts, boxes, scores):
    """The reference's CPU path for one frame, through the oracle port (oracle/pp_oracle.c)."""
    v, c, n = O.pointpillars_voxelization(pts, geom["voxel_size"], geom["point_cloud_range"],
                                          geom["max_voxel_points"], geom["max_voxels"])
    coors = np.concatenate([np.zeros((len(c), 1), np.int64), c], 1)
    feat = O.pillar_feature_net(v, n, coors, [pfn], geom["voxel_size"], geom["point_cloud_range"])
    canvas = O.scatter_dense(feat, coors.astype(np.int32), 1, 1, 496, 432)
    keep = O.multiclass_nms(boxes, scores, NMS_SCORE_THR, NMS_IOU_THR, 2)[0]
    return canvas, keep


def calibration():
    """port vs the unmodified reference, measured in the build container (scripts/calibrate_cpu_baseline.py)."""
    c = load_json(CPU_CALIB)
    if not c:
        return {"port_speedup_over_reference": None, "calibration": "profiles/ref_vs_port_cpu.json missing"}
    return {"port_speedup_over_reference": c["port_speedup_over_reference"],
            "calibration": "profiles/ref_vs_port_cpu.json: unmodified reference %.2f s/frame vs port %.2f s/frame on the same "
                           "D1M + NMS20k inputs (%d-thread build container; per stage: voxelize %.1fx, PFN+scatter %.1fx, "
                           "NMS %.0fx)" % (c["frame_median_s"]["reference"], c["frame_median_s"]["port"],
                                           c["host"]["cpu_count"], c["per_stage_speedup_median"]["voxelize"],
                                           c["per_stage_speedup_median"]["pfn_scatter"], c["per_stage_speedup_median"]["nms"])}


def cpu_baseline(seconds):
    """Oracle port on the host cores, single thread (the reference's numba kernels are single-threaded,
    ops/ops_numba.py:171,242); a bounded sample of whole frames."""
    from objectdetection_3d_b200 import synth
    from oracle import oracle as O
    O.lib()
    geom, pfn = synth.G_KITTI, synth.pfn_params(9, 63, seed=5)
    pts, boxes, scores = make_frame(3000)
    cpu_frame(O, geom, pfn, pts[:50_000], boxes[:2000], scores[:2000])       # warm-up (page-in)
    t0 = time.perf_counter()
    frames = 0
    while frames < 3 or (time.perf_counter() - t0 < seconds and frames < 64):
        cpu_frame(O, geom, pfn, pts, boxes, scores)
        frames += 1
    dt = time.perf_counter() - t0
    out = {"value": frames / dt, "unit": "frames/s", "cores": 1, "kind": "port",
           "sample": "%d whole frames (same D1M tile + NMS20k) in %.1f s, oracle/pp_oracle.c (gcc -O2) single thread" % (frames, dt)}
    out.update(calibration())
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: the reference is
    pure Python and its checkout is not on the GPU box) on all host threads, one frame per thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from objectdetection_3d_b200 import synth
    from oracle import oracle as O
    O.lib()
    geom, pfn = synth.G_KITTI, synth.pfn_params(9, 63, seed=5)
    threads = max(1, min(os.cpu_count() or 1, 32))
    frames = [make_frame(3000 + i) for i in range(min(threads, 4))]
    pool = ThreadPoolExecutor(threads)

    def step():
        futs = [pool.submit(cpu_frame, O, geom, pfn, *frames[i % len(frames)]) for i in range(threads)]
        for f in futs:
            f.result()

    for _ in range(min(args.warmup, 1)):
        step()
    # bound the run: each step is `threads` whole frames; cap the timed steps so the run ends in minutes
    t_probe = time.perf_counter(); step(); per_step = time.perf_counter() - t_probe
    steps = max(1, min(args.steps, int(120.0 / max(per_step, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = steps * threads / dt
    cb = {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
          "sample": "%d steps x %d whole frames (one per host thread), oracle/pp_oracle.c (gcc -O2)" % (steps, threads)}
    cb.update(calibration())
    line = {"impl": "reference", "metric": "voxelize+scatter+NMS frames/s", "value": value, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(args.gpus), "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled from a thread every ~2 ms (the timed region
    of a 20-step run lasts a few tens of ms; nvidia-smi -lms 100 would return nothing), nvidia-smi as the fallback."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, torch, local):
        self.sm, self.bits, self.h, self.nv, self.smax = [], 0, None, None, None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(local)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            self.nv = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:       # noqa: BLE001 -- any NVML problem -> fallback
            self.err = "%s: %s" % (type(e).__name__, e)
        self.thread = None

    def _loop(self):
        nv, h = self.nv, self.h
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:        # noqa: BLE001
                break
            time.sleep(0.002)

    def start(self):
        if self.h is not None:
            self._stop.clear()
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % getattr(self, "err", "?")],
                    "samples": 0}
        self._stop.set()
        self.thread.join(timeout=2)
        self.thread = None
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.smax,
                "reasons": sorted(n for b, n in self.REASONS.items() if self.bits & b), "samples": len(self.sm),
                "how": "NVML polled every ~2 ms between the first and the last event of the timed region"}


def bind_to_gpu_numa(torch, local):
    """Run this rank on the CPUs of its GPU's NUMA node, so that the pinned host buffers it first-touches (and the
    threads that feed them) sit next to the GPU's PCIe root.  Reports what it found instead of failing silently."""
    info = {"node": None}
    try:
        pr = torch.cuda.get_device_properties(local)
        dev = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % dev).read())
        info["sysfs_numa_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["host_numa_nodes"] = len(nodes)
        if node < 0:
            info["why"] = "the platform reports no NUMA affinity for this GPU (single-node VM)"
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["node"] = node
            info["cpus"] = len(cpus)
    except (OSError, ValueError, AttributeError) as e:
        info["why"] = "%s: %s" % (type(e).__name__, e)
    return info


def naive_gpu_frame(torch, pts_voxels, num, coors, pfn_t, geom, boxes, scores):
    """The "naive GPU" comparison of SURVEY.md 8(d): stages 2-3 as the reference writes them (eager torch ops, one
    small kernel each; model/PointPillars.py:480-526,565-571 and model/utils.py:353-426 restated), on CUDA tensors."""
    vx, vy = geom["voxel_size"][0], geom["voxel_size"][1]
    x_off, y_off = vx / 2 + geom["point_cloud_range"][0], vy / 2 + geom["point_cloud_range"][1]
    P = pts_voxels.shape[1]
    mean = pts_voxels[:, :, :3].sum(dim=1, keepdim=True) / num.view(-1, 1, 1).float()
    f_cluster = pts_voxels[:, :, :3] - mean
    f_center = torch.zeros_like(pts_voxels[:, :, :2])
    f_center[:, :, 0] = pts_voxels[:, :, 0] - (coors[:, 3].float().unsqueeze(1) * vx + x_off)
    f_center[:, :, 1] = pts_voxels[:, :, 1] - (coors[:, 2].float().unsqueeze(1) * vy + y_off)
    feats = torch.cat([pts_voxels, f_cluster, f_center], dim=-1)
    mask = (num.unsqueeze(1) > torch.arange(P, device=num.device).unsqueeze(0)).unsqueeze(-1).float()
    feats = feats * mask
    x = torch.nn.functional.linear(feats, pfn_t["w"])
    x = torch.relu(x * pfn_t["scale"] + pfn_t["shift"])
    x = torch.cat([x.max(dim=1)[0], num.float().unsqueeze(1)], dim=1)
    canvas = torch.zeros((1, x.shape[1], 1, 496, 432), device=x.device)
    canvas[coors[:, 0], :, coors[:, 1], coors[:, 2], coors[:, 3]] = x
    canvas = canvas.view(1, -1, 496, 432)
    # multiclass_nms, nms_dim == 2: the greedy loop over sorted candidates with one device synchronisation per kept box
    from objectdetection_3d_b200 import ops_torch
    rect = ops_torch.bbox2rotated_corners2D(boxes)
    sc = scores[:, 0]
    order = torch.argsort(sc, descending=True, stable=True)
    order = order[sc[order] > NMS_SCORE_THR]
    r = rect[order]
    area = (r[:, 2] - r[:, 0]) * (r[:, 3] - r[:, 1])
    alive = torch.ones(len(r), dtype=torch.bool, device=r.device)
    keep = []
    i = 0
    n = len(r)
    while i < n:
        keep.append(i)
        lt = torch.maximum(r[i, :2], r[:, :2])
        rb = torch.minimum(r[i, 2:], r[:, 2:])
        wh = (rb - lt).clamp(min=0)
        ov = wh[:, 0] * wh[:, 1]
        iou = ov / (area[i] + area - ov).clamp(min=1e-6)
        alive &= ~(iou > NMS_IOU_THR)
        alive[: i + 1] = False
        nxt = torch.nonzero(alive)
        if nxt.numel() == 0:
            break
        i = int(nxt[0])                                   # the per-box synchronisation of the reference's loop
    return canvas, order[torch.tensor(keep, device=r.device)]


def run_ours(args):
    import torch
    import torch.distributed as dist
    from objectdetection_3d_b200 import _lib, model_utils, pipeline, pointpillars, sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(torch, local)
    if world > 1:
        # NCCL prints its version banner on stdout; keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _lib.load()
    K, W = args.steps, max(args.warmup, 3)
    FPS = FRAMES_PER_STEP
    geom, pfn = synth.G_KITTI, synth.pfn_params(9, 63, seed=5)

    # B64-style sharding: frame i of the job -> rank i mod world; every rank holds RING_TILES distinct tiles
    host_pts, host_boxes, host_scores, host_frames = [], [], [], []
    for i in range(RING_TILES):
        p, b, s = make_frame(3000 + sharding.frame_of_rank(i, rank, world))
        host_pts.append(torch.from_numpy(p).pin_memory())
        host_boxes.append(torch.from_numpy(b).pin_memory())
        host_scores.append(torch.from_numpy(s).pin_memory())
        # e2e: a frame's inputs travel as ONE pinned buffer (points | boxes | scores) = one H2D copy per frame
        host_frames.append(torch.cat([host_pts[-1].reshape(-1), host_boxes[-1].reshape(-1),
                                      host_scores[-1].reshape(-1)]).pin_memory())
    d_pts = [t.to(dev) for t in host_pts]
    d_boxes = [t.to(dev) for t in host_boxes]
    d_scores = [t.to(dev) for t in host_scores]
    pipe = pipeline.FramePipeline(geom, pfn, N_POINTS, device=dev)
    nms = pipeline.NmsStage(N_BOXES, device=dev)
    canvases = [pipe.new_canvas() for _ in range(RING_CANVAS)]
    stream = torch.cuda.current_stream()

    def frame(i, p=pipe):
        j = i % RING_TILES
        p.run(d_pts[j], canvases[i % RING_CANVAS], stream)
        nms.run(d_boxes[j], d_scores[j], NMS_SCORE_THR, NMS_IOU_THR, 0, stream)

    def barrier():
        torch.cuda.synchronize()
        sharding.barrier(dist if world > 1 else None)
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        return sharding.max_over_ranks(ms, dist if world > 1 else None, dev)

    for i in range(W):
        frame(i)
    torch.cuda.synchronize()
    m_pillars = int(pipe.voxel_num.item())
    keep_n = int(nms.count.item())

    # ---- frames in flight ----------------------------------------------------------------------
    # Frames are independent, so the host keeps N_SLOTS frames in flight, each on its own stream with its own
    # pipeline buffers: the single-CTA NMS sweep of one frame overlaps the grid-filling kernels of the others,
    # and (e2e) the H2D copy of frame i+1 overlaps the kernels of frame i.  A slot is reused only after its
    # stream has been synchronised, i.e. after that frame's results are complete (e2e: on the host).
    N_SLOTS = max(1, min(args.slots, RING_TILES))
    slots = []
    for k in range(N_SLOTS):
        sl = {"stream": torch.cuda.Stream(device=dev),
              "pipe": pipeline.FramePipeline(geom, pfn, N_POINTS, device=dev),
              "pipe_given": pipeline.FramePipeline(geom, pfn, N_POINTS, order=_lib.ORDER_GIVEN, device=dev),
              "nms": pipeline.NmsStage(N_BOXES, device=dev),
              "frame": torch.empty((host_frames[0].numel(),), dtype=torch.float32, device=dev),
              "canvas": canvases[k % RING_CANVAS],
              "keep": torch.empty((N_BOXES,), dtype=torch.int64).pin_memory(),
              "cnt": torch.empty((2,), dtype=torch.int32).pin_memory(), "graphs": {}}
        slots.append(sl)
    h2d = host_pts[0].numel() * 4 + host_boxes[0].numel() * 4 + host_scores[0].numel() * 4
    d2h = N_BOXES * 8 + 8

    def enqueue(sl, j, mode, st):
        """One frame of slot `sl` on tile j: the C-ABI calls (and, e2e, the host copies) on stream st."""
        p = sl["pipe_given"] if mode == "given" else sl["pipe"]
        if mode in ("e2e", "h2d"):
            sl["frame"].copy_(host_frames[j], non_blocking=True)
            if mode == "h2d":
                return
            np_, nb_ = host_pts[j].numel(), host_boxes[j].numel()
            pts_ = sl["frame"][:np_].view(host_pts[j].shape)
            boxes_ = sl["frame"][np_:np_ + nb_].view(host_boxes[j].shape)
            scores_ = sl["frame"][np_ + nb_:].view(host_scores[j].shape)
        else:
            pts_, boxes_, scores_ = d_pts[j], d_boxes[j], d_scores[j]
        p.run(pts_, sl["canvas"], st)
        sl["nms"].run(boxes_, scores_, NMS_SCORE_THR, NMS_IOU_THR, 0, st)
        if mode == "e2e":
            sl["keep"].copy_(sl["nms"].keep, non_blocking=True)
            sl["cnt"][0:1].copy_(sl["nms"].count, non_blocking=True)
            sl["cnt"][1:2].copy_(p.voxel_num, non_blocking=True)

    def capture(mode):
        """The host cost of ~20 launches per frame would cap the rate: each slot's frame (always on its own tile) is
        captured once into a CUDA graph and replayed with a single launch."""
        for k, sl in enumerate(slots):
            st = sl["stream"]
            with torch.cuda.stream(st):
                enqueue(sl, k % RING_TILES, mode, st)          # warm-up outside capture (lazy attribute setup)
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                enqueue(sl, k % RING_TILES, mode, st)
            sl["graphs"][mode] = g

    def submit(i, mode):
        sl = slots[i % N_SLOTS]
        with torch.cuda.stream(sl["stream"]):
            sl["graphs"][mode].replay()

    def run_in_flight(steps, mode, sampler=None):
        """`steps` batches of FPS frames; returns ms (max over ranks)."""
        if mode not in slots[0]["graphs"]:
            capture(mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if sampler:
            sampler.start()
        e0.record(stream)
        for sl in slots:
            sl["stream"].wait_stream(stream)
        for i in range(steps * FPS):
            if i >= N_SLOTS:
                slots[i % N_SLOTS]["stream"].synchronize()      # frame i - N_SLOTS is complete
            submit(i, mode)
        for sl in slots:
            sl["stream"].synchronize()
            stream.wait_stream(sl["stream"])
        e1.record(stream)
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), clocks

    # ---- headline: device-resident inputs, K steps of FPS frames, N_SLOTS in flight ---------------
    sampler = ClockSampler(torch, local)
    for _ in range(2):
        run_in_flight(max(W, (2 * N_SLOTS + FPS - 1) // FPS), "resident")        # warm-up: every slot's graph has run
    l0 = _lib.launch_count()
    enqueue(slots[0], 0, "resident", stream)                    # count this library's kernels in one frame
    torch.cuda.synchronize()
    launches_per_frame = _lib.launch_count() - l0
    ms_total, clocks = run_in_flight(K, "resident", sampler)
    launches = launches_per_frame * K * FPS                     # replayed from the per-slot CUDA graphs
    value = world * K * FPS / (ms_total * 1e-3)

    # ---- e2e: host buffers in, host results out, through the same C ABI -------------------------
    run_in_flight(max(W, (2 * N_SLOTS + FPS - 1) // FPS), "e2e")
    # the host-buffer path must give the device-resident path's results (checked on tile 0)
    submit(0, "e2e")
    slots[0]["stream"].synchronize()
    frame(0)
    torch.cuda.synchronize()
    assert int(slots[0]["cnt"][0]) == int(nms.count.item()) and int(slots[0]["cnt"][1]) == int(pipe.voxel_num.item())
    assert torch.equal(slots[0]["keep"][:int(slots[0]["cnt"][0])], nms.keep[:int(nms.count.item())].cpu())
    ms_e2e, _ = run_in_flight(K, "e2e")
    e2e_value = world * K * FPS / (ms_e2e * 1e-3)
    # what the host can deliver: the same pinned buffers through the same streams with no kernels behind them
    run_in_flight(W, "h2d")
    ms_h2d, _ = run_in_flight(K, "h2d")
    h2d_ceiling = world * K * FPS / (ms_h2d * 1e-3)

    line = {"metric": "voxelize+scatter+NMS frames/s", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "frames_in_flight": N_SLOTS, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(world),
            "observed": {"pillars": m_pillars, "nms_kept": keep_n, "host_numa": numa,
                         "launches_per_frame": int(launches_per_frame),
                         "execution": "%d independent frames in flight per GPU, one CUDA graph per frame slot" % N_SLOTS},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d * FPS, "d2h_bytes_per_step": d2h * FPS,
                    "ms_per_step": ms_e2e / K,
                    "h2d_only_frames_per_s": h2d_ceiling, "fraction_of_h2d_ceiling": e2e_value / h2d_ceiling,
                    "h2d_GBps_per_gpu": h2d * K * FPS / (ms_h2d * 1e-3) / 1e9,
                    "note": "one pinned host buffer per frame (points | boxes | scores) -> one H2D copy -> same C-ABI calls -> "
                            "D2H keep list + counts; %d frames in flight on %d streams, a frame's results are on the host "
                            "before its slot is reused; the canvas stays on the device for the backbone.  "
                            "h2d_only = the same copies with no kernels (what the host side of PCIe delivers at this "
                            "rank count)" % (N_SLOTS, N_SLOTS)}}

    if not args.quick:
        line.update(detail_legs(args, torch, dist, dev, world, rank, pipe, nms, slots, d_pts, d_boxes, d_scores, host_pts,
                                host_boxes, host_scores, canvases, stream, frame, run_in_flight, barrier, geom, pfn,
                                m_pillars, keep_n, K))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def detail_legs(args, torch, dist, dev, world, rank, pipe, nms, slots, d_pts, d_boxes, d_scores, host_pts, host_boxes,
                host_scores, canvases, stream, frame, run_in_flight, barrier, geom, pfn, m_pillars, keep_n, K):
    """Everything that explains the headline: latency view, stage split, per-kernel event times, roofline, the other
    workloads of SURVEY.md 8(d), the drop-in signature path, the naive-GPU and CPU baselines."""
    from objectdetection_3d_b200 import _lib, model_utils, pipeline, pointpillars, synth
    FPS = FRAMES_PER_STEP
    out = {}

    def per_call_us(fn, reps):
        """us per call on the launching stream: the best of 5 equal chunks of `reps` calls (a chunk hit by a one-off
        stall -- a lazy allocation, a clock ramp -- does not set the figure)."""
        per = max(1, reps // 5)
        best = None
        torch.cuda.synchronize()
        for c in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(per):
                fn(c * per + i)
            e1.record(stream)
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / per
            best = us if best is None or us < best else best
        return best

    reps = 50
    out["single_stream_ms_per_frame"] = per_call_us(frame, reps) * 1e-3
    ms_given, _ = run_in_flight(max(4, K // 4), "given")
    out["given_order"] = {"value": world * max(4, K // 4) * FPS / (ms_given * 1e-3), "unit": "frames/s",
                          "note": "PP_ORDER_GIVEN: points already in processing order (32-bit keys, no reflectance order)"}

    # ---- stage times, single stream (latency view) ----------------------------------------------
    t_ve = per_call_us(lambda i: pipe.run(d_pts[i % RING_TILES], canvases[i % RING_CANVAS], stream), reps)
    t_vf = per_call_us(lambda i: pipe.run(d_pts[i % RING_TILES], canvases[i % RING_CANVAS], stream, fused="features"), reps)
    t_vox = per_call_us(lambda i: pipe.voxelize(d_pts[i % RING_TILES], stream), reps)
    t_enc = per_call_us(lambda i: pipe.encode_scatter(canvases[i % RING_CANVAS], stream), reps)
    t_nms = per_call_us(lambda i: nms.run(d_boxes[i % RING_TILES], d_scores[i % RING_TILES], NMS_SCORE_THR, NMS_IOU_THR, 0, stream), reps)

    def nms_us(mode, boxes, scores, sthr, ithr, r=10):
        st_ = pipeline.NmsStage(boxes.shape[0], device=dev, iou_mode=mode)
        for _ in range(2):
            st_.run(boxes, scores, sthr, ithr, 0, stream)
        t = per_call_us(lambda i: st_.run(boxes, scores, sthr, ithr, 0, stream), r)
        return t, int(st_.count.item())

    t_rot, kept_rot = nms_us(_lib.NMS_ROT_BEV, d_boxes[0], d_scores[0], NMS_SCORE_THR, NMS_IOU_THR, 20)
    t_b3d, kept_b3d = nms_us(_lib.NMS_BOX3D, d_boxes[0], d_scores[0], NMS_SCORE_THR, NMS_IOU_THR, 10)

    # ---- per-kernel durations with CUDA events on the launching stream (library profiler) --------
    _lib.profile(True)
    for i in range(reps):
        frame(i)
    torch.cuda.synchronize()
    _lib.profile(False)
    kern = {k: {"launches_per_frame": c / reps, "us_per_frame": 1e3 * ms / reps, "avg_us": 1e3 * ms / c}
            for k, (c, ms) in _lib.profile_report().items()}
    out["kernels"] = kern

    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json")) or {}
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    vox_b, enc_b = pipe.algorithmic_bytes(N_POINTS, m_pillars)
    canvas_b = (pipe.U + 1) * pipe.D * pipe.H * pipe.W * 4
    vox_out_b = m_pillars * pipe.P * pipe.C * 4 + m_pillars * 16
    # algorithmic bytes per launch of each kernel that moves boundary data (DESIGN.md "Kernels"); scan / rank kernels
    # move only workspace traffic and have 0 algorithmic bytes.  In the one-call frame the canvas zeros are written by
    # the three per-point / per-cell kernels (a third each) and the features by the gather kernel.
    kbytes = {"vox_count_kernel": N_POINTS * pipe.C * 4 + canvas_b // 3,
              "vox_place_kernel": canvas_b // 3, "vox_first_kernel": canvas_b // 3,
              "vox_gather_pfn_kernel": vox_out_b + 2 * m_pillars * (pipe.U + 1) * 4,
              "vox_gather_kernel": vox_out_b,
              "scatter_canvas_kernel": canvas_b + m_pillars * (pipe.U + 1) * 4,
              "pfn_fused_small_kernel": vox_out_b + m_pillars * (pipe.U + 1) * 4}
    ncu = load_json(NCU_TABLE) or {}
    frame_us = max(sum(v["us_per_frame"] for v in kern.values()), 1e-9)

    def roof(name, bound):
        k = kern[name]
        b = kbytes.get(name, 0)
        ach = b / (k["avg_us"] * 1e-6) / 1e9
        # the library's profiler names the PFN-fused gather separately; in the ncu table it is a template instance
        alias = {"vox_gather_pfn_kernel": "vox_gather_kernel<unsigned long long, 1, 1>",
                 "scatter_canvas_kernel": "scatter_canvas_wave_kernel", "sort_pass_kernel": "sort_pass_kernel<2, 0>",
                 "nms_mask_kernel": "nms_mask_kernel<1, 0>", "nms_filter_kernel": "nms_filter_kernel<0>"}
        tab = ncu.get("kernels", {})
        t = tab.get(alias.get(name, name)) or tab.get(name, {})
        return {"kernel": name, "bound": bound, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": t.get("dram_bytes"), "limiter": t.get("limiter"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b, "avg_launch_us": k["avg_us"],
                "share_of_frame": k["us_per_frame"] / frame_us}

    hbm_path = [k for k in kern if k.startswith(("vox_", "pfn_", "scatter_"))]
    if hbm_path:
        dom = max(hbm_path, key=lambda k: kern[k]["us_per_frame"])
        r = roof(dom, "hbm")
        r["how"] = ("CUDA events on the launching stream around every launch (pp_profile_*), separate single-stream pass of "
                    "%d frames after the timed region; traffic / limiter from %s" % (reps, os.path.relpath(NCU_TABLE, ROOT)))
        longest = max(kern, key=lambda k: kern[k]["us_per_frame"])
        r["longest_kernel_of_frame"] = dict(roof(longest, "latency" if longest.startswith(("nms_", "sort_")) else "hbm"),
                                            us_per_frame=kern[longest]["us_per_frame"])
        out["roofline"] = r
    ve_b = vox_b + enc_b
    out["stages"] = {
        "voxelize+scatter": {"algorithmic_MB": ve_b / 1e6, "us": t_ve, "GBps": ve_b / (t_ve * 1e-6) / 1e9,
                             "frac": ve_b / (t_ve * 1e-6) / 1e9 / hbm_peak, "target_frac": 0.6,
                             "sum_of_kernel_us": sum(v["us_per_frame"] for k, v in kern.items() if k.startswith("vox_")),
                             "note": "the product path, ONE call (pp_voxelize_scatter): single stream, back-to-back frames"},
        "voxelize_features+canvas_kernel": {"us": t_vf, "note": "pp_voxelize_features + pp_scatter_mapped"},
        "voxelize": {"algorithmic_MB": vox_b / 1e6, "us": t_vox, "frac": vox_b / (t_vox * 1e-6) / 1e9 / hbm_peak,
                     "note": "stand-alone pp_voxelize (drop-in of points_to_voxel)"},
        "decorate_pfn_scatter": {"algorithmic_MB": enc_b / 1e6, "us": t_enc, "frac": enc_b / (t_enc * 1e-6) / 1e9 / hbm_peak,
                                 "note": "stand-alone pp_pillar_features + pp_scatter_mapped"},
        "nms_20k": {"us": t_nms, "target_us": 1000.0, "round2_review_target_us": 120.0, "kept": keep_n,
                    "pair_test": "xy rectangle of the rotated box (the reference's nms_dim == 2)"},
        "nms_20k_rotated_bev": {"us": t_rot, "target_us": 1000.0, "round2_review_target_us": 400.0, "kept": kept_rot,
                                "pair_test": "rotated BEV footprint IoU (north-star extension)"},
        "nms_20k_box3d": {"us": t_b3d, "round2_review_target_us": 5000.0, "kept": kept_b3d,
                          "pair_test": "oriented 3-D box IoU (the reference's nms_dim == 3, config.yaml:6)"}}

    if world > 1:
        return out

    # ---- the other workloads of SURVEY.md 8(d): device-resident, single stream ------------------
    def vox_workload(name, pts, g, order, with_pfn, r=20):
        t_pts = torch.from_numpy(pts).to(dev)
        ppfn = synth.pfn_params(9, 63 if g is synth.G_KITTI else 19, seed=5)
        wp = pipeline.FramePipeline(g, ppfn, len(pts), order=order, device=dev)
        cv = wp.new_canvas() if with_pfn else None
        fn = (lambda i: wp.run(t_pts, cv, stream)) if with_pfn else (lambda i: wp.voxelize(t_pts, stream))
        for i in range(3):
            fn(i)
        us = per_call_us(fn, r)
        m = int(wp.voxel_num.item())
        vb, eb = wp.algorithmic_bytes(len(pts), m)
        b = vb + (eb if with_pfn else 0)
        return {"us": us, "frames_per_s": 1e6 / us, "pillars": m, "algorithmic_MB": b / 1e6,
                "frac": b / (us * 1e-6) / 1e9 / hbm_peak, "what": "voxelize+PFN+scatter" if with_pfn else "voxelize"}

    wl = {}
    f120 = synth.forest_tile(n=120_000)
    wl["F120k_gref_verbatim"] = vox_workload("F120k", f120, synth.G_REF, _lib.ORDER_REFLECTANCE_DESC, False)
    wl["F120k_pillar250"] = vox_workload("F120k", f120, synth.G_REF_PILLAR, _lib.ORDER_REFLECTANCE_DESC, False)
    wl["D1M_overflow"] = vox_workload("D1M-overflow", synth.uniform_tile(n=N_POINTS), synth.G_KITTI, _lib.ORDER_REFLECTANCE_DESC, True)
    wl["D1M_ties"] = vox_workload("D1M-ties", synth.dense_tile(n=N_POINTS, seed=2025, ties=True), synth.G_KITTI,
                                  _lib.ORDER_REFLECTANCE_DESC, True)
    wl["D1M_given_order"] = vox_workload("D1M", synth.dense_tile(n=N_POINTS), synth.G_KITTI, _lib.ORDER_GIVEN, True)
    grid = {}
    for L in (40.0, 200.0):
        bx, sc = synth.nms_boxes(n=N_BOXES, seed=4, extent=L)
        bx, sc = torch.from_numpy(bx).to(dev), torch.from_numpy(sc).to(dev)
        for mode, mname in ((_lib.NMS_AABB2D, "aabb2d"), (_lib.NMS_ROT_BEV, "rot_bev"), (_lib.NMS_BOX3D, "box3d")):
            for sthr in (0.05, 0.1, 0.3, 0.5, 0.7):
                for ithr in (1e-5, 0.1, 0.5):
                    us, kept = nms_us(mode, bx, sc, sthr, ithr, 5)
                    grid["L%d_%s_s%g_i%g" % (L, mname, sthr, ithr)] = {"us": round(us, 1), "kept": kept}
    wl["NMS20k_grid"] = grid
    out["workloads"] = wl

    # ---- the drop-in signature path: what a user of the reference's API calls --------------------
    vox_mod = pointpillars.PointPillarsVoxelization("cuda", geom["voxel_size"], geom["point_cloud_range"],
                                                    geom["max_voxel_points"], geom["max_voxels"])
    net = pointpillars.PillarFeatureNet(4, [64], geom["voxel_size"], geom["point_cloud_range"]).to(dev).eval()
    with torch.no_grad():
        l = net.pfn_layers[0]
        l.linear.weight.copy_(torch.from_numpy(pfn["weight"])); l.norm.weight.copy_(torch.from_numpy(pfn["gamma"]))
        l.norm.bias.copy_(torch.from_numpy(pfn["beta"])); l.norm.running_mean.copy_(torch.from_numpy(pfn["mean"]))
        l.norm.running_var.copy_(torch.from_numpy(pfn["var"]))
    mid = pointpillars.SparseMiddleExtractor([1, 496, 432])
    np_pts = [t.numpy() for t in host_pts[:4]]
    np_boxes = [t.numpy() for t in host_boxes[:4]]
    np_scores = [t.numpy() for t in host_scores[:4]]

    def dropin_frame(i):
        with torch.no_grad():
            v, c, n = vox_mod(np_pts[i % 4])                                       # numpy in (pageable H2D), device out
            c4 = torch.nn.functional.pad(c, (1, 0), value=0)
            f = net(v, n, c4)
            cv = mid(f, c4, 1)
            keep = model_utils.multiclass_nms(torch.from_numpy(np_boxes[i % 4]).to(dev), torch.from_numpy(np_scores[i % 4]).to(dev),
                                              NMS_SCORE_THR, NMS_IOU_THR, 2)
            return cv, keep[0].cpu()

    for i in range(2):
        dropin_frame(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nd = 12
    for i in range(nd):
        dropin_frame(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["e2e_dropin"] = {"value": nd / dt, "unit": "frames/s", "ms_per_frame": 1e3 * dt / nd,
                         "note": "reference-contract calls, one frame at a time, host wall clock: PointPillarsVoxelization.forward"
                                 "(numpy) -> PillarFeatureNet -> SparseMiddleExtractor -> multiclass_nms(...)[0].cpu(); pageable H2D, "
                                 "per-call output allocation, the .item() sync of the slice bound"}

    # ---- naive GPU: stages 2-3 as eager torch ops on the same device ------------------------------
    pipe.voxelize(d_pts[0], stream)
    torch.cuda.synchronize()
    m = int(pipe.voxel_num.item())
    v_t, n_t = pipe.voxels[:m].clone(), pipe.num[:m].clone().long()
    c_t = torch.nn.functional.pad(pipe.coors[:m][:, [2, 1, 0]].long(), (1, 0), value=0)
    pfn_t = {"w": pipe.w, "scale": pipe.scale, "shift": pipe.shift}
    with torch.no_grad():
        cv_naive, keep_naive = naive_gpu_frame(torch, v_t, n_t, c_t, pfn_t, geom, d_boxes[0], d_scores[0])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            naive_gpu_frame(torch, v_t, n_t, c_t, pfn_t, geom, d_boxes[0], d_scores[0])
        torch.cuda.synchronize()
        t_naive = (time.perf_counter() - t0) / 3
    nms.run(d_boxes[0], d_scores[0], NMS_SCORE_THR, NMS_IOU_THR, 0, stream)
    torch.cuda.synchronize()
    same_keep = bool(torch.equal(keep_naive, nms.keep[:int(nms.count.item())]))
    out["naive_gpu"] = {"ms_per_frame_stages_2_3": 1e3 * t_naive, "ours_us_stages_2_3": t_enc + t_nms,
                        "keep_list_identical": same_keep,
                        "note": "PFN + index_put scatter + greedy NMS loop written as the reference writes them (eager torch ops on "
                                "CUDA tensors, one synchronisation per kept box), same B200"}
    out["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

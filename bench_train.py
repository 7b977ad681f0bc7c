#!/usr/bin/env python
"""Training-step benchmark of BASELINE.json configs[4] ("PointPillars training step with focal/smooth-L1 losses, batch 16
per GPU, NCCL grad allreduce at 2/4/8 B200").

    python bench_train.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        bench_train.py --gpus N --steps 10 --warmup 3

One step = pipeline/pipeline.py:485-499 on a batch of 16 synthetic forest tiles per GPU (120k points each, G_kitti
pillars): voxelize (this library) -> PFN (training mode) -> dense scatter (this library, with backward) -> stock dense
backbone stand-in -> 1x1 conv heads -> assign_bboxes + encode (this library) -> focal / smooth-L1 / cross-entropy
losses -> backward with DistributedDataParallel's NCCL gradient all-reduce -> clip_grad_value_(2) -> AdamW.
Replicas only exchange gradients (weak scaling); timing is CUDA events per rank, max over ranks; rank 0 prints one
JSON line.  This is the secondary benchmark: the driver's headline contract is bench.py.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def make_batch(torch, synth, geom, batch, n_points, seed, dev):
    """F120k-style tiles on G_kitti + 20-60 ground-truth boxes per frame (9 parameters, small rx / ry)."""
    rng = np.random.default_rng(seed)
    rg = geom["point_cloud_range"]
    pts, gts, labels = [], [], []
    sizes = np.asarray(synth.ANCHOR_SIZES, dtype=np.float64)
    for b in range(batch):
        p = synth.forest_tile(n=n_points, seed=seed * 1000 + b, point_cloud_range=rg)
        pts.append(torch.from_numpy(p).to(dev))
        g = int(rng.integers(20, 61))
        box = np.zeros((g, 9))
        box[:, 0] = rng.uniform(rg[0] + 2, rg[3] - 2, g)
        box[:, 1] = rng.uniform(rg[1] + 2, rg[4] - 2, g)
        box[:, 2] = rg[2]
        box[:, 3:6] = sizes[rng.integers(0, len(sizes), g)] * np.exp(rng.normal(0, 0.1, (g, 3)))
        box[:, 6:8] = rng.uniform(-0.1, 0.1, (g, 2))
        box[:, 8] = rng.uniform(0, np.pi, g)
        gts.append(torch.from_numpy(box.astype(np.float32)).to(dev))
        labels.append(torch.zeros((g,), dtype=torch.long, device=dev))
    return pts, gts, labels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16, help="frames per GPU per step")
    ap.add_argument("--points", type=int, default=120_000)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from objectdetection_3d_b200 import _lib, sharding, synth, training

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench_train.py needs a GPU"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)                       # NCCL's banner must not land on stdout
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _lib.load()
    torch.manual_seed(1234)                 # identical initial weights on every rank
    geom = synth.G_KITTI
    net = training.TrainableNet(geom, synth.ANCHOR_SIZES, synth.ANCHOR_ROTATIONS, [[0.08, 0.2]]).to(dev)
    net.train()
    ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else None
    step = training.TrainStep(net, geom, dev, ddp)
    batches = [make_batch(torch, synth, geom, args.batch, args.points, 100 + rank + world * i, dev) for i in range(2)]
    grad_bytes = sum(p.numel() * 4 for p in net.parameters())
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        sharding.barrier(dist if world > 1 else None)
        torch.cuda.synchronize()

    def timed(steps, sync=True):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(steps):
            if sync or ddp is None:
                loss = step(*batches[i % 2])
            else:
                with ddp.no_sync():         # the same step without the gradient all-reduce
                    loss = step(*batches[i % 2])
        e1.record(stream)
        barrier()
        return sharding.max_over_ranks(e0.elapsed_time(e1), dist if world > 1 else None, dev), float(loss)

    W, K = max(args.warmup, 3), args.steps
    timed(W)
    l0 = _lib.launch_count()
    ms, loss = timed(K)
    launches = _lib.launch_count() - l0
    ms_nosync = timed(K, sync=False)[0] if ddp is not None else ms
    # stage split of one step (events between the phases; separate pass)
    marks = []

    class Marker(list):
        def append(self, name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            marks.append((name, ev))

    step(*batches[0], marks=Marker())
    torch.cuda.synchronize()
    stages = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}
    if rank == 0:
        frames = world * args.batch * K
        line = {"metric": "PointPillars training step frames/s", "value": frames / (ms * 1e-3), "unit": "frames/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "T16: batch %d per GPU of %d-point forest tiles on G_kitti, 20-60 boxes per frame" % (args.batch, args.points),
                           "model": "PillarFeatureNet(64) + dense scatter + stock dense backbone stand-in + Anchor3DHead (12 anchors per cell, 248x216 map)",
                           "optimizer": "AdamW lr 1e-4 betas (0.95, 0.99) wd 0.01, clip_grad_value_(2)",
                           "parallelism": "DistributedDataParallel over %d GPU(s): NCCL gradient all-reduce" % world},
                "allreduce": {"grad_bytes": grad_bytes, "ms_per_step_with": ms / K, "ms_per_step_without": ms_nosync / K,
                              "share": max(0.0, 1.0 - ms_nosync / ms) if world > 1 else 0.0,
                              "how": "the same K steps inside DistributedDataParallel.no_sync()"},
                "stage_ms": stages, "loss": loss, "gpu_launches": int(launches)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

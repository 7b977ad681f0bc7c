"""CPU: the training-step harness of configs[4] (objectdetection_3d_b200/training.py, bench_train.py).
(1) the plain-torch loss restatements against the REFERENCE's own loss modules (build container only);
(2) world_size 2 over gloo: DistributedDataParallel around the stand-in backbone -- after one backward every rank holds
    the mean of the per-rank gradients (the one exchange step of the path), and no_sync() skips it."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import ref_shim


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not mounted")
def test_losses_match_reference():
    from objectdetection_3d_b200 import training
    if ref_shim.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    from losses.cross_entropy import CrossEntropyLoss
    from losses.focal_loss import FocalLoss
    from losses.smooth_L1 import SmoothL1Loss
    g = torch.Generator().manual_seed(0)
    pred = torch.randn(4000, 1, generator=g)
    tgt = (torch.rand(4000, generator=g) < 0.1).long().neg() + 1          # 0 = the class, 1 = background (num_classes == 1)
    for af in (37, None):
        assert torch.equal(training.focal_loss(pred, tgt, af), FocalLoss(gamma=2.0, alpha=0.25, loss_weight=1.0)(pred, tgt, avg_factor=af))
    a, b = torch.randn(300, 9, generator=g), torch.randn(300, 9, generator=g) * 0.2
    assert torch.equal(training.smooth_l1_loss(a, b, 300), SmoothL1Loss(beta=0.11, loss_weight=2.0)(a, b, avg_factor=300))
    d, t = torch.randn(300, 2, generator=g), torch.randint(0, 2, (300,), generator=g)
    assert torch.equal(training.cross_entropy_loss(d, t, 300), CrossEntropyLoss(loss_weight=0.2)(d, t, avg_factor=300))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _net():
    from objectdetection_3d_b200 import training
    torch.manual_seed(7)
    return training.DenseBackboneStandIn(in_channels=4, channels=(4, 8), layers=(1, 1), up_channels=4)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = _net()
    ddp = torch.nn.parallel.DistributedDataParallel(net)
    x = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(100 + rank))
    ddp(x).square().mean().backward()
    synced = [p.grad.numpy().copy() for p in net.parameters()]       # (numpy: pickled by value through the queue)
    net.zero_grad()
    with ddp.no_sync():
        ddp(x).square().mean().backward()
    local = [p.grad.numpy().copy() for p in net.parameters()]
    q.put((rank, synced, local))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_gradient_allreduce_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, l0), (_, s1, l1) = out
    for a, b, c, d in zip(s0, s1, l0, l1):
        assert np.array_equal(a, b)                                # every rank holds the same reduced gradient ...
        assert np.allclose(a, 0.5 * (c + d), rtol=1e-5, atol=1e-7)      # ... the mean of the per-rank gradients
    assert any(not np.array_equal(c, d) for c, d in zip(l0, l1))   # (the ranks saw different data)

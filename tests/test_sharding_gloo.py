"""CPU, world_size 2 over gloo: the N>1 path of the benchmark / batch driver -- frames shard by
`i mod world`, every rank runs its own frames (here through the CPU oracle as a stand-in for the GPU path),
results gathered in frame order equal the single-process results; timing is the max over ranks."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _frame_result(i):
    from objectdetection_3d_b200 import synth
    from oracle import oracle as O
    g = synth.G_KITTI
    pts = synth.dense_tile(n=20_000, seed=3000 + i, n_cells=500, n_clusters=20)
    v, c, n = O.pointpillars_voxelization(pts, g["voxel_size"], g["point_cloud_range"], 32, 12000)
    return (int(v.shape[0]), int(n.sum()), float(v.sum()))


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    from objectdetection_3d_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.frames_of_rank(n_frames, rank, world)
    local = [(i, _frame_result(i)) for i in mine]
    allres = sharding.gather_results(local)
    t = sharding.max_over_ranks(1.0 + rank)
    if rank == 0:
        q.put((allres, t, mine))
    dist.barrier()
    dist.destroy_process_group()


def test_frames_shard_without_collectives_on_the_data_path():
    from objectdetection_3d_b200 import sharding
    n_frames, world = 5, 2
    assert sharding.frames_of_rank(8, 1, 4) == [1, 5]
    assert [sharding.frame_of_rank(k, 1, 4) for k in range(2)] == sharding.frames_of_rank(8, 1, 4)
    assert sorted(sum((sharding.frames_of_rank(n_frames, r, world) for r in range(world)), [])) == list(range(n_frames))
    assert sharding.job_throughput([8, 8], [1.0, 2.0]) == 8.0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    allres, t, mine = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert mine == [0, 2, 4] and t == 2.0
    assert [i for i, _ in allres] == list(range(n_frames))
    for i, res in allres:
        assert res == _frame_result(i)


@pytest.mark.reference
def test_dropin_install_swaps_reference_symbols():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference checkout not mounted")
    R = ref_shim.load()
    from objectdetection_3d_b200 import dropin, model_utils, ops_numba, pointpillars
    orig = R.pp.multiclass_nms
    try:
        done = dropin.install()
        assert ("model.PointPillars", "multiclass_nms") in done
        assert R.pp.multiclass_nms is model_utils.multiclass_nms
        assert R.ops_numba.points_to_voxel is ops_numba.points_to_voxel
        assert R.pp.PillarFeatureNet is pointpillars.PillarFeatureNet
        assert R.utils.BBoxCoder is model_utils.BBoxCoder
    finally:
        dropin.uninstall()
    assert R.pp.multiclass_nms is orig

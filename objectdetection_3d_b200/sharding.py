"""Per-frame data parallelism: frames are independent (model/PointPillars.py:114, :1016 loop over frames),
so a batch shards across ranks with no collective on the data path.  bench.py and the gloo test use these helpers;
torch.distributed appears only for the barrier, the max-over-ranks of the timing and the host-side result gather."""


def frames_of_rank(n_frames, rank, world):
    """Frame i of the job runs on rank i mod world."""
    return list(range(rank, n_frames, world))


def frame_of_rank(local_index, rank, world):
    """Job-wide index of the local_index-th frame this rank runs (inverse of frames_of_rank)."""
    return rank + world * local_index


def job_throughput(frames_per_rank, seconds_per_rank):
    """Whole-job frames/s: all frames over the slowest rank's time."""
    return sum(frames_per_rank) / max(seconds_per_rank)


def _active(dist):
    return dist is not None and dist.is_available() and dist.is_initialized()


def barrier(dist=None):
    """Rendezvous of all ranks (no-op for one process)."""
    if dist is None:
        import torch.distributed as dist
    if _active(dist):
        dist.barrier()


def gather_results(local_results, group=None):
    """Collect every rank's per-frame Python results on all ranks, in frame order.
    local_results: list of (frame_index, payload).  Uses all_gather_object (host side, after the path)."""
    import torch.distributed as dist
    if not _active(dist):
        return sorted(local_results)
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local_results, group=group)
    return sorted(x for part in out for x in part)


def max_over_ranks(value, dist=None, device="cpu", group=None):
    """The slowest rank's value (timings are per-rank CUDA events; the job's time is their maximum)."""
    import torch
    if dist is None:
        import torch.distributed as dist
    if not _active(dist):
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())

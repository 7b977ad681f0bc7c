"""Per-frame data parallelism: frames are independent (model/PointPillars.py:114, :1016 loop over frames),
so a batch shards across ranks with no collective on the data path."""


def frames_of_rank(n_frames, rank, world):
    """Frame i of the job runs on rank i mod world."""
    return list(range(rank, n_frames, world))


def job_throughput(frames_per_rank, seconds_per_rank):
    """Whole-job frames/s: all frames over the slowest rank's time."""
    return sum(frames_per_rank) / max(seconds_per_rank)


def gather_results(local_results, group=None):
    """Collect every rank's per-frame Python results on all ranks, in frame order.
    local_results: list of (frame_index, payload).  Uses all_gather_object (host side, after the path)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return sorted(local_results)
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local_results, group=group)
    return sorted(x for part in out for x in part)


def max_over_ranks(value, device="cpu", group=None):
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())

"""Frame sharding for multi-GPU runs: frames are independent (no temporal state in either matcher,
SURVEY.md 8(e)), so a stream is split into contiguous per-rank chunks and no collective touches the
data path.  Only the timing reduction (max over ranks) uses torch.distributed."""
from __future__ import annotations


def shard_range(n_frames: int, rank: int, world: int):
    """Contiguous, balanced split: returns (start, stop) of rank's frames; sizes differ by at most 1."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """All-reduce MAX of a per-rank scalar (device-timed milliseconds)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_throughput(frames_per_rank: int, ms_per_rank: float, units_per_frame: float, dist=None, device=None):
    """value = units all ranks processed / max-over-ranks time.  -> (units per second, ms_max, total frames)"""
    ms_max = max_over_ranks(ms_per_rank, dist, device)
    total = sum_over_ranks(frames_per_rank, dist, device)
    return total * units_per_frame / (ms_max * 1e-3), ms_max, int(round(total))

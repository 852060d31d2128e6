"""Batch-axis sharding of the hot path across GPUs (one process per GPU).

Every op of the backbone is per sample, so rank r owns samples [lo, hi) of the global batch together with
their (h, c) states, which never leave the rank; there is NO data-path collective.  The voxelizer shards
by time window the same way.  The only communication is the benchmark's timing reduction."""
from typing import List, Tuple

import torch
import torch.distributed as dist


def batch_slice(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) share of `global_batch` samples for `rank` (first ranks get the remainder)."""
    assert 0 <= rank < world and global_batch >= 0
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_states(states, lo: int, hi: int):
    """Slice a list of (h, c) states (or None entries) to the local samples."""
    if states is None:
        return None
    return [None if s is None else (s[0][lo:hi], s[1][lo:hi]) for s in states]


def window_shares(n_windows: int, world: int) -> List[Tuple[int, int]]:
    """Per-rank [lo, hi) ranges of event windows for the voxelizer."""
    return [batch_slice(n_windows, r, world) for r in range(world)]


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reported as the max over ranks (bench.py)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

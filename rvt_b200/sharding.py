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


def flat_gradient_view(params) -> torch.Tensor:
    """One flat fp32 tensor aliasing every ``p.grad`` when they already sit back to back in one allocation
    (the layout rvt_b200.train hands to autograd); otherwise None."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or any(g.dtype != torch.float32 or not g.is_contiguous() for g in grads):
        return None
    base = grads[0].untyped_storage().data_ptr()
    off = grads[0].storage_offset()
    for g in grads:
        if g.untyped_storage().data_ptr() != base or g.storage_offset() != off:
            return None
        off += g.numel()
    total = off - grads[0].storage_offset()
    return grads[0].new_empty(0).set_(grads[0].untyped_storage(), grads[0].storage_offset(), (total,))


def allreduce_gradients(module_or_params, group=None, average: bool = True) -> int:
    """Data-parallel gradient reduction of the training step: ONE collective — a NCCL
    all-reduce(SUM) over a flat fp32 gradient buffer — then / world.  In place on ``p.grad``.
    Returns the number of collectives issued (0 when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    params = list(module_or_params.parameters()) if hasattr(module_or_params, 'parameters') else list(module_or_params)
    params = [p for p in params if p.grad is not None]
    if not params:
        return 0
    flat = flat_gradient_view(params)
    copied = flat is None
    if copied:
        flat = torch.cat([p.grad.reshape(-1).float() for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(world)
    if copied:
        off = 0
        for p in params:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
    return 1

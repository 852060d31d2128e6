"""CPU, world_size 2 over gloo: the N>1 host logic of the bench / module — balanced disjoint batch shares,
state slicing, window shares, and the max-over-ranks timing reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rvt_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, global_batch, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = sharding.batch_slice(global_batch, rank, world)
    owned = torch.zeros(global_batch, dtype=torch.int64)
    owned[lo:hi] = 1
    dist.all_reduce(owned)                                   # every sample owned exactly once
    states = [(torch.arange(global_batch).view(-1, 1, 1, 1).float(), torch.zeros(global_batch, 1, 1, 1)), None]
    local = sharding.shard_states(states, lo, hi)
    mx = sharding.max_over_ranks(10.0 + rank)
    q.put((rank, lo, hi, owned.tolist(), local[0][0].flatten().tolist(), local[1], mx))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    world, gb = 2, 17
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, gb, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, owned0, loc0, none0, mx0), (r1, lo1, hi1, owned1, loc1, none1, mx1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 9, 9, 17)
    assert owned0 == owned1 == [1] * gb
    assert loc0 == list(map(float, range(0, 9))) and loc1 == list(map(float, range(9, 17)))
    assert none0 is None and none1 is None
    assert mx0 == mx1 == 11.0


def test_slices_are_balanced_partitions():
    for gb in (0, 1, 8, 24, 63):
        for world in (1, 2, 4, 8):
            parts = [sharding.batch_slice(gb, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == gb
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.window_shares(5, 2) == [(0, 3), (3, 5)]

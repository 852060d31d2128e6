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


def _run_ranks(target, world, *extra, attempts=2):
    """Spawn `world` gloo ranks and collect one queue item per rank; a rendezvous hiccup (the free port picked a moment ago
    got taken, a slow container) gets one retry on a fresh port."""
    import queue as _queue
    last = None
    for _ in range(attempts):
        ctx = mp.get_context('spawn')
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=target, args=(r, world, port, *extra, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            res = [q.get(timeout=120) for _ in range(world)]
            for p in procs:
                p.join(60)
            if all(p.exitcode == 0 for p in procs):
                return res
            last = [p.exitcode for p in procs]
        except _queue.Empty:
            last = 'timeout'
        for p in procs:
            if p.is_alive():
                p.terminate()
    raise AssertionError(f'gloo ranks failed: {last}')


def test_two_rank_sharding_gloo():
    world, gb = 2, 17
    res = sorted(_run_ranks(_worker, world, gb))
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


# ---- training step: single flat all-reduce of the gradients ----
def _grad_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    res = {}
    for layout in ('flat', 'separate'):
        ps = [torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2))]
        if layout == 'flat':
            flat = torch.arange(21, dtype=torch.float32) * (rank + 1)
            off = 0
            for p in ps:
                p.grad = flat[off:off + p.numel()].view(p.shape)
                off += p.numel()
            assert sharding.flat_gradient_view(ps) is not None
        else:
            for i, p in enumerate(ps):
                p.grad = torch.full(p.shape, float((rank + 1) * (i + 1)))
            assert sharding.flat_gradient_view(ps) is None
        n = sharding.allreduce_gradients(ps)
        res[layout] = (n, [p.grad.flatten().tolist() for p in ps])
    q.put((rank, res))
    dist.destroy_process_group()


def test_allreduce_gradients_two_ranks_gloo():
    world = 2
    res = dict(_run_ranks(_grad_worker, world))
    for rank in range(world):
        n, g = res[rank]['flat']
        assert n == 1                                              # ONE collective
        assert sum(g, []) == [i * 1.5 for i in range(21)]          # mean of x1 and x2
        n, g = res[rank]['separate']
        assert n == 1
        assert g == [[1.5] * 12, [3.0] * 5, [4.5] * 4]


def test_grad_sink_hands_autograd_one_flat_buffer():
    """rvt_b200.train._GradSink: accumulators -> parameter gradients as views of one allocation that autograd keeps
    as .grad (so the all-reduce needs no flatten copy); draining zeroes the accumulators; a second backward
    accumulates into the same buffer."""
    import rvt_b200
    from rvt_b200 import train
    from oracle import backbone_oracle as bo
    from tests.golden_configs import BACKBONE_CASES, spec_of
    from tests.test_host_cpu import make_cfg
    spec = spec_of(BACKBONE_CASES['tiny_p6'])
    m = rvt_b200.RNNDetector(make_cfg(spec))
    m.load_state_dict(bo.synth_params(spec, 1), strict=True)
    eng = m._train_engine()
    dev = torch.device('cpu')
    for rep in range(2):
        acc = eng.acc(dev)
        for v in acc.values():
            v.fill_(1.0)
        eng.dirty = True
        tok = train._GradSink.apply(eng, *eng.params)
        tok.sum().backward()
        assert float(eng._acc_flat.abs().sum()) == 0.0 and not eng.dirty
        flat = sharding.flat_gradient_view(list(m.parameters()))
        assert flat is not None and flat.numel() == sum(p.numel() for p in m.parameters())
        w = m.stages[0].lstm.conv1x1.weight
        assert torch.equal(w.grad, torch.full_like(w, float(rep + 1)))     # lstm.G passes through unchanged

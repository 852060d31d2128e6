"""GPU: SURVEY.md 8 f3 (harness glue inside the sequence) and f4 (preprocessing neighbours of the voxelizer) through the
C-ABI, bit-exact against the reference's goldens / the oracle (integer, byte and copy work: the bar is equality)."""
import os

import numpy as np
import pytest
import torch

from oracle import backbone_oracle as bo
from oracle import neighbours_oracle as no
from oracle import voxel_oracle as vo
from tests.golden_configs import BACKBONE_CASES, MIXED_DENSITY_CASES, VOXEL_CASES, make_time_glitched, make_voxel_events
from tests.helpers import GOLD
from tests.test_gpu_backbone import build_module

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _cuda(*arrs):
    return tuple(torch.from_numpy(np.ascontiguousarray(a)).to(DEV) for a in arrs)


# ------------------------------------------------------------------------------------------------- f4
def test_downsample_ev_repr_bit_exact():
    import rvt_b200
    from rvt_b200 import preprocessing as pp
    gold = np.load(os.path.join(GOLD, 'neigh.npz'))
    c = VOXEL_CASES['uniform']
    x, y, p, t = make_voxel_events(c)
    sh = rvt_b200.StackedHistogram(c['bins'], c['height'], c['width'], 10).construct(*_cuda(x, y, p, t))
    got = pp.downsample_ev_repr(sh.unsqueeze(0), 0.5)
    assert got.shape == (1, 2 * c['bins'], c['height'] // 2, c['width'] // 2) and got.dtype == torch.uint8
    assert np.array_equal(got[0].cpu().numpy(), gold['ds_u8'])
    odd = torch.from_numpy(gold['ds_odd_in']).to(DEV)
    assert np.array_equal(pp.downsample_ev_repr(odd).cpu().numpy(), gold['ds_odd'])
    # int8 representations go through the same byte gather (the reference shifts by 128 and back, :470-476)
    i8 = torch.from_numpy(np.random.RandomState(1).randint(-128, 128, (10, 90, 160)).astype(np.int8)).to(DEV)
    assert np.array_equal(pp.downsample_ev_repr(i8).cpu().numpy(), no.downsample_ev_repr(i8.cpu().numpy()))
    # BASELINE size: 20 x 720 x 1280
    big = torch.randint(0, 11, (20, 720, 1280), dtype=torch.uint8, device=DEV)
    assert torch.equal(pp.downsample_ev_repr(big), big[:, 1::2, 1::2])


@pytest.mark.parametrize('n', [1, 17, 4096, 4097, 20000, 3_000_001])
def test_correct_time_bit_exact(n):
    from rvt_b200 import preprocessing as pp
    tg = make_time_glitched(31, n)
    got = pp.correct_time_(torch.from_numpy(tg.copy()).to(DEV)).cpu().numpy()
    assert np.array_equal(got, no.correct_time(tg))
    if n == 20000:
        assert np.array_equal(got, np.load(os.path.join(GOLD, 'neigh.npz'))['ct'])
    assert np.all(got[1:] >= got[:-1])                       # size-independent property: monotone, idempotent
    assert np.array_equal(pp.correct_time_(torch.from_numpy(got.copy()).to(DEV)).cpu().numpy(), got)


def test_event_window_indices_bit_exact():
    from rvt_b200 import preprocessing as pp
    gold = np.load(os.path.join(GOLD, 'neigh.npz'))
    ts = np.sort(np.random.RandomState(32).randint(0, 2_000_000, 50000).astype(np.int64))
    q = np.arange(50_000, 2_000_000, 50_000, dtype=np.int64)
    (tsd, qd) = _cuda(ts, q)
    s_d, e_d = pp.event_window_indices(tsd, qd, None, 50)
    s_n, e_n = pp.event_window_indices(tsd, qd, 3000, None)
    assert np.array_equal(s_d.cpu().numpy(), gold['win_start_dt']) and np.array_equal(e_d.cpu().numpy(), gold['win_end'])
    assert np.array_equal(s_n.cpu().numpy(), gold['win_start_n']) and torch.equal(e_n, e_d)
    # duplicates, queries below / above the range, empty array
    ts2 = np.array([5, 5, 5, 7, 7, 9], np.int64)
    q2 = np.array([0, 5, 6, 7, 9, 10], np.int64)
    for side in ('left', 'right'):
        got = pp.searchsorted(*_cuda(ts2, q2), side=side).cpu().numpy()
        assert np.array_equal(got, np.searchsorted(ts2, q2, side=side))
    assert np.array_equal(pp.searchsorted(torch.empty(0, dtype=torch.int64, device=DEV), _cuda(q2)[0]).cpu().numpy(), np.zeros(6, np.int64))


@pytest.mark.parametrize('name', list(MIXED_DENSITY_CASES))
def test_mixed_density_matches_reference_golden(name):
    import rvt_b200
    c = MIXED_DENSITY_CASES[name]
    x, y, p, t = make_voxel_events(c)
    md = rvt_b200.MixedDensityEventStack(c['bins'], c['height'], c['width'], c['cutoff'])
    got = md.construct(*_cuda(x, y, p, t))
    assert got.dtype == torch.int8 and tuple(got.shape) == md.get_shape()
    assert np.array_equal(got.cpu().numpy(), np.load(os.path.join(GOLD, 'neigh.npz'))[name])
    got2 = md.construct(*_cuda(x, y, p, t))                  # scratch left clean
    assert torch.equal(got, got2)


def test_mixed_density_large_vs_oracle_and_properties():
    import rvt_b200
    n, H, W, bins = 3_000_000, 360, 640, 10
    x, y, p, t = vo.synth_events(41, n, H, W, hot_fraction=0.02, hot_pixels=8)
    md = rvt_b200.MixedDensityEventStack(bins, H, W, None)
    got = md.construct(*_cuda(x, y, p, t)).cpu().numpy()
    assert np.array_equal(got, no.mixed_density_stack(x, y, p, t, bins, H, W, None))
    # last channel = wrapped total signed count per pixel
    tot = np.zeros(H * W, np.int64)
    np.add.at(tot, x + W * y, p * 2 - 1)
    assert np.array_equal(got[-1].reshape(-1), (((tot + 128) & 0xFF) - 128).astype(np.int8))
    with pytest.raises(AssertionError):
        md.construct(*_cuda(x, y, p * 3, t))


# ------------------------------------------------------------------------------------------------- f3
def _harness_reference(m, xs, states, is_first, selected):
    """the reference's time loop (modules/detection.py:117-159) with its own glue semantics, on chained forward()"""
    if states is not None and is_first is not None:
        for (h, c) in states:                               # RNNStates.reset -> recursive_reset (in place)
            assert h.requires_grad is False
            h[is_first] = 0
            c[is_first] = 0
    feats = {}
    with torch.no_grad():
        for t, x in enumerate(xs):
            out, states = m(x, states)
            if selected[t]:
                for k, v in out.items():                    # BackboneFeatureSelector.add_backbone_features
                    feats.setdefault(k, []).append(v[selected[t]])
    return {k: torch.cat(v, dim=0) for k, v in feats.items()}, states


@pytest.mark.parametrize('graph', [False, True])
def test_sequence_with_reset_mask_and_feature_selection(graph):
    import rvt_b200
    case = dict(BACKBONE_CASES['tiny_p6'])
    m, _, _ = build_module(case)
    L, B = 5, 4
    xs = torch.stack([bo.synth_events_tensor(300 + t, B, 20, 64, 96) for t in range(L)]).to(DEV)
    with torch.no_grad():
        _, st0 = m(bo.synth_events_tensor(299, B, 20, 64, 96).to(DEV).float())
    clone = lambda st: [(h.clone(), c.clone()) for h, c in st]
    is_first = torch.tensor([False, True, False, True], device=DEV)
    selected = [[], [0, 2], [], [1], [0, 1, 2, 3]]
    flat = [t * B + b for t, sel in enumerate(selected) for b in sel]
    S = 12                                                   # static capacity, 7 slots used
    sel_idx = torch.full((S,), -1, dtype=torch.int32, device=DEV)
    sel_idx[:len(flat)] = torch.tensor(flat, dtype=torch.int32)
    ref_feats, ref_states = _harness_reference(m, list(xs), clone(st0), is_first, selected)

    st_in = clone(st0)
    if graph:
        g = rvt_b200.GraphedCallable(lambda: m.forward_sequence(xs, st_in, reset_mask=is_first, select=sel_idx), warmup=1)
        for (h, c), (h0, c0) in zip(st_in, st0):            # the warm-up / capture runs reset the static states in place: refill
            h.copy_(h0)
            c.copy_(c0)
        outs, states, feats = g()
    else:
        outs, states, feats = m.forward_sequence(xs, st_in, reset_mask=is_first, select=sel_idx)
    torch.cuda.synchronize()
    for (h, c) in st_in:                                     # in-place reset of the caller's tensors, like the reference
        assert float(h[is_first].abs().max()) == 0 and float(c[is_first].abs().max()) == 0
    for (h, c), (h2, c2) in zip(states, ref_states):
        assert torch.equal(h, h2) and torch.equal(c, c2)
    for k in (1, 2, 3, 4):
        assert torch.equal(feats[k][:len(flat)], ref_feats[k])
        assert float(feats[k][len(flat):].abs().max()) == 0  # empty slots are zero

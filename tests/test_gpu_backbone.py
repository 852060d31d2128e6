"""GPU: the CUDA backbone (rvt_b200.RNNDetector, all arithmetic through the C-ABI) against
  (1) the REFERENCE's committed golden outputs (tests/golden, minted from /root/reference),
  (2) the fp32 CPU oracle on the same seeded inputs, operator by operator (taps),
over multi-step sequences with state carry and harness-style in-place resets.

Tolerances (relative = max|a-b| / max|b|):
  per operator, identical inputs: 1e-3 — tests/test_gpu_ops.py (the north-star parity bar).
  TOL_FP32 = 2e-2   END-TO-END envelope of this file, against the PURE-fp32 reference over whole
                    multi-step sequences: fp16 operand rounding (2^-11 per operand; the reference's
                    own `precision: 16`, config/general.yaml:6) compounds through ~50 chained GEMMs
                    and is rescaled by the LSTM gates (|h| < 1).  Measured: 2.5e-4 after the stem,
                    <= 1.1e-2 at stage 4 (profiles/parity_r01.md).  The reference under AMP deviates
                    from its own fp32 run by the same mechanism; an oracle emulating fp16 operand
                    rounding decorrelates after the first rounding, so it is no tighter.
  Long sequences: the fp16-operand deviation of the recurrent states keeps growing with the step count for the REFERENCE's
  own AMP run too (measured on the B200, profiles/amp_envelope_r02.json: AMP-reference vs fp32-reference 2.4e-2 at step 20 of
  an RVT-B 1Mpx sequence, rvt_b200 vs fp32 2.9e-2).  A fixed bar therefore only makes sense for short sequences: steps < 5 are
  held to TOL_FP32, later steps to tol_at(step) = TOL_FP32 * (1 + (step - 4) / 8)  (6e-2 at step 20), and the drift itself is
  pinned against the AMP reference in tests/test_gpu_parity_envelope.py (ours <= 1.5 x AMP-reference at every reported step)."""
import json
import os

import pytest
import torch

from oracle import backbone_oracle as bo
from tests.golden_configs import BACKBONE_CASES, spec_of
from tests.helpers import check_against_golden, rel_err, replay_case
from tests.test_host_cpu import make_cfg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_FP32 = 2e-2


def tol_at(step: int) -> float:
    return TOL_FP32 if step < 5 else TOL_FP32 * (1 + (step - 4) / 8)


def build_module(case):
    import rvt_b200
    spec = spec_of(case)
    m = rvt_b200.build_recurrent_backbone(make_cfg(spec))
    params = bo.synth_params(spec, case['seed'], case.get('gamma_mode', 'uniform'))
    m.load_state_dict(params, strict=True)
    return m.cuda().eval(), params, spec


def _report(name, rows):
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', f'parity_{name}.json'), 'w') as f:
        json.dump(rows, f, indent=1)


@pytest.mark.parametrize('name', list(BACKBONE_CASES))
def test_operator_taps_match_oracle(name):
    """Every operator's output (residual stream after conv+LN / attention / MLP, then h, c) vs the
    pure-fp32 oracle end to end (TOL_FP32), step by step; the report pinpoints the first
    diverging operator."""
    case = BACKBONE_CASES[name]
    m, params, spec = build_module(case)
    o_states = None
    rows, worst = [], 0.0
    taps_gpu = {}
    m.debug_taps = taps_gpu

    def step(x, states, mask):
        with torch.no_grad():
            return m(x, states, mask)

    for stepi, feats, states in replay_case(case, step, device='cuda'):
        x = bo.synth_events_tensor(case['seed'] * 1000 + stepi, case['batch'], 20, case['height'], case['width']).float()
        mask = None
        if spec.enable_masking:
            import numpy as np
            rs = np.random.RandomState(case['seed'] + 77 + stepi)
            mask = torch.from_numpy(rs.uniform(size=(case['batch'], case['height'] // 4, case['width'] // 4)) < 0.2)
        taps_cpu = {}
        with torch.no_grad():
            o_out, o_states = bo.backbone_forward(x, o_states, params, spec, mask, taps=taps_cpu)
        if case.get('reset_at') == stepi:
            for (hh, cc) in o_states:
                hh[0] = 0
                cc[0] = 0
        for k, v in taps_gpu.items():
            e = rel_err(v, taps_cpu[k])
            rows.append((stepi, k, e))
            worst = max(worst, e)
        for s in range(4):
            for tag, i in (('h', 0), ('c', 1)):
                e = rel_err(states[s][i], o_states[s][i])
                rows.append((stepi, f'stage{s}.{tag}', e))
                worst = max(worst, e)
            assert feats[s + 1].shape == o_out[s + 1].shape
            assert feats[s + 1].dtype == torch.float32
    _report(name, rows)
    bad = [r for r in rows if not r[2] <= tol_at(r[0])]
    assert not bad, f'first diverging operator: {bad[0]} (worst {worst:.3e})'


@pytest.mark.parametrize('name', list(BACKBONE_CASES))
def test_backbone_matches_reference_golden(name):
    case = BACKBONE_CASES[name]
    m, _, _ = build_module(case)

    def step(x, states, mask):
        with torch.no_grad():
            return m(x, states, mask)

    worst = check_against_golden(name, case, step, tol=tol_at, device='cuda')
    print(f'{name}: worst rel err vs reference golden {worst:.3e}')


def test_uint8_input_and_folded_padding_equal_float_padded():
    """uint8 event tensors and the folded zero-padding (pad_to_hw) give the same result as the
    harness path (float32 + explicit pad, modules/detection.py:133-134, utils/padding.py:29-44)."""
    case = dict(BACKBONE_CASES['tiny_p6'])
    m, _, spec = build_module(case)
    x8 = bo.synth_events_tensor(5, 2, 20, 60, 90).cuda()            # un-padded 60x90 -> model 64x96
    xf = torch.nn.functional.pad(x8.float(), (0, 6, 0, 4))
    with torch.no_grad():
        ref, _ = m(xf)
        m.pad_to_hw = (64, 96)
        got, _ = m(x8)
    for s in range(1, 5):
        assert torch.equal(ref[s], got[s])


def test_states_are_harness_compatible():
    """Returned states: fp32, no grad, channels-last strides like the reference's; in-place index
    reset works and non-contiguous / None states are accepted back."""
    case = BACKBONE_CASES['tiny_p6']
    m, _, _ = build_module(case)
    x = bo.synth_events_tensor(1, 2, 20, 64, 96).float().cuda()
    with torch.no_grad():
        out, st = m(x)
        for (h, c) in st:
            assert h.dtype == c.dtype == torch.float32 and not h.requires_grad
            assert h.permute(0, 2, 3, 1).is_contiguous()
            h[torch.tensor([True, False], device='cuda')] = 0
        st2 = [(h.contiguous(), c.contiguous()) for h, c in st]     # NCHW-contiguous copies
        st2[1] = None
        out2, _ = m(x, st2)
    assert set(out2) == {1, 2, 3, 4}
    out3, st3 = m(x)   # grad mode: the autograd-visible training path (rvt_b200/train.py), states attached to the graph
    assert all(h.requires_grad and c.requires_grad for h, c in st3) and out3[4].requires_grad


@pytest.mark.parametrize('name', ['tiny_p6', 'dws_hidden', 'rvt_t_gen1'])
def test_forward_sequence_equals_chained_forward(name):
    """forward_sequence (sequential and 4-stream wavefront schedule) is bit-identical to chaining
    the reference-API forward() over the timesteps, including a non-None initial state."""
    case = BACKBONE_CASES[name]
    m, _, spec = build_module(case)
    L = 4
    xs = [bo.synth_events_tensor(900 + t, case['batch'], 20, case['height'], case['width']).cuda() for t in range(L)]
    with torch.no_grad():
        _, st0 = m(xs[0].float())
        ref_out, st = [], st0
        for t in range(L):
            o, st = m(xs[t], st)
            ref_out.append(o)
        for wf in (False, True):
            outs, st2 = m.forward_sequence(xs, st0, wavefront=wf)
            torch.cuda.synchronize()
            for t in range(L):
                for k in (1, 2, 3, 4):
                    assert torch.equal(outs[t][k], ref_out[t][k]), (wf, t, k)
            for (h, c), (h2, c2) in zip(st, st2):
                assert torch.equal(h, h2) and torch.equal(c, c2)


def test_graphed_sequence_matches_eager():
    """rvt_b200.capture_sequence: replaying the captured multi-stream graph (twice, with refilled static
    inputs) reproduces the eager forward_sequence bit for bit."""
    import rvt_b200
    case = BACKBONE_CASES['tiny_p6']
    m, _, _ = build_module(case)
    L = 3
    xs = torch.stack([bo.synth_events_tensor(700 + t, 2, 20, 64, 96) for t in range(L)]).cuda()
    g = rvt_b200.capture_sequence(m, xs)
    for rep in range(2):
        fresh = torch.stack([bo.synth_events_tensor(800 + 10 * rep + t, 2, 20, 64, 96) for t in range(L)]).cuda()
        xs.copy_(fresh)
        outs, st = g()
        torch.cuda.synchronize()
        with torch.no_grad():
            ref_outs, ref_st = m.forward_sequence(fresh, None, wavefront=False)
        for t in range(L):
            for k in (1, 2, 3, 4):
                assert torch.equal(outs[t][k], ref_outs[t][k]), (rep, t, k)
        for (h, c), (h2, c2) in zip(st, ref_st):
            assert torch.equal(h, h2) and torch.equal(c, c2)


def test_graphed_wavefront_sequence_at_bench_shape_matches_eager():
    """BASELINE configs[1] shape (RVT-B 1Mpx, bs 8, 21 timesteps): the captured 4-stream wavefront graph, replayed twice
    with refilled inputs, is bit-identical to the strictly sequential eager schedule.  Guards the cross-stream lifetime of
    the per-step fp16 feature copies under capture (a recycled block would be a silent race in the replayed graph)."""
    import rvt_b200
    case = dict(BACKBONE_CASES['rvt_b_1mpx'])
    m, _, _ = build_module(case)
    m.pad_to_hw = (384, 640)
    L, B = 21, 8
    g = torch.Generator(device='cuda').manual_seed(5)

    def fresh():
        v = torch.randint(1, 11, (L, B, 20, 360, 640), generator=g, device='cuda', dtype=torch.uint8)
        return v * (torch.randint(0, 10, v.shape, generator=g, device='cuda', dtype=torch.uint8) == 0)

    xs = fresh()
    graphed = rvt_b200.capture_sequence(m, xs)
    for rep in range(2):
        new = fresh()
        xs.copy_(new)
        outs, st = graphed()
        torch.cuda.synchronize()
        with torch.no_grad():
            ref_outs, ref_st = m.forward_sequence(new, None, wavefront=False)
        torch.cuda.synchronize()
        for t in (0, 1, 7, 19, 20):
            for k in (1, 2, 3, 4):
                assert torch.equal(outs[t][k], ref_outs[t][k]), (rep, t, k)
        for (h, c), (h2, c2) in zip(st, ref_st):
            assert torch.equal(h, h2) and torch.equal(c, c2)

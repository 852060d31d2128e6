"""GPU: the CUDA backbone (rvt_b200.RNNDetector, all arithmetic through the C-ABI) against
  (1) the REFERENCE's committed golden outputs (tests/golden, minted from /root/reference),
  (2) the fp32 CPU oracle on the same seeded inputs, operator by operator (taps),
over multi-step sequences with state carry and harness-style in-place resets.

Tolerance: 1e-3 relative (max|a-b| / max|b|) against the pure-fp32 reference — the north-star
'1e-3 relative fp16/bf16' bar; the CUDA path uses fp16 operands with fp32 accumulation."""
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
TOL = 1e-3


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


@pytest.mark.parametrize('name', ['tiny_p6', 'small_dh24', 'dws_hidden', 'dws_xh', 'ls_init_mask'])
def test_operator_taps_match_oracle(name):
    """Every operator's output (residual stream after conv+LN / attention / MLP, then h, c) vs the
    fp32 oracle, step by step; the report pinpoints the first diverging operator."""
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
            for tag, a, b in (('h', states[s][0], o_states[s][0]), ('c', states[s][1], o_states[s][1])):
                e = rel_err(a, b)
                rows.append((stepi, f'stage{s}.{tag}', e))
                worst = max(worst, e)
            assert feats[s + 1].shape == o_out[s + 1].shape
            assert feats[s + 1].dtype == torch.float32
    _report(name, rows)
    bad = [r for r in rows if not r[2] <= TOL]
    assert not bad, f'first diverging operator: {bad[0]} (worst {worst:.3e})'


@pytest.mark.parametrize('name', list(BACKBONE_CASES))
def test_backbone_matches_reference_golden(name):
    case = BACKBONE_CASES[name]
    m, _, _ = build_module(case)

    def step(x, states, mask):
        with torch.no_grad():
            return m(x, states, mask)

    worst = check_against_golden(name, case, step, tol=TOL, device='cuda')
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
    with pytest.raises(NotImplementedError):
        m(x)        # grad mode: backward kernels not built -> loud failure, not a silent fallback

"""GPU: the training step end to end — rvt_b200.RNNDetector under autograd over an unrolled sequence
(modules/detection.py:150-199) against (a) the REFERENCE's gradients committed in tests/golden/backbone_grads_*.npz
and (b) autograd through the fp32 oracle on the same inputs.

Tolerance: fp16 tensor-core operands and fp16 branch-internal gradient signals (as the reference under
precision-16 AMP) against a pure-fp32 comparator: per-parameter rel-L2 <= 3e-2, loss <= 1e-3 relative."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import backbone_oracle as bo  # noqa: E402
from tests.golden_configs import BACKBONE_CASES, spec_of  # noqa: E402
from tests.helpers import GOLD, GRAD_CASES, grad_sub, case_inputs, train_loss  # noqa: E402
from tests.test_host_cpu import make_cfg  # noqa: E402

GRAD_TOL = 3e-2


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    return torch.device('cuda:0')


def build(case, dev):
    import rvt_b200
    spec = spec_of(case)
    params = bo.synth_params(spec, case['seed'], case.get('gamma_mode', 'uniform'))
    m = rvt_b200.RNNDetector(make_cfg(spec))
    m.load_state_dict(params, strict=True)
    return m.to(dev).train(), params, spec


def run_ours(m, xs, dev):
    outs, st = [], None
    for x in xs:
        o, st = m(x.to(dev), st)
        outs.append(o)
    return outs, st


@pytest.mark.parametrize('wavefront', [False, True])
@pytest.mark.parametrize('name', list(GRAD_CASES))
def test_gradients_match_reference_golden(name, wavefront, dev):
    case = BACKBONE_CASES[name]
    m, params, spec = build(case, dev)
    m.train_wavefront = wavefront        # stage-per-stream schedule (forward and, through autograd, backward)
    xs = case_inputs(case, GRAD_CASES[name])
    outs, st = run_ours(m, xs, dev)
    for o in outs:
        for s in range(1, 5):
            assert o[s].requires_grad and o[s].dtype == torch.float32
    loss = train_loss(outs, st)
    loss.backward()
    gold = np.load(os.path.join(GOLD, f'backbone_grads_{name}.npz'))
    assert abs(float(loss.detach()) - float(gold['loss'])) <= 1e-3 * abs(float(gold['loss']))
    worst = {}
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        assert torch.isfinite(p.grad).all(), k
        ref = torch.from_numpy(gold['g.' + k]).double()
        got = p.grad.detach().cpu().contiguous().reshape(-1)[::grad_sub(name)].double()
        # subsample error measured against the subsample's own norm (floor: 1/sqrt(stride) of the full norm)
        den = max(float(ref.norm()), float(gold['n.' + k]) / np.sqrt(grad_sub(name)) * 0.1)
        worst[k] = float((got - ref).norm()) / max(den, 1e-20)
    bad = {k: v for k, v in worst.items() if v > GRAD_TOL}
    print(f'{name}: worst grad rel-L2 vs reference golden: {max(worst.values()):.3e} ({max(worst, key=worst.get)})')
    assert not bad, bad


@pytest.mark.parametrize('name,steps', [('tiny_p6', 3), ('rvt_t_gen1', 2)])
def test_gradients_match_oracle_autograd(name, steps, dev):
    case = BACKBONE_CASES[name]
    m, params, spec = build(case, dev)
    xs = case_inputs(case, steps)
    outs, st = run_ours(m, xs, dev)
    loss = train_loss(outs, st)
    loss.backward()
    po = {k: v.to(dev).requires_grad_(True) for k, v in params.items()}
    o_outs, o_st = [], None
    for x in xs:
        o, o_st = bo.backbone_forward(x.to(dev).float(), o_st, po, spec)
        o_outs.append(o)
    loss_o = train_loss(o_outs, o_st)
    loss_o.backward()
    assert abs(float(loss.detach()) - float(loss_o.detach())) <= 1e-3 * abs(float(loss_o.detach()))
    worst = {}
    for k, p in m.named_parameters():
        worst[k] = float((p.grad - po[k].grad).double().norm() / po[k].grad.double().norm().clamp_min(1e-20))
    print(f'{name}: worst grad rel-L2 vs oracle autograd: {max(worst.values()):.3e} ({max(worst, key=worst.get)})')
    bad = {k: v for k, v in worst.items() if v > GRAD_TOL}
    assert not bad, bad


def test_training_forward_equals_inference_forward(dev):
    """The autograd-visible forward (three-kernel operators) and the fused inference forward agree to fp16 rounding."""
    case = BACKBONE_CASES['tiny_p6']
    m, _, _ = build(case, dev)
    xs = case_inputs(case, 2)
    outs_t, st_t = run_ours(m, xs, dev)
    with torch.no_grad():
        outs_i, st_i = run_ours(m, xs, dev)
    for s in range(4):
        for a, b in zip(st_t[s], st_i[s]):
            assert float((a.detach() - b).abs().max() / b.abs().max()) < 1e-2


def test_input_state_gradients_and_detached_states(dev):
    """dL/d(h_prev, c_prev) flow to caller-provided states; detached states (TBPTT boundary,
    modules/utils/detection.py:84-93) stop the graph; a second backward pass starts from clean accumulators."""
    case = BACKBONE_CASES['tiny_p6']
    m, params, spec = build(case, dev)
    xs = case_inputs(case, 2)
    with torch.no_grad():
        _, st0 = m(xs[0].to(dev), None)
    st_in = [(h.clone().requires_grad_(True), c.clone().requires_grad_(True)) for h, c in st0]
    o, st = m(xs[1].to(dev), st_in)
    loss = train_loss([o], st)
    loss.backward()
    po = {k: v.to(dev).requires_grad_(True) for k, v in params.items()}
    st_o = [(h.detach().clone().requires_grad_(True), c.detach().clone().requires_grad_(True)) for h, c in st0]
    oo, sto = bo.backbone_forward(xs[1].to(dev).float(), st_o, po, spec)
    train_loss([oo], sto).backward()
    for s in range(4):
        for a, b in zip(st_in[s], st_o[s]):
            e = float((a.grad - b.grad).double().norm() / b.grad.double().norm().clamp_min(1e-20))
            assert e < GRAD_TOL, (s, e)
    g1 = {k: p.grad.clone() for k, p in m.named_parameters()}
    # second, independent pass with detached states: gradients must not contain leftovers of the first
    m.zero_grad(set_to_none=True)
    o, st = m(xs[1].to(dev), [(h.detach(), c.detach()) for h, c in st_in])
    train_loss([o], st).backward()
    for k, p in m.named_parameters():
        e = float((p.grad - g1[k]).double().norm() / g1[k].double().norm().clamp_min(1e-20))
        assert e < 1e-3, (k, e)


def test_optimizer_step_repacks_weights(dev):
    """After an in-place parameter update (optimizer.step) the next forward uses the new weights: the packed
    copies refresh, and the output tracks the oracle evaluated at the updated parameters."""
    case = BACKBONE_CASES['tiny_p6']
    m, _, spec = build(case, dev)
    xs = case_inputs(case, 1)
    opt = torch.optim.SGD(m.parameters(), lr=1e-4)
    o, st = m(xs[0].to(dev), None)
    before = [h.detach().clone() for h, _ in st]
    train_loss([o], st).backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    o, st = m(xs[0].to(dev), None)
    po = {k: v.detach().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        _, sto = bo.backbone_forward(xs[0].to(dev).float(), None, po, spec)
    moved = 0.0
    for s in range(4):
        e = float((st[s][0].detach() - sto[s][0]).abs().max() / sto[s][0].abs().max())
        assert e < 2e-2, (s, e)
        moved = max(moved, float((st[s][0].detach() - before[s]).abs().max()))
    assert moved > 1e-3      # the step really changed the function


def test_token_mask_gradients(dev):
    """enable_masking: masked stage-1 tokens take the mask token (maxvit_rnn.py:174-176); its gradient and all others
    match autograd through the oracle."""
    case = BACKBONE_CASES['ls_init_mask']
    spec = spec_of(case)
    params = bo.synth_params(spec, case['seed'], 'uniform')
    import rvt_b200
    m = rvt_b200.RNNDetector(make_cfg(spec))
    m.load_state_dict(params, strict=True)
    m = m.to(dev).train()
    xs = case_inputs(case, 2)
    b, h, w = case['batch'], case['height'], case['width']
    masks = [torch.from_numpy(np.random.RandomState(5 + t).uniform(size=(b, h // 4, w // 4)) < 0.2).to(dev) for t in range(2)]
    outs, st = [], None
    for x, mk in zip(xs, masks):
        o, st = m(x.to(dev), st, mk)
        outs.append(o)
    train_loss(outs, st).backward()
    po = {k: v.to(dev).requires_grad_(True) for k, v in params.items()}
    o_outs, o_st = [], None
    for x, mk in zip(xs, masks):
        o, o_st = bo.backbone_forward(x.to(dev).float(), o_st, po, spec, mk)
        o_outs.append(o)
    train_loss(o_outs, o_st).backward()
    worst = {k: float((p.grad - po[k].grad).double().norm() / po[k].grad.double().norm().clamp_min(1e-20))
             for k, p in m.named_parameters()}
    assert 'stages.0.mask_token' in worst
    bad = {k: v for k, v in worst.items() if v > GRAD_TOL}
    assert not bad, bad

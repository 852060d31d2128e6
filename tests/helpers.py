"""Shared test utilities: replaying golden cases through any backbone implementation."""
import os

import numpy as np
import torch

from oracle import backbone_oracle as bo
from tests.golden_configs import spec_of

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| — the '1e-3 relative' metric used throughout."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def replay_case(case, step_fn, device='cpu', input_dtype=torch.float32):
    """Drive `step_fn(x, states, mask) -> (feats, states)` over the case's seeded inputs,
    mimicking the harness' in-place state reset (modules/utils/detection.py:96-113).
    Yields (step, feats, states)."""
    spec = spec_of(case)
    b, h, w = case['batch'], case['height'], case['width']
    states = None
    for step in range(case['steps']):
        x = bo.synth_events_tensor(case['seed'] * 1000 + step, b, spec.input_channels, h, w)
        x = x.to(device=device, dtype=input_dtype)
        mask = None
        if spec.enable_masking:
            rs = np.random.RandomState(case['seed'] + 77 + step)
            mask = torch.from_numpy(rs.uniform(size=(b, h // 4, w // 4)) < 0.2).to(device)
        feats, states = step_fn(x, states, mask)
        if case.get('reset_at') == step:
            for (hh, cc) in states:
                assert hh.requires_grad is False and cc.requires_grad is False
                hh[0] = 0
                cc[0] = 0
        yield step, feats, states


def check_against_golden(name, case, step_fn, tol, device='cpu', input_dtype=torch.float32):
    """Compare states at the saved steps with the REFERENCE's outputs in tests/golden."""
    gold = np.load(os.path.join(GOLD, f'backbone_{name}.npz'))
    stride = case.get('sub', 1)
    save_steps = case.get('save_steps', [case['steps'] - 1])
    worst = 0.0
    tol_fn = tol if callable(tol) else (lambda _step: tol)
    for step, feats, states in replay_case(case, step_fn, device, input_dtype):
        if step not in save_steps:
            continue
        tol = tol_fn(step)
        for s in range(4):
            hh, cc = states[s]
            assert feats[s + 1].shape == hh.shape
            for tag, t in (('h', hh), ('c', cc)):
                got = t.detach().float().cpu().contiguous().reshape(-1)[::stride]
                ref = torch.from_numpy(gold[f'step{step}_{tag}{s}'])
                e = rel_err(got, ref)
                worst = max(worst, e)
                assert e <= tol, f'{name} step {step} stage {s} {tag}: rel err {e:.3e} > {tol}'
            ref_sum = float(gold[f'step{step}_h{s}_sum'])
            ref_abs = float(gold[f'step{step}_h{s}_abssum'])
            got_sum = float(hh.detach().double().sum().cpu())
            assert abs(got_sum - ref_sum) <= max(tol, 1e-6) * ref_abs, (name, step, s, got_sum, ref_sum)
    return worst


# ---------------------------------------------------------------------------------------------
# training-step parity: one deterministic scalar loss over an unrolled sequence (features of every
# step and stage, plus the final cell states), used identically for the reference (golden minting),
# the oracle and the CUDA path.  Sum-based so gradient magnitudes stay O(0.1-1) (no fp16 underflow).
# ---------------------------------------------------------------------------------------------
def train_loss(outs_per_step, final_states) -> torch.Tensor:
    loss = 0.0
    for t, feats in enumerate(outs_per_step):
        for s in range(1, 5):
            f = feats[s].float()
            loss = loss + (0.5 + 0.25 * t) / s * 0.5 * ((f - 0.1) ** 2).sum() / f.shape[0]
    for (_, c) in final_states:
        loss = loss + 0.1 * (c.float() ** 2).sum() / c.shape[0]
    return loss


GRAD_CASES = {
    # name -> (backbone case, unrolled steps)
    'tiny_p6': 3,
    'small_dh24': 2,
    'rvt_b_1mpx_bs3': 2,      # BASELINE configs[2] per-GPU shape (C up to 512, P = 60, bs 3)
}
GRAD_SUB = 29   # stride of the committed gradient subsamples


def grad_sub(name: str) -> int:
    """stride of the committed gradient subsample of a case (bigger models: sparser)"""
    return 211 if name == 'rvt_b_1mpx_bs3' else GRAD_SUB


def case_inputs(case, steps):
    spec = spec_of(case)
    return [bo.synth_events_tensor(case['seed'] * 1000 + t, case['batch'], spec.input_channels, case['height'], case['width'])
            for t in range(steps)]

"""Measure, on the B200, how far the REFERENCE's own mixed-precision run (`precision: 16`, config/general.yaml:6: the op
sequence under torch.autocast(float16), fp32 residual stream / states) sits from its pure-fp32 run, next to how
far rvt_b200 sits from the same fp32 run — per operator on identical inputs, and over a 21-step sequence with states carried.
The port of the reference op sequence is oracle/backbone_oracle.py (pinned to the reference by tests/golden).

  python profiles/amp_envelope.py [--batch 2] [--steps 21] [--out gpurun_out/amp_envelope.json]

Test infrastructure / measurement only (uses oracle/); nothing in rvt_b200/ imports it."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import backbone_oracle as bo                   # noqa: E402
from tests.golden_configs import BACKBONE_CASES, spec_of   # noqa: E402
from tests.test_host_cpu import make_cfg                   # noqa: E402


def errs(a, b):
    a, b = a.detach().double(), b.detach().double()
    return {'rel_max': float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)),
            'rel_l2': float((a - b).norm() / b.norm().clamp_min(1e-30))}


def amp():
    return torch.autocast('cuda', dtype=torch.float16)


def per_operator(case, dev):
    """Every operator fed the fp32 run's own input for it: AMP port vs fp32 port."""
    spec = spec_of(case)
    params = {k: v.to(dev) for k, v in bo.synth_params(spec, case['seed']).items()}
    rows = []
    states = None
    for step in range(2):
        x = bo.synth_events_tensor(case['seed'] * 1000 + step, case['batch'], 20, case['height'], case['width']).to(dev).float()
        taps = {}
        prev = states
        with torch.no_grad():
            _, states = bo.backbone_forward(x, prev, params, spec, taps=taps)
            cur = x
            for s in range(4):
                pre = f'stages.{s}.'
                f = spec.patch_size if s == 0 else 2
                with amp():
                    got = bo.downsample_cf2cl(cur, params, pre + 'downsample_cf2cl.', f, spec.overlap)
                rows.append((step, pre + 'downsample', errs(got.float(), taps[pre + 'downsample'])))
                xin = taps[pre + 'downsample']
                for kind, window in (('att_window', True), ('att_grid', False)):
                    bp = f'{pre}att_blocks.0.{kind}.'
                    t2 = {}
                    with amp():
                        bo.partition_attention_cl(xin, params, bp, spec.partition_size, window, spec.dim_head, spec.norm_eps, None, t2)
                    rows.append((step, bp + 'x_attn', errs(t2[bp + 'x_attn'].float(), taps[bp + 'x_attn'])))
                    # MLP half fed the fp32 x_attn
                    xa = taps[bp + 'x_attn']
                    with amp():
                        m = bo.mlp_branch(xa, params, bp, spec.norm_eps)
                        g2 = params.get(bp + 'ls2.gamma')
                        y = xa + (m * g2 if g2 is not None else m)
                    rows.append((step, bp + 'x_mlp', errs(y.float(), taps[bp + 'x_mlp'])))
                    xin = taps[bp + 'x_mlp']
                with amp():
                    h1, c1 = bo.dws_conv_lstm(taps[pre + 'pre_lstm'].permute(0, 3, 1, 2), None if prev is None else prev[s], params,
                                              pre + 'lstm.', spec.dws_conv, spec.dws_conv_only_hidden)
                rows.append((step, pre + 'lstm.h', errs(h1.float(), states[s][0])))
                rows.append((step, pre + 'lstm.c', errs(c1.float(), states[s][1])))
                cur = states[s][0]
    return rows


def sequence(case, batch, steps, dev):
    """fp32 port vs AMP port vs rvt_b200 over `steps` timesteps, states carried."""
    import rvt_b200
    spec = spec_of(case)
    cpu_params = bo.synth_params(spec, case['seed'])
    params = {k: v.to(dev) for k, v in cpu_params.items()}
    m = rvt_b200.build_recurrent_backbone(make_cfg(spec))
    m.load_state_dict(cpu_params, strict=True)
    m = m.to(dev).eval()
    st32 = st16 = sto = None
    rows = []
    report = sorted({0, 1, 4, 9, steps - 1})
    for t in range(steps):
        x8 = bo.synth_events_tensor(case['seed'] * 1000 + t, batch, 20, case['height'], case['width']).to(dev)
        with torch.no_grad():
            _, st32 = bo.backbone_forward(x8.float(), st32, params, spec)
            with amp():
                _, st16 = bo.backbone_forward(x8.float(), st16, params, spec)
            st16 = [(h.float(), c.float()) for h, c in st16]
            _, sto = m(x8.float(), sto)
        if t in report:
            for s in range(4):
                for tag, i in (('h', 0), ('c', 1)):
                    rows.append({'step': t, 'stage': s, 'state': tag,
                                 'amp_vs_fp32': errs(st16[s][i], st32[s][i]),
                                 'ours_vs_fp32': errs(sto[s][i], st32[s][i]),
                                 'ours_vs_amp': errs(sto[s][i], st16[s][i])})
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--steps', type=int, default=21)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'amp_envelope.json'))
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False            # the fp32 comparator is true fp32
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device('cuda:0')
    case = BACKBONE_CASES['rvt_b_1mpx']
    ops_rows = per_operator(case, dev)
    seq_rows = sequence(case, args.batch, args.steps, dev)
    worst_op = {'rel_max': max(r[2]['rel_max'] for r in ops_rows), 'rel_l2': max(r[2]['rel_l2'] for r in ops_rows)}
    last = [r for r in seq_rows if r['step'] == args.steps - 1]
    summary = {
        'per_operator_amp_vs_fp32_worst': worst_op,
        f'step{args.steps - 1}_worst_amp_vs_fp32': max(r['amp_vs_fp32']['rel_max'] for r in last),
        f'step{args.steps - 1}_worst_ours_vs_fp32': max(r['ours_vs_fp32']['rel_max'] for r in last),
        f'step{args.steps - 1}_worst_ours_vs_amp': max(r['ours_vs_amp']['rel_max'] for r in last),
    }
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, 'w') as f:
        json.dump({'config': f"RVT-B 1Mpx 384x640, batch {args.batch}, {args.steps} steps; per-operator: batch {case['batch']}, 2 steps",
                   'summary': summary, 'per_operator': ops_rows, 'sequence': seq_rows}, f, indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == '__main__':
    main()

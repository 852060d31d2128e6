"""GPU: every hot-path operator IN ISOLATION, fed the fp32 oracle's own input for that operator,
against the fp32 oracle's output of the same operator (oracle pinned to the reference by
tests/golden).  This is the north-star parity bar: 1e-3 relative per kernel.

  rel-L2  = ||a-b||_2 / ||b||_2          <= 1e-3   (asserted)
  rel-max = max|a-b| / max|b|            <= 5e-3   (asserted; the tail of fp16 operand rounding
                                                    over 1e5..1e6 outputs sits at ~4-5 sigma)
The CUDA operators use fp16 tensor-core operands with fp32 accumulation and fp32
LayerNorm/softmax/gates; the oracle is pure fp32."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import backbone_oracle as bo
from tests.golden_configs import BACKBONE_CASES, spec_of
from tests.helpers import rel_err
from tests.test_gpu_backbone import build_module

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_L2, TOL_MAX = 1e-3, 5e-3


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize('name', list(BACKBONE_CASES))
def test_each_operator_with_oracle_inputs(name):
    from rvt_b200 import ops
    case = BACKBONE_CASES[name]
    m, params, spec = build_module(case)
    dev = torch.device('cuda:0')
    packed = m._ensure_packed(dev)
    o_states = None
    rows = []

    def rec(step, key, got, ref):
        rows.append((step, key, rel_l2(got, ref), rel_err(got, ref)))

    steps = min(case['steps'], 2)
    for step in range(steps):
        x = bo.synth_events_tensor(case['seed'] * 1000 + step, case['batch'], 20, case['height'], case['width']).float()
        mask = None
        if spec.enable_masking:
            rs = np.random.RandomState(case['seed'] + 77 + step)
            mask = torch.from_numpy(rs.uniform(size=(case['batch'], case['height'] // 4, case['width'] // 4)) < 0.2)
        taps = {}
        prev = o_states
        with torch.no_grad():
            o_out, o_states = bo.backbone_forward(x, prev, params, spec, mask, taps=taps)
        cur, cur_nchw = x.to(dev), True
        for s, (st, pk) in enumerate(zip(m.stages, packed)):
            d = st.downsample_cf2cl
            pre = f'stages.{s}.'
            tm = mask.to(dev) if (mask is not None and s == 0) else None
            got = ops.downsample_cf2cl(cur, cur_nchw, pk['conv_w'], st.dim, d.kernel_size, d.factor, d.padding,
                                       pk['ds_ln_w'], pk['ds_ln_b'], token_mask=tm, mask_token=pk['mask_token'])
            rec(step, pre + 'downsample', got, taps[pre + 'downsample'])
            if s > 0:       # fp16 channels-last input (the LSTM's fp16 copy of h feeds the next stage's conv)
                got16 = ops.downsample_cf2cl(cur.half(), False, pk['conv_w'], st.dim, d.kernel_size, d.factor, d.padding,
                                             pk['ds_ln_w'], pk['ds_ln_b'])
                rec(step, pre + 'downsample(f16 in)', got16, taps[pre + 'downsample'])
            if s == 0:      # stem fast path (space-to-depth) and uint8 input must agree with the generic gather path
                s2d = torch.empty(cur.numel(), dtype=torch.float16, device=dev)
                got2 = ops.downsample_cf2cl(cur.to(torch.uint8), True, pk['conv_w_s2d'], st.dim, d.kernel_size, d.factor,
                                            d.padding, pk['ds_ln_w'], pk['ds_ln_b'], token_mask=tm,
                                            mask_token=pk['mask_token'], s2d_scratch=s2d)
                rec(step, pre + 'downsample(s2d,u8)', got2, taps[pre + 'downsample'])
                cu8 = cur.to(torch.uint8)
                if pk['conv_w_u8'] is not None and ops.stem_u8_ok(cu8, 20, d.kernel_size, d.factor, d.padding, cu8.shape[2:], st.dim):
                    got3 = ops.downsample_cf2cl(cu8, True, pk['conv_w_u8'], st.dim, d.kernel_size, d.factor, d.padding,
                                                pk['ds_ln_w'], pk['ds_ln_b'], token_mask=tm, mask_token=pk['mask_token'],
                                                stem_mode=2)
                    rec(step, pre + 'downsample(u8 smem patch)', got3, taps[pre + 'downsample'])
            xin = taps[pre + 'downsample']
            b, hh, ww, c = xin.shape
            for bi, blk in enumerate(pk['blocks']):
                tp = f"{pre}att_blocks.{bi // 2}.{'att_grid' if blk['grid'] else 'att_window'}."
                rows_s = ops.attention_scratch_rows(b, hh, ww, blk['part'])
                sq = torch.empty(rows_s * 3 * c, dtype=torch.float16, device=dev)
                so = torch.empty(rows_s * c, dtype=torch.float16, device=dev)
                sxn = torch.empty(max(rows_s, ((b * hh * ww + 127) // 128) * 128) * c, dtype=torch.float16, device=dev)
                xg = xin.to(dev).contiguous()
                ops.partition_attention_(xg, blk, sq, so, sxn)
                rec(step, tp + 'x_attn', xg, taps[tp + 'x_attn'])
                xg = taps[tp + 'x_attn'].to(dev).contiguous()
                sh = torch.empty(((b * hh * ww + 127) // 128) * 128 * blk['hidden'], dtype=torch.float16, device=dev)
                ops.mlp_block_(xg, blk, sh, sxn)
                rec(step, tp + 'x_mlp', xg, taps[tp + 'x_mlp'])
                xin = taps[tp + 'x_mlp']
            hp = cp = None
            if prev is not None and prev[s] is not None:
                hp, cp = (t.permute(0, 2, 3, 1).contiguous().to(dev) for t in prev[s])
            sxh = torch.empty(((b * hh * ww + 127) // 128) * 128 * 2 * c, dtype=torch.float16, device=dev)
            hg, cg = ops.dws_conv_lstm(taps[pre + 'pre_lstm'].to(dev).contiguous(), hp, cp, pk, st.lstm.ks, sxh)
            rec(step, pre + 'lstm.h', hg, o_states[s][0].permute(0, 2, 3, 1))
            rec(step, pre + 'lstm.c', cg, o_states[s][1].permute(0, 2, 3, 1))
            cur, cur_nchw = o_states[s][0].permute(0, 2, 3, 1).contiguous().to(dev), False
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    tag = os.environ.get('RVT_PARITY_TAG', '')            # tests/test_gpu_variants.py: keep the variants' numbers apart
    with open(os.path.join(ROOT, 'gpurun_out', f'op_parity_{name}{tag}.json'), 'w') as f:
        json.dump(rows, f, indent=1)
    bad = [r for r in rows if not (r[2] <= TOL_L2 and r[3] <= TOL_MAX)]
    worst = max(r[2] for r in rows), max(r[3] for r in rows)
    print(f'{name}: worst rel-L2 {worst[0]:.2e}, worst rel-max {worst[1]:.2e} over {len(rows)} operator outputs')
    assert not bad, f'{bad[0]} (worst L2 {worst[0]:.2e}, max {worst[1]:.2e})'

"""Training step of the recurrent backbone: autograd-visible forward + analytic backward on the CUDA library.

The reference trains by letting PyTorch autograd differentiate ``RNNDetector.forward`` across the
unrolled sequence (modules/detection.py:150-199: L timesteps, states carried, one loss, TBPTT with
``RNNStates.save_states_and_detach`` between batches).  Here every ``RNNDetectorStage.forward``
(maxvit_rnn.py:169-182) is ONE ``torch.autograd.Function`` whose forward runs the training-mode CUDA
operators (they additionally save the intermediates the gradient needs) and whose backward composes
the building blocks of include/rvt_b200.h "Training step".  Autograd only does the plumbing between
stages / timesteps (summing the gradients that reach a state from the next stage, the next timestep
and the detection head).

Parameter gradients never go through per-timestep tensors: the weight-gradient kernels ADD into
persistent fp32 accumulators (one flat buffer per model), so all unrolled timesteps accumulate in place.
A single ``_GradSink`` node per chained sequence owns the parameters as autograd inputs; it sits below
every stage node of the sequence, therefore runs last, turns the accumulators into parameter gradients
(LayerScale finishing, weight layouts) and returns them to autograd once — so ``param.grad``, DDP hooks
and GradScaler behave exactly as with the reference module.

Numerics: fp32 residual-stream / state gradients, fp16 gradient signals inside a branch and fp16
tensor-core operands with fp32 accumulation (the reference under ``precision: 16`` back-propagates fp16
through its Linear / conv layers the same way; use a GradScaler as the reference harness does).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import _lib, ops, packing

_ru = ops.round_up


class TrainEngine:
    """Per-model training state: packed weights (forward + transposed for the data gradients),
    the gradient accumulators and the sink bookkeeping."""

    def __init__(self, model):
        self.model = model
        self.params: List[torch.nn.Parameter] = [p for _, p in model.named_parameters()]
        self.names: List[str] = [n for n, _ in model.named_parameters()]
        self._packed = None
        self._packed_key = None
        self._acc: Optional[Dict[str, torch.Tensor]] = None
        self._acc_flat: Optional[torch.Tensor] = None
        self.dirty = False
        self.gen = 0
        self._zero_tok = None

    def invalidate(self):
        """Forget the packed weights (next forward re-packs; used before a CUDA-graph capture so the refresh is recorded)."""
        self._packed = self._packed_key = None

    # ------------------------------------------------------------------ packed weights
    def packed(self, device):
        m = self.model
        key = m._param_key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        L = _lib.lib()
        f32 = lambda t: None if t is None else t.detach().to(device=device, dtype=torch.float32).contiguous()
        plw = packing.pack_linear_weight
        out = []
        for s, st in enumerate(m.stages):
            c = st.dim
            d = st.downsample_cf2cl
            if st.lstm.dws_conv:
                raise NotImplementedError('rvt_b200 training: dws_conv=True is not built (released configs use False)')
            w = d.conv.weight.detach().to(device).float()
            cin_p = _ru(st.dim_in, 8)                                      # im2col channel groups of 8 (the stem's 20 -> 24)
            k = d.kernel_size * d.kernel_size * cin_p
            ldc = k
            w2 = torch.zeros(c, d.kernel_size, d.kernel_size, cin_p, device=device)
            w2[..., :st.dim_in] = w.permute(0, 2, 3, 1)
            w2 = w2.reshape(c, k)                                          # K order (ky, kx, ci)
            e = {
                'conv_w': packing.pack_conv_weight(w, channels_last_input=s > 0, bn=L.rvt_conv_tile_n(c)),
                'conv_w_u8': (packing.pack_stem_weight_u8(w) if s == 0 and d.kernel_size == 7 and d.factor == 4 else None),
                'conv_wT': (plw(w2.t().contiguous(), L.rvt_tile_n(ldc, c))
                            if (s > 0 and ldc == k and L.rvt_tile_n(ldc, c) > 0) else None),
                'k': k, 'ldc': ldc, 'cin_p': cin_p,
                'ds_ln_w': f32(getattr(d.norm, 'weight', None)), 'ds_ln_b': f32(getattr(d.norm, 'bias', None)),
                'mask_token': f32(st.mask_token.reshape(-1)) if st.mask_token is not None else None,
                'blocks': [],
            }
            for bi, pair in enumerate(st.att_blocks):
                for kind, att in (('att_window', pair.att_window), ('att_grid', pair.att_grid)):
                    sa, mlp = att.self_attn, att.mlp
                    fc1, fc2 = mlp.net[0][0], mlp.net[2]
                    hid = fc1.weight.shape[0]
                    wq, wp = sa.qkv.weight.detach().to(device).float(), sa.proj.weight.detach().to(device).float()
                    w1, w2_ = fc1.weight.detach().to(device).float(), fc2.weight.detach().to(device).float()
                    e['blocks'].append({
                        'prefix': f'stages.{s}.att_blocks.{bi}.{kind}.',
                        'grid': 0 if att.window else 1, 'map_mode': 1 if att.window else 2,
                        'part': att.partition_size, 'dh': att.dim_head, 'eps': att.eps, 'hidden': hid,
                        'n1_w': f32(getattr(att.norm1, 'weight', None)), 'n1_b': f32(getattr(att.norm1, 'bias', None)),
                        'wqkv': plw(wq, L.rvt_tile_n(3 * c, c)), 'bqkv': f32(getattr(sa.qkv, 'bias', None)),
                        'wproj': plw(wp, L.rvt_tile_n(c, c)), 'bproj': f32(getattr(sa.proj, 'bias', None)),
                        'g1': f32(getattr(att.ls1, 'gamma', None)),
                        'n2_w': f32(att.norm2.weight), 'n2_b': f32(att.norm2.bias),
                        'w1': plw(w1, L.rvt_tile_n(hid, c)), 'b1': f32(getattr(fc1, 'bias', None)),
                        'w2': plw(w2_, L.rvt_tile_n(c, hid)), 'b2': f32(getattr(fc2, 'bias', None)),
                        'g2': f32(getattr(att.ls2, 'gamma', None)),
                        # data-gradient GEMMs: dX = dY W  ==  linear with weight W^T
                        'wqkvT': plw(wq.t().contiguous(), L.rvt_tile_n(c, 3 * c)),
                        'wprojT': plw(wp.t().contiguous(), L.rvt_tile_n(c, c)),
                        'w1T': plw(w1.t().contiguous(), L.rvt_tile_n(c, hid)),
                        'w2T': plw(w2_.t().contiguous(), L.rvt_tile_n(hid, c)),
                        # fp32 copies for the LayerScale finishing
                        'wproj_f32': wp, 'w2_f32': w2_,
                    })
            wl = st.lstm.conv1x1.weight.detach().to(device).float().reshape(4 * c, 2 * c)
            lw, lb = packing.pack_lstm_weight(st.lstm.conv1x1.weight.to(device), st.lstm.conv1x1.bias.to(device), L.rvt_lstm_cw(c))
            e['lstm_w'], e['lstm_b'] = lw, lb
            e['lstm_wxT'] = plw(wl[:, :c].t().contiguous(), L.rvt_tile_n(c, 4 * c))
            e['lstm_whT'] = plw(wl[:, c:].t().contiguous(), L.rvt_tile_n(c, 4 * c))
            out.append(e)
        self._packed, self._packed_key = out, key
        return out

    # ------------------------------------------------------------------ accumulators
    def _acc_layout(self):
        m = self.model
        lay = []
        for s, st in enumerate(m.stages):
            c = st.dim
            d = st.downsample_cf2cl
            k = d.kernel_size * d.kernel_size * _ru(st.dim_in, 8)
            pre = f'stages.{s}.'
            lay += [(pre + 'conv.GT', (k, c)), (pre + 'conv.ln_w', (c,)), (pre + 'conv.ln_b', (c,))]
            if st.mask_token is not None:
                lay += [(pre + 'mask_token', (c,))]
            for bi, pair in enumerate(st.att_blocks):
                for kind, att in (('att_window', pair.att_window), ('att_grid', pair.att_grid)):
                    bp = f'{pre}att_blocks.{bi}.{kind}.'
                    hid = att.mlp.net[0][0].weight.shape[0]
                    lay += [(bp + 'n1_w', (c,)), (bp + 'n1_b', (c,)), (bp + 'qkv.G', (3 * c, c)), (bp + 'qkv.s', (3 * c,)),
                            (bp + 'proj.G0', (c, c)), (bp + 'proj.s0', (c,)), (bp + 'n2_w', (c,)), (bp + 'n2_b', (c,)),
                            (bp + 'fc1.G', (hid, c)), (bp + 'fc1.s', (hid,)), (bp + 'fc2.G0T', (hid, c)), (bp + 'fc2.s0', (c,))]
            lay += [(pre + 'lstm.G', (4 * c, 2 * c)), (pre + 'lstm.s', (4 * c,))]
        return lay

    def acc(self, device) -> Dict[str, torch.Tensor]:
        if self._acc is None or self._acc_flat.device != device:
            lay = self._acc_layout()
            sizes = []
            for _, shp in lay:
                n = 1
                for v in shp:
                    n *= v
                sizes.append(_ru(n, 4))                       # 16-byte aligned segments
            flat = torch.zeros(sum(sizes), dtype=torch.float32, device=device)
            self._acc, off = {}, 0
            for (name, shp), n in zip(lay, sizes):
                numel = 1
                for v in shp:
                    numel *= v
                self._acc[name] = flat[off:off + numel].view(*shp)
                off += n
            self._acc_flat = flat
        return self._acc

    # ------------------------------------------------------------------ drain: accumulators -> parameter gradients
    def drain(self, device) -> Dict[str, torch.Tensor]:
        """Linear finishing of the accumulated raw gradients; zeroes the accumulators.  Called by the sink."""
        grads: Dict[str, torch.Tensor] = {}
        if not self.dirty:
            return grads
        A = self.acc(device)
        pk = self._packed if self._packed is not None else self.packed(device)
        m = self.model
        for s, st in enumerate(m.stages):
            c = st.dim
            d = st.downsample_cf2cl
            pre = f'stages.{s}.'
            cin_p = _ru(st.dim_in, 8)
            gt = A[pre + 'conv.GT'].t()                              # [c, K], K order (ky, kx, ci padded to 8)
            g = gt.reshape(c, d.kernel_size, d.kernel_size, cin_p)[..., :st.dim_in].permute(0, 3, 1, 2)
            grads[pre + 'downsample_cf2cl.conv.weight'] = g.contiguous()
            if st.mask_token is not None:
                grads[pre + 'mask_token'] = A[pre + 'mask_token'].reshape(1, 1, 1, c).clone()
            if d.norm_affine:
                grads[pre + 'downsample_cf2cl.norm.weight'] = A[pre + 'conv.ln_w'].clone()
                grads[pre + 'downsample_cf2cl.norm.bias'] = A[pre + 'conv.ln_b'].clone()
            for blk in pk[s]['blocks']:
                bp = blk['prefix']
                if blk['n1_w'] is not None:
                    grads[bp + 'norm1.weight'] = A[bp + 'n1_w'].clone()
                    grads[bp + 'norm1.bias'] = A[bp + 'n1_b'].clone()
                grads[bp + 'self_attn.qkv.weight'] = A[bp + 'qkv.G'].clone()
                grads[bp + 'self_attn.qkv.bias'] = A[bp + 'qkv.s'].clone()
                for gname, G0n, s0n, wname, bname, w32, bias in (
                        ('ls1.gamma', 'proj.G0', 'proj.s0', 'self_attn.proj.weight', 'self_attn.proj.bias', blk['wproj_f32'], blk['bproj']),
                        ('ls2.gamma', 'fc2.G0', 'fc2.s0', 'mlp.net.2.weight', 'mlp.net.2.bias', blk['w2_f32'], blk['b2'])):
                    G0 = A[bp + 'fc2.G0T'].t() if G0n == 'fc2.G0' else A[bp + G0n]
                    s0 = A[bp + s0n]
                    gamma = blk['g1'] if gname == 'ls1.gamma' else blk['g2']
                    if gamma is not None:
                        # out = x + gamma * (a W^T + b):  dW = gamma[:,None] * G0,  db = gamma * s0,
                        # dgamma = sum_t dout * (a W^T + b) = rowsum(W * G0) + b * s0      (maxvit.py:45-53)
                        dg = (w32 * G0).sum(1)
                        if bias is not None:
                            dg = dg + bias * s0
                        grads[bp + gname] = dg
                        grads[bp + wname] = gamma[:, None] * G0
                        grads[bp + bname] = gamma * s0
                    else:
                        grads[bp + wname] = G0.contiguous().clone()
                        grads[bp + bname] = s0.clone()
                grads[bp + 'norm2.weight'] = A[bp + 'n2_w'].clone()
                grads[bp + 'norm2.bias'] = A[bp + 'n2_b'].clone()
                grads[bp + 'mlp.net.0.0.weight'] = A[bp + 'fc1.G'].clone()
                grads[bp + 'mlp.net.0.0.bias'] = A[bp + 'fc1.s'].clone()
            grads[pre + 'lstm.conv1x1.weight'] = A[pre + 'lstm.G'].reshape(4 * c, 2 * c, 1, 1).clone()
            grads[pre + 'lstm.conv1x1.bias'] = A[pre + 'lstm.s'].clone()
        self._acc_flat.zero_()
        self.dirty = False
        self.gen += 1
        return grads

    def zero_token_grad(self, device):
        if self._zero_tok is None or self._zero_tok.device != device:
            self._zero_tok = torch.zeros(1, device=device)
        return self._zero_tok

    # ------------------------------------------------------------------ one stage, forward
    def stage_forward(self, s: int, cur: torch.Tensor, cur_nchw: bool, hp, cp, token_mask=None):
        """Training-mode RNNDetectorStage.forward; returns (h_new, c_new, saved)."""
        m = self.model
        st = m.stages[s]
        d = st.downsample_cf2cl
        c = st.dim
        dev = cur.device
        pk = self._packed[s]                          # refreshed once per forward call (forward_train)
        L = _lib.lib()
        stream = torch.cuda.current_stream(dev).cuda_stream
        ptr = _lib.ptr
        if cur_nchw:
            b, cin, hin, win = cur.shape
        else:
            b, hin, win, cin = cur.shape
        vh, vw = m.pad_to_hw if (s == 0 and m.pad_to_hw is not None) else (hin, win)
        ks, stride, pad = d.kernel_size, d.factor, d.padding
        hh, ww = (vh + 2 * pad - ks) // stride + 1, (vw + 2 * pad - ks) // stride + 1
        n_tok = b * hh * ww
        n_pad = _ru(n_tok, 128)
        f16 = lambda n: torch.empty(n, dtype=torch.float16, device=dev)
        raw = torch.empty((b, hh, ww, c), dtype=torch.float32, device=dev)
        x = torch.empty_like(raw)
        conv_w, stem_mode = pk['conv_w'], 0
        if s == 0 and pk['conv_w_u8'] is not None and ops.stem_u8_ok(cur, cin, ks, stride, pad, (vh, vw), c):
            conv_w, stem_mode = pk['conv_w_u8'], 2
        if token_mask is not None:
            assert pk['mask_token'] is not None, 'No mask token present in this stage'
            token_mask = token_mask.to(device=dev, dtype=torch.uint8).contiguous()
            assert tuple(token_mask.shape) == (b, hh, ww)
        _lib.check(L.rvt_downsample_cf2cl_train(
            ptr(cur), ops._IN_DTYPES[cur.dtype], int(cur_nchw), b, cin, hin, win, ks, stride, pad, hh, ww, c, ptr(conv_w),
            ptr(pk['ds_ln_w']), ptr(pk['ds_ln_b']), 1e-5, ptr(token_mask), ptr(pk['mask_token']), ptr(x), ptr(raw), None,
            stem_mode, stream), 'downsample_cf2cl_train')
        saved = {'cur': cur, 'cur_nchw': cur_nchw, 'raw': raw, 'geom': (b, cin, hin, win, ks, stride, pad, hh, ww), 'blocks': [],
                 'token_mask': token_mask}
        for blk in pk['blocks']:
            rows = ops.attention_scratch_rows(b, hh, ww, blk['part'])
            hid = blk['hidden']
            qkv, o = f16(rows * 3 * c), f16(rows * c)
            xn1, xn2 = f16(rows * c), f16(n_pad * c)
            x1 = torch.empty_like(x)
            _lib.check(L.rvt_partition_attention_train(
                ptr(x), ptr(x1), b, hh, ww, c, blk['part'][0], blk['part'][1], blk['grid'], blk['dh'], ptr(blk['n1_w']),
                ptr(blk['n1_b']), blk['eps'], ptr(blk['wqkv']), ptr(blk['bqkv']), ptr(blk['wproj']), ptr(blk['bproj']),
                ptr(blk['g1']), ptr(qkv), ptr(o), ptr(xn1), stream), 'partition_attention_train')
            pre_, act = f16(n_pad * hid), f16(n_pad * hid)
            x2 = torch.empty_like(x)
            _lib.check(L.rvt_mlp_block_train(
                ptr(x1), ptr(x2), n_tok, c, hid, ptr(blk['n2_w']), ptr(blk['n2_b']), blk['eps'], ptr(blk['w1']), ptr(blk['b1']),
                ptr(blk['w2']), ptr(blk['b2']), ptr(blk['g2']), ptr(pre_), ptr(act), ptr(xn2), stream), 'mlp_block_train')
            saved['blocks'].append({'x_in': x, 'qkv': qkv, 'o': o, 'x_mid': x1, 'pre': pre_, 'act': act, 'rows': rows,
                                    'xn1': xn1, 'xn2': xn2})
            x = x2
        xh, gates = f16(n_pad * 2 * c), f16(n_tok * 4 * c)
        h_new, c_new = torch.empty_like(x), torch.empty_like(x)
        _lib.check(L.rvt_dws_conv_lstm_train(ptr(x), ptr(hp), ptr(cp), b, hh, ww, c, ptr(pk['lstm_w']), ptr(pk['lstm_b']),
                                             ptr(h_new), ptr(c_new), ptr(xh), ptr(gates), stream), 'dws_conv_lstm_train')
        saved.update({'xh': xh, 'gates': gates, 'shape': (b, hh, ww, c), 'x_lstm': x})
        return h_new, c_new, saved

    # ------------------------------------------------------------------ one stage, backward
    def stage_backward(self, s: int, saved, cp, c_new, dh, dc, need_in: bool, need_hp: bool, need_cp: bool):
        """Returns (d_cur or None, dh_prev or None, dc_prev or None); parameter gradients go to the accumulators."""
        b, hh, ww, c = saved['shape']
        dev = c_new.device
        pk = self._packed[s]                          # the weights of the forward this backward belongs to
        A = self.acc(dev)
        pre = f'stages.{s}.'
        n_tok, n_pad = b * hh * ww, _ru(b * hh * ww, 128)
        shape = (b, hh, ww, c)
        f16 = lambda n: torch.empty(n, dtype=torch.float16, device=dev)
        self.dirty = True
        # ---- Conv-LSTM (rnn.py:55-67)
        dpre = f16(n_pad * 4 * c)
        dc_prev = torch.empty(shape, dtype=torch.float32, device=dev) if need_cp else None
        ops.lstm_gates_bwd(saved['gates'], cp, c_new, dh, dc, n_tok, c, dpre, dc_prev)
        ops.gemm_tn(dpre, 4 * c, saved['xh'], 2 * c, n_tok, A[pre + 'lstm.G'], colsum1=A[pre + 'lstm.s'])
        dres = torch.empty(shape, dtype=torch.float32, device=dev)
        ops.linear_ex(dpre, n_tok, 4 * c, c, pk['lstm_wxT'], dres)
        dh_prev = None
        if need_hp:
            dh_prev = torch.empty(shape, dtype=torch.float32, device=dev)
            ops.linear_ex(dpre, n_tok, 4 * c, c, pk['lstm_whT'], dh_prev)
        del dpre
        # ---- attention / MLP blocks in reverse (maxvit.py:267-270)
        for blk, sv in zip(reversed(pk['blocks']), reversed(saved['blocks'])):
            bp, hid, eps = blk['prefix'], blk['hidden'], blk['eps']
            # MLP half: x2 = x1 + g2 * (fc2(gelu(fc1(norm2(x1)))) )
            d0, d1 = f16(n_pad * c), (f16(n_pad * c) if blk['g2'] is not None else None)
            ops.gather_cast(dres, 0, None, blk['g2'], d0, d1)
            ops.gemm_tn(sv['act'], hid, d0, c, n_tok, A[bp + 'fc2.G0T'], colsum2=A[bp + 'fc2.s0'])
            dpre_m = f16(n_pad * hid)
            ops.linear_ex(d1 if d1 is not None else d0, n_tok, c, hid, blk['w2T'], dpre_m, act=2, aux=sv['pre'])
            ops.gemm_tn(dpre_m, hid, sv['xn2'], c, n_tok, A[bp + 'fc1.G'], colsum1=A[bp + 'fc1.s'])
            dxn = f16(n_pad * c)
            ops.linear_ex(dpre_m, n_tok, hid, c, blk['w1T'], dxn)
            ops.ln_bwd(sv['x_mid'], dxn, shape, 0, None, blk['n2_w'], True, eps, dres, None, A[bp + 'n2_w'], A[bp + 'n2_b'])
            del dpre_m, dxn, d0, d1
            # attention half: x1 = x0 + g1 * proj(attn(partition(norm1(x0))))
            rows, mm, part = sv['rows'], blk['map_mode'], blk['part']
            d0, d1 = f16(rows * c), (f16(rows * c) if blk['g1'] is not None else None)
            ops.gather_cast(dres, mm, part, blk['g1'], d0, d1)
            ops.gemm_tn(d0, c, sv['o'], c, rows, A[bp + 'proj.G0'], colsum1=A[bp + 'proj.s0'])
            do = f16(rows * c)
            ops.linear_ex(d1 if d1 is not None else d0, rows, c, c, blk['wprojT'], do)
            groups_rows = b * (hh // part[0]) * (ww // part[1]) * _lib.lib().rvt_rows_per_group(part[0] * part[1])
            dqkv = f16(rows * 3 * c) if groups_rows == rows else torch.zeros(rows * 3 * c, dtype=torch.float16, device=dev)
            ops.attn_core_bwd(sv['qkv'], sv['o'], do, dqkv, shape, part, blk['dh'])
            do_ln = blk['n1_w'] is not None
            ops.gemm_tn(dqkv, 3 * c, sv['xn1'], c, rows, A[bp + 'qkv.G'], colsum1=A[bp + 'qkv.s'])
            dxn = f16(rows * c)
            ops.linear_ex(dqkv, rows, 3 * c, c, blk['wqkvT'], dxn)
            ops.ln_bwd(sv['x_in'] if do_ln else None, dxn, shape, mm, part, blk['n1_w'], do_ln, eps, dres, None,
                       A[bp + 'n1_w'] if do_ln else None, A[bp + 'n1_b'] if do_ln else None)
            del d0, d1, do, dqkv, dxn
        # ---- mask token (maxvit_rnn.py:174-176): masked tokens took the token's value -> their gradient goes to it
        if saved.get('token_mask') is not None:
            mk = saved['token_mask'].reshape(-1, 1).to(torch.float32)
            d2 = dres.view(-1, c)
            A[pre + 'mask_token'].add_((d2 * mk).sum(0))
            d2.mul_(1.0 - mk)
        # ---- downsample conv + LayerNorm (maxvit.py:174-178)
        bq, cin, hin, win, ks, stride, pad, _, _ = saved['geom']
        dy16 = f16(n_pad * c)
        ops.ln_bwd(saved['raw'], dres, shape, 0, None, pk['ds_ln_w'], True, 1e-5, None, dy16, A[pre + 'conv.ln_w'] if pk['ds_ln_w'] is not None else None,
                   A[pre + 'conv.ln_b'] if pk['ds_ln_w'] is not None else None)
        ldc = pk['ldc']
        col = f16(n_tok * ldc)
        src = saved['cur']
        if saved['cur_nchw']:
            # the stem's NCHW events -> channels-last fp16 (channels padded to a multiple of 8) so im2col is 16-byte vectors
            nhwc = f16(bq * hin * win * pk['cin_p'])
            ops.nchw_to_nhwc_f16(src, pk['cin_p'], nhwc)
            src = nhwc.view(bq, hin, win, pk['cin_p'])
        ops.im2col(src, False, ks, stride, pad, hh, ww, col)
        ops.gemm_tn(col, ldc, dy16, c, n_tok, A[pre + 'conv.GT'])
        d_cur = None
        if need_in:
            if pk['conv_wT'] is None:
                raise NotImplementedError('input gradient of this downsample geometry is not built')
            dcol = f16(n_pad * ldc)
            ops.linear_ex(dy16, n_tok, c, ldc, pk['conv_wT'], dcol)
            d_cur = torch.empty((bq, hin, win, cin), dtype=torch.float32, device=dev)
            ops.col2im(dcol, bq, cin, hin, win, ks, stride, pad, hh, ww, d_cur)
        return d_cur, dh_prev, dc_prev


class _GradSink(torch.autograd.Function):
    """Owns the parameters of one chained sequence in the autograd graph; see the module docstring."""

    @staticmethod
    def forward(ctx, engine: TrainEngine, *params):
        ctx.engine = engine
        ctx.device = next(p.device for p in params)
        return torch.zeros(1, device=ctx.device)

    @staticmethod
    def backward(ctx, _dtoken):
        eng = ctx.engine
        grads = eng.drain(ctx.device)
        # hand the gradients to autograd as views of ONE flat fp32 buffer (parameter order): AccumulateGrad keeps
        # them as .grad, so the data-parallel reduction is a single NCCL all-reduce over that buffer with no
        # flatten / unflatten copies (rvt_b200.sharding.allreduce_gradients).
        want = [(i, name, p) for i, (name, p) in enumerate(zip(eng.names, eng.params))
                if ctx.needs_input_grad[i + 1] and name in grads]
        out = [None] * len(eng.params)
        if want:
            flat = torch.empty(sum(p.numel() for _, _, p in want), dtype=torch.float32, device=ctx.device)
            off = 0
            for i, name, p in want:
                seg = flat[off:off + p.numel()].view(p.shape)
                seg.copy_(grads[name].reshape(p.shape))
                out[i] = seg if p.dtype == torch.float32 else seg.to(p.dtype)
                off += p.numel()
        return (None, *out)


class _StageFn(torch.autograd.Function):
    """One RNNDetectorStage.forward (maxvit_rnn.py:169-182) as a single autograd node."""

    @staticmethod
    def forward(ctx, engine: TrainEngine, s: int, cur_nchw: bool, token, cur, hp, cp, token_mask=None):
        h_new, c_new, saved = engine.stage_forward(s, cur, cur_nchw, hp, cp, token_mask)
        ctx.engine, ctx.s, ctx.saved = engine, s, saved
        ctx.save_for_backward(cp, c_new)
        return h_new, c_new

    @staticmethod
    def backward(ctx, dh, dc):
        eng = ctx.engine
        if ctx.saved is None:
            raise RuntimeError('rvt_b200: backward through the same stage a second time (retain_graph=True or two losses '
                               'backpropagated separately) is not supported: the saved fp16 intermediates are released after '
                               'the first pass and the weight-gradient accumulators already hold it; sum the losses and call '
                               'backward once, or re-run the forward')
        cp, c_new = ctx.saved_tensors
        dh = None if dh is None else dh.contiguous().float()
        dc = None if dc is None else dc.contiguous().float()
        need = ctx.needs_input_grad                    # (engine, s, cur_nchw, token, cur, hp, cp, token_mask)
        d_cur, dh_prev, dc_prev = eng.stage_backward(ctx.s, ctx.saved, cp, c_new, dh, dc, need[4], need[5], need[6])
        ctx.saved = None
        return None, None, None, eng.zero_token_grad(c_new.device), d_cur, dh_prev, dc_prev, None


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    """logical NCHW state -> contiguous [B, H, W, C] fp32, inside the autograd graph."""
    v = t.permute(0, 2, 3, 1)
    if v.dtype != torch.float32:
        v = v.float()
    return v.contiguous()


def forward_train(model, x: torch.Tensor, prev_states, token_mask):
    """RNNDetector.forward under grad mode (maxvit_rnn.py:93-105).

    ``model.train_wavefront = True`` (opt-in) enqueues stage s on its own CUDA stream (stage s of step t only waits for
    stage s-1 of step t and — by stream order — for its own step t-1), exactly like ``forward_sequence``: consecutive
    forward calls then overlap across stages, and because autograd replays every node's backward on the stream of its
    forward, the backward pass gets the mirrored wavefront for free.  The current stream waits for all stage streams
    before this function returns, so callers may use the outputs as usual."""
    eng: TrainEngine = model._train_engine()
    token = None
    for st in prev_states:
        if st is not None:
            t = getattr(st[0], '_rvt_token', None)
            if t is not None and t[1] == eng.gen and t[0].device == x.device:
                token = t[0]
                break
    if token is None:
        if eng.dirty:
            # a previous backward never reached its sink (it raised part-way, or its graph was dropped): its partial sums
            # must not leak into this sequence's gradients
            eng._acc_flat.zero_()
            eng.dirty = False
            eng.gen += 1
        token = _GradSink.apply(eng, *eng.params)
    x = model._prep_input(x)
    dev = x.device
    eng.packed(dev)                                   # re-pack if an optimizer step changed the parameters
    eng.acc(dev)                                      # accumulators exist before any stage stream may touch them
    n = model.num_stages
    main = torch.cuda.current_stream(dev)
    wavefront = bool(getattr(model, 'train_wavefront', False))
    if wavefront:
        if getattr(eng, '_streams', None) is None or eng._streams[0].device != dev:
            eng._streams = [torch.cuda.Stream(dev, priority=(-1 if s >= 2 else 0)) for s in range(n)]
        streams = eng._streams
        capturing = torch.cuda.is_current_stream_capturing()
        ready = torch.cuda.Event()
        ready.record(main)                             # inputs / states / re-packed weights prepared on the caller's stream
    states, output = [], {}
    cur, cur_nchw = x, True
    prev_done = None
    for s in range(n):
        if wavefront:
            streams[s].wait_event(ready)
            if prev_done is not None:
                streams[s].wait_event(prev_done)
            ctx_mgr = torch.cuda.stream(streams[s])
        else:
            ctx_mgr = _NullCtx()
        with ctx_mgr:
            hp = cp = None
            if prev_states[s] is not None:
                hp, cp = (_nhwc(t) for t in prev_states[s])
            h_new, c_new = _StageFn.apply(eng, s, cur_nchw, token, cur, hp, cp, token_mask if s == 0 else None)
            if wavefront:
                prev_done = torch.cuda.Event()
                prev_done.record(streams[s])
                if not capturing:
                    h_new.record_stream(main)
                    c_new.record_stream(main)
                    if s + 1 < n:
                        h_new.record_stream(streams[s + 1])
        h_nchw, c_nchw = h_new.permute(0, 3, 1, 2), c_new.permute(0, 3, 1, 2)
        h_nchw._rvt_token = (token, eng.gen)
        states.append((h_nchw, c_nchw))
        output[s + 1] = h_nchw
        cur, cur_nchw = h_new, False
    if wavefront:
        for st_ in streams:
            main.wait_stream(st_)
    return output, states


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False

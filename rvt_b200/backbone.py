"""Drop-in mirror of the reference's recurrent backbone module API.

``RNNDetector`` has the constructor, ``forward`` signature, ``get_stage_dims`` /
``get_strides`` and the exact ``state_dict`` key names/shapes of
``models/detection/recurrent_backbone/maxvit_rnn.py:23-105`` (reference), so released
checkpoints load strictly and ``YoloXDetector.forward_backbone`` (detector.py:34-41) can call
it unchanged.  All arithmetic runs in the sm_100a CUDA library through the C-ABI
(include/rvt_b200.h); PyTorch only owns device memory, streams and parameters.  There is no
CPU or PyTorch-op fallback: non-CUDA inputs raise.

Numerics contract: fp32 residual stream and (h, c) states as in the reference
under AMP; fp16 tensor-core operands with fp32 accumulation (the reference's
``precision: 16``); fp32 LayerNorm / softmax / gates.

Internal layout is channels-last; features and states are returned as logical-NCHW views with
channels-last strides — the same strides the reference's own (h, c) have — so the harness'
``state[idx] = 0`` in-place resets (modules/utils/detection.py:96-113) and feature indexing
keep working.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops, packing

LstmState = Optional[Tuple[torch.Tensor, torch.Tensor]]
LstmStates = List[LstmState]


def _cfg_get(cfg, key, default=None):
    if hasattr(cfg, 'get'):
        v = cfg.get(key, default)
        return default if v is None else v
    return getattr(cfg, key, default)


def _cfg(cfg, key):
    try:
        return cfg[key]
    except (TypeError, KeyError):
        return getattr(cfg, key)


class _Params(nn.Module):
    """A named bag of parameters (keeps the reference's state_dict key names)."""

    def __init__(self, **shapes):
        super().__init__()
        for name, shape in shapes.items():
            self.register_parameter(name, nn.Parameter(torch.empty(*shape)))


def _linear(n_out, n_in, bias=True):
    m = _Params(weight=(n_out, n_in), **({'bias': (n_out,)} if bias else {}))
    nn.init.kaiming_uniform_(m.weight, a=math.sqrt(5))
    if bias:
        bound = 1 / math.sqrt(n_in)
        nn.init.uniform_(m.bias, -bound, bound)
    return m


def _layernorm(c, affine=True):
    if not affine:
        return nn.Module()
    m = _Params(weight=(c,), bias=(c,))
    nn.init.ones_(m.weight)
    nn.init.zeros_(m.bias)
    return m


class _LayerScale(nn.Module):
    def __init__(self, dim, init_values):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))


class _Downsample(nn.Module):
    """ConvDownsampling_Cf2Cl parameters (maxvit.py:143-172)."""

    def __init__(self, dim_in, dim_out, factor, cfg):
        super().__init__()
        assert factor in (2, 4, 8)
        self.overlap = bool(_cfg_get(cfg, 'overlap', True))
        self.norm_affine = bool(_cfg_get(cfg, 'norm_affine', True))
        dtype = _cfg_get(cfg, 'type', 'patch')
        if dtype != 'patch':
            raise NotImplementedError(dtype)
        self.factor = factor
        self.kernel_size = (factor - 1) * 2 + 1 if self.overlap else factor
        self.padding = self.kernel_size // 2 if self.overlap else 0
        self.conv = _Params(weight=(dim_out, dim_in, self.kernel_size, self.kernel_size))
        nn.init.kaiming_uniform_(self.conv.weight, a=math.sqrt(5))
        self.norm = _layernorm(dim_out, self.norm_affine)


class _SelfAttention(nn.Module):
    def __init__(self, dim, dim_head, bias):
        super().__init__()
        self.qkv = _linear(3 * dim, dim, bias)
        self.proj = _linear(dim, dim, bias)


class _MLP(nn.Module):
    """Non-gated MLP parameters with the reference's Sequential nesting (maxvit.py:103-115):
    net.0.0 = Linear(C, 4C), net.1 = Dropout, net.2 = Linear(4C, C)."""

    def __init__(self, dim, ratio, bias):
        super().__init__()
        inner = int(dim * ratio)
        self.net = nn.ModuleList([nn.ModuleList([_linear(inner, dim, bias), nn.Identity()]), nn.Identity(),
                                  _linear(dim, inner, bias)])


class _PartitionAttention(nn.Module):
    """PartitionAttentionCl parameters (maxvit.py:193-250)."""

    def __init__(self, dim, window: bool, cfg, skip_first_norm: bool):
        super().__init__()
        if _cfg_get(cfg, 'use_torch_mha', False):
            raise NotImplementedError('use_torch_mha=True (TorchMHSAWrapperCl) is not built; released configs use False')
        if _cfg_get(cfg, 'mlp_gated', False):
            raise NotImplementedError('mlp_gated=True (GLU) is not built; released configs use False')
        act = _cfg_get(cfg, 'mlp_activation', 'gelu')
        if act != 'gelu':
            raise NotImplementedError(f'mlp_activation={act}; only gelu (released configs) is built')
        for k in ('drop_path', 'drop_mlp'):
            if float(_cfg_get(cfg, k, 0.0)) != 0.0:
                raise NotImplementedError(f'{k} > 0 (training-time regulariser) is not built')
        self.window = window
        self.eps = float(_cfg_get(cfg, 'norm_eps', 1e-5))
        self.dim_head = int(_cfg_get(cfg, 'dim_head', 32))
        part = _cfg(cfg, 'partition_size')
        self.partition_size = (part, part) if isinstance(part, int) else tuple(int(v) for v in part)
        assert len(self.partition_size) == 2
        bias = bool(_cfg_get(cfg, 'attention_bias', True))
        ls = float(_cfg_get(cfg, 'ls_init_value', 1e-5))
        self.norm1 = nn.Identity() if skip_first_norm else _layernorm(dim)
        self.self_attn = _SelfAttention(dim, self.dim_head, bias)
        self.ls1 = _LayerScale(dim, ls) if ls > 0 else nn.Identity()
        self.norm2 = _layernorm(dim)
        self.mlp = _MLP(dim, _cfg_get(cfg, 'mlp_ratio', 4), bool(_cfg_get(cfg, 'mlp_bias', True)))
        self.ls2 = _LayerScale(dim, ls) if ls > 0 else nn.Identity()


class _AttentionPair(nn.Module):
    def __init__(self, dim, skip_first_norm, cfg):
        super().__init__()
        self.att_window = _PartitionAttention(dim, True, cfg, skip_first_norm)
        self.att_grid = _PartitionAttention(dim, False, cfg, False)


class _ConvLSTM(nn.Module):
    """DWSConvLSTM2d parameters (rnn.py:11-34)."""

    def __init__(self, dim, dws_conv, only_hidden, ks, drop):
        super().__init__()
        if float(drop) != 0.0:
            raise NotImplementedError('drop_cell_update > 0 is not built')
        self.dim, self.dws_conv, self.only_hidden, self.ks = dim, bool(dws_conv), bool(only_hidden), int(ks)
        if self.dws_conv:
            d = dim if self.only_hidden else 2 * dim
            self.conv3x3_dws = _Params(weight=(d, 1, ks, ks), bias=(d,))
            nn.init.kaiming_uniform_(self.conv3x3_dws.weight, a=math.sqrt(5))
            nn.init.uniform_(self.conv3x3_dws.bias, -1 / ks, 1 / ks)
        else:
            self.conv3x3_dws = nn.Identity()
        self.conv1x1 = _Params(weight=(4 * dim, 2 * dim, 1, 1), bias=(4 * dim,))
        nn.init.kaiming_uniform_(self.conv1x1.weight, a=math.sqrt(5))
        nn.init.uniform_(self.conv1x1.bias, -1 / math.sqrt(2 * dim), 1 / math.sqrt(2 * dim))


class RNNDetectorStage(nn.Module):
    """Mirror of maxvit_rnn.py:130-182; NCHW in/out at the API."""

    def __init__(self, dim_in, stage_dim, spatial_downsample_factor, num_blocks, enable_token_masking,
                 T_max_chrono_init, stage_cfg):
        super().__init__()
        assert isinstance(num_blocks, int) and num_blocks > 0
        lstm_cfg = _cfg(stage_cfg, 'lstm')
        self.dim_in, self.dim = dim_in, stage_dim
        self.downsample_cf2cl = _Downsample(dim_in, stage_dim, spatial_downsample_factor, _cfg(stage_cfg, 'downsample'))
        self.att_blocks = nn.ModuleList([_AttentionPair(stage_dim, i == 0, _cfg(stage_cfg, 'attention'))
                                         for i in range(num_blocks)])
        self.lstm = _ConvLSTM(stage_dim, _cfg(lstm_cfg, 'dws_conv'), _cfg(lstm_cfg, 'dws_conv_only_hidden'),
                              _cfg(lstm_cfg, 'dws_conv_kernel_size'), _cfg_get(lstm_cfg, 'drop_cell_update', 0))
        if enable_token_masking:
            self.mask_token = nn.Parameter(torch.zeros(1, 1, 1, stage_dim))
            nn.init.normal_(self.mask_token, std=.02)
        else:
            self.mask_token = None


class RNNDetector(nn.Module):
    """B200-native ``MaxViTRNNDetector`` (reference maxvit_rnn.py:23-105)."""

    def __init__(self, mdl_config):
        super().__init__()
        in_channels = _cfg(mdl_config, 'input_channels')
        embed_dim = _cfg(mdl_config, 'embed_dim')
        dim_multiplier = tuple(_cfg(mdl_config, 'dim_multiplier'))
        num_blocks = tuple(_cfg(mdl_config, 'num_blocks'))
        t_max = tuple(_cfg(mdl_config, 'T_max_chrono_init'))      # read, unused (as in the reference)
        enable_masking = _cfg(mdl_config, 'enable_masking')
        num_stages = len(num_blocks)
        assert num_stages == 4
        assert isinstance(embed_dim, int)
        assert num_stages == len(dim_multiplier) == len(t_max)
        # the reference's optional torch.compile switch (maxvit_rnn.py:43-52) has no meaning here:
        # the step is already a fixed sequence of fused kernels (CUDA-graph capturable).
        patch_size = _cfg(_cfg(mdl_config, 'stem'), 'patch_size')
        self.stage_dims = [embed_dim * x for x in dim_multiplier]
        self.stages = nn.ModuleList()
        self.strides = []
        input_dim, stride = in_channels, 1
        for i, (nb, tm) in enumerate(zip(num_blocks, t_max)):
            f = patch_size if i == 0 else 2
            self.stages.append(RNNDetectorStage(input_dim, self.stage_dims[i], f, nb, enable_masking and i == 0,
                                                tm, _cfg(mdl_config, 'stage')))
            stride *= f
            self.strides.append(stride)
            input_dim = self.stage_dims[i]
        self.num_stages = num_stages
        self._packed = None
        self._packed_key = None
        self._scratch: Dict[str, torch.Tensor] = {}
        self._train = None
        # Optional: model input resolution.  When set, an un-padded event tensor (e.g. 360x640) is accepted and the
        # bottom/right zero padding the harness would add (utils/padding.py:29-44) is folded into the stem conv's bounds
        # checks.  Read from the backbone config's `in_res_hw` key when a caller put it there (the reference keeps it one
        # level up, config/modifier.py:28-34), else set by hand.
        hw = _cfg_get(mdl_config, 'in_res_hw', None)
        self.pad_to_hw: Optional[Tuple[int, int]] = tuple(int(v) for v in hw) if hw is not None else None
        # test hook: when a dict, forward() stores clones of the residual stream after every
        # operator (keys mirror oracle.backbone_oracle taps) so parity failures localise.
        self.debug_taps: Optional[Dict[str, torch.Tensor]] = None

    # ---- reference API -------------------------------------------------------------------
    def get_stage_dims(self, stages: Tuple[int, ...]) -> Tuple[int, ...]:
        idx = [x - 1 for x in stages]
        assert min(idx) >= 0 and max(idx) < len(self.stages), idx
        return tuple(self.stage_dims[i] for i in idx)

    def get_strides(self, stages: Tuple[int, ...]) -> Tuple[int, ...]:
        idx = [x - 1 for x in stages]
        assert min(idx) >= 0 and max(idx) < len(self.stages), idx
        return tuple(self.strides[i] for i in idx)

    def _train_engine(self):
        if getattr(self, '_train', None) is None:
            from . import train
            self._train = train.TrainEngine(self)
        return self._train

    # ---- packed weights ------------------------------------------------------------------
    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _ensure_packed(self, device):
        key = self._param_key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        L = _lib.lib()
        f32 = lambda t: None if t is None else t.detach().to(device=device, dtype=torch.float32).contiguous()
        packed = []
        for s, st in enumerate(self.stages):
            c = st.dim
            d = st.downsample_cf2cl
            e = {'conv_w': packing.pack_conv_weight(d.conv.weight.to(device), channels_last_input=s > 0,
                                                    bn=L.rvt_conv_tile_n(c)),
                 'conv_w_s2d': (packing.pack_stem_weight_s2d(d.conv.weight.to(device), d.factor)
                                if s == 0 and (d.factor * st.dim_in) % 8 == 0 else None),
                 'conv_w_u8': (packing.pack_stem_weight_u8(d.conv.weight.to(device))
                               if s == 0 and d.kernel_size == 7 and d.factor == 4 else None),
                 'ds_ln_w': f32(getattr(d.norm, 'weight', None)), 'ds_ln_b': f32(getattr(d.norm, 'bias', None)),
                 'mask_token': f32(st.mask_token.reshape(-1)) if st.mask_token is not None else None,
                 'blocks': []}
            for pair in st.att_blocks:
                for att in (pair.att_window, pair.att_grid):
                    sa, mlp = att.self_attn, att.mlp
                    fc1, fc2 = mlp.net[0][0], mlp.net[2]
                    bn1, bn2, _ = _lib.mlp_tiles(c, fc1.weight.shape[0])
                    if L.rvt_attention_is_fused(c, att.dim_head):
                        wqkv, bqkv = packing.pack_qkv_weight(sa.qkv.weight.to(device),
                                                             getattr(sa.qkv, 'bias', None), att.dim_head)
                        bqkv = bqkv.to(device)
                        wproj = packing.pack_linear_weight(sa.proj.weight.to(device), c)
                    else:
                        wqkv = packing.pack_linear_weight(sa.qkv.weight.to(device), L.rvt_tile_n(3 * c, c))
                        bqkv = f32(getattr(sa.qkv, 'bias', None))
                        wproj = packing.pack_linear_weight(sa.proj.weight.to(device), L.rvt_tile_n(c, c))
                    e['blocks'].append({
                        'grid': 0 if att.window else 1, 'part': att.partition_size, 'dh': att.dim_head, 'eps': att.eps,
                        'n1_w': f32(getattr(att.norm1, 'weight', None)), 'n1_b': f32(getattr(att.norm1, 'bias', None)),
                        'wqkv': wqkv, 'bqkv': bqkv, 'wproj': wproj,
                        'bproj': f32(getattr(sa.proj, 'bias', None)),
                        'g1': f32(getattr(att.ls1, 'gamma', None)),
                        'n2_w': f32(att.norm2.weight), 'n2_b': f32(att.norm2.bias),
                        'hidden': fc1.weight.shape[0],
                        'w1': packing.pack_linear_weight(fc1.weight.to(device), bn1),
                        'b1': f32(getattr(fc1, 'bias', None)),
                        'w2': packing.pack_linear_weight(fc2.weight.to(device), bn2),
                        'b2': f32(getattr(fc2, 'bias', None)),
                        'g2': f32(getattr(att.ls2, 'gamma', None)),
                    })
            lw, lb = packing.pack_lstm_weight(st.lstm.conv1x1.weight.to(device), st.lstm.conv1x1.bias.to(device),
                                              L.rvt_lstm_cw(c))
            e['lstm_w'], e['lstm_b'] = lw, lb
            if st.lstm.dws_conv:
                e['dw_w'] = packing.pack_dw_weight(st.lstm.conv3x3_dws.weight.to(device))
                e['dw_b'] = f32(st.lstm.conv3x3_dws.bias)
                e['dws_mode'] = 1 if st.lstm.only_hidden else 2
            else:
                e['dw_w'] = e['dw_b'] = None
                e['dws_mode'] = 0
            packed.append(e)
        self._packed, self._packed_key = packed, key
        return packed

    def _scratch_buf(self, name, numel, dtype, device):
        buf = self._scratch.get(name)
        if buf is None or buf.numel() < numel or buf.device != device or buf.dtype != dtype:
            buf = torch.empty(numel, dtype=dtype, device=device)
            self._scratch[name] = buf
        return buf

    # ---- forward ---------------------------------------------------------------------------
    @staticmethod
    def _as_nhwc_f32(t: torch.Tensor) -> torch.Tensor:
        """logical NCHW state -> contiguous [B,H,W,C] fp32 (free when already channels-last fp32)."""
        v = t.detach().permute(0, 2, 3, 1)
        if v.dtype != torch.float32 or not v.is_contiguous():
            v = v.to(torch.float32).contiguous()
        return v

    def forward(self, x: torch.Tensor, prev_states: Optional[LstmStates] = None,
                token_mask: Optional[torch.Tensor] = None):
        if not x.is_cuda:
            raise RuntimeError('rvt_b200.RNNDetector runs on CUDA (sm_100a) only; there is no CPU fallback')
        if prev_states is None:
            prev_states = [None] * self.num_stages
        assert len(prev_states) == self.num_stages
        assert x.dim() == 4
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training step (modules/detection.py:150-199): autograd-visible path, rvt_b200/train.py
            from . import train
            return train.forward_train(self, x, list(prev_states), token_mask)
        packed = self._ensure_packed(x.device)
        x = self._prep_input(x)
        states: LstmStates = []
        output: Dict[int, torch.Tensor] = {}
        cur, cur_nchw = x, True
        for s in range(self.num_stages):
            h_new, c_new, h16 = self._stage_step(s, packed[s], cur, cur_nchw, prev_states[s],
                                                 token_mask if s == 0 else None)
            h_nchw, c_nchw = h_new.permute(0, 3, 1, 2), c_new.permute(0, 3, 1, 2)
            states.append((h_nchw, c_nchw))
            output[s + 1] = h_nchw
            cur, cur_nchw = (h16 if h16 is not None else h_new), False
        return output, states

    def _prep_input(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError('rvt_b200.RNNDetector runs on CUDA (sm_100a) only; there is no CPU fallback')
        assert x.dim() == 4 and x.shape[1] == self.stages[0].dim_in
        if x.dtype not in (torch.uint8, torch.float16, torch.float32):
            x = x.to(torch.float32)
        return x.contiguous()

    def _stage_step(self, s: int, pk: dict, cur: torch.Tensor, cur_nchw: bool, prev_state: LstmState,
                    token_mask: Optional[torch.Tensor], want_h16: bool = True, h_out: Optional[torch.Tensor] = None):
        """One RNNDetectorStage.forward (maxvit_rnn.py:169-182) enqueued on the CURRENT stream:
        downsample(+LN) -> [window block, grid block] x num_blocks -> Conv-LSTM.  Scratch buffers are
        per stage, so different stages may run concurrently on different streams."""
        st = self.stages[s]
        d = st.downsample_cf2cl
        c = st.dim
        dev = cur.device
        b = cur.shape[0]
        taps = self.debug_taps
        conv_w, s2d, stem_mode = pk['conv_w'], None, 0
        vhw = self.pad_to_hw if (s == 0 and self.pad_to_hw is not None) else (cur.shape[2], cur.shape[3])
        if s == 0 and pk['conv_w_u8'] is not None and ops.stem_u8_ok(cur, st.dim_in, d.kernel_size, d.factor, d.padding, vhw, c):
            conv_w, stem_mode = pk['conv_w_u8'], 2
        elif s == 0 and pk['conv_w_s2d'] is not None:
            vw = self.pad_to_hw[1] if self.pad_to_hw is not None else cur.shape[3]
            if ops.stem_uses_s2d(st.dim_in, d.factor, d.kernel_size, d.padding, vw):
                conv_w = pk['conv_w_s2d']
                s2d = self._scratch_buf('s2d', b * cur.shape[2] * vw * st.dim_in, torch.float16, dev)
        if s == 0 and token_mask is not None:
            assert st.mask_token is not None, 'No mask token present in this stage'
        split_ws = None
        if s > 0 and c >= 256:
            n_out = b * (cur.shape[1] // d.factor) * (cur.shape[2] // d.factor)
            splits = _lib.lib().rvt_conv_split_k(n_out, c, st.dim_in * d.kernel_size ** 2)
            if splits > 1:
                split_ws = self._scratch_buf(f'convws{s}', splits * n_out * c, torch.float32, dev)
        xs = ops.downsample_cf2cl(
            cur, cur_nchw, conv_w, c, d.kernel_size, d.factor, d.padding, pk['ds_ln_w'], pk['ds_ln_b'],
            virtual_hw=self.pad_to_hw if s == 0 else None,
            token_mask=token_mask if s == 0 else None, mask_token=pk['mask_token'], s2d_scratch=s2d,
            stem_mode=stem_mode, split_ws=split_ws)
        _, hh, ww, _ = xs.shape
        n_tok = b * hh * ww
        if taps is not None:
            taps[f'stages.{s}.downsample'] = xs.clone()
        for bi, blk in enumerate(pk['blocks']):
            tap_prefix = f"stages.{s}.att_blocks.{bi // 2}.{'att_grid' if blk['grid'] else 'att_window'}."
            rows = ops.attention_scratch_rows(b, hh, ww, blk['part'])
            sq = self._scratch_buf(f'qkv{s}', rows * 3 * c, torch.float16, dev)
            so = self._scratch_buf(f'o{s}', rows * c, torch.float16, dev)
            sx = self._scratch_buf(f'xn{s}', max(rows, ((n_tok + 127) // 128) * 128) * c, torch.float16, dev)
            ops.partition_attention_(xs, blk, sq, so, sx)
            if taps is not None:
                taps[tap_prefix + 'x_attn'] = xs.clone()
            sh = self._scratch_buf(f'hidden{s}', ((n_tok + 127) // 128) * 128 * blk['hidden'], torch.float16, dev)
            ops.mlp_block_(xs, blk, sh, sx)
            if taps is not None:
                taps[tap_prefix + 'x_mlp'] = xs.clone()
        hp = cp = None
        if prev_state is not None:
            hp, cp = (self._as_nhwc_f32(t) for t in prev_state)
            assert hp.shape == xs.shape and cp.shape == xs.shape
        sxh = None
        if c >= 128 and pk['dws_mode'] == 0:
            sxh = self._scratch_buf(f'xh{s}', ((n_tok + 127) // 128) * 128 * 2 * c, torch.float16, dev)
        # fp16 copy of h_t for the next stage's im2col loader (half the bytes, no conversion)
        h16 = torch.empty(xs.shape, dtype=torch.float16, device=dev) if (want_h16 and s + 1 < self.num_stages) else None
        h_new, c_new = ops.dws_conv_lstm(xs, hp, cp, pk, st.lstm.ks, sxh, h16, h_out=h_out)
        return h_new, c_new, h16

    @torch.no_grad()
    def forward_sequence(self, xs, prev_states: Optional[LstmStates] = None, token_masks=None,
                         wavefront: bool = True, input_ready=None, reset_mask: Optional[torch.Tensor] = None,
                         select: Optional[torch.Tensor] = None):
        """Run consecutive timesteps (extension of the reference API; the reference's time loop lives
        in the harness, modules/detection.py:131-148 / :231-243).

        xs: sequence (list or [L, B, C, H, W] tensor) of event tensors.  Returns
        ([{1..4: feat} per step], final states) exactly as L chained ``forward`` calls would.

        wavefront=True exploits the recurrence structure: stage s at step t depends only on
        (s-1, t) and (s, t-1), so the four stages run on four CUDA streams, stage s working on step
        t while stage s-1 already works on t+1.  The small-grid kernels of the wide late stages then
        overlap the big grids of the early stages instead of leaving SMs idle.

        input_ready: optional per-step CUDA events (e.g. of host->device copies on a copy stream) the
        first stage waits on.  After the call ``self.last_step_events[t]`` is the event recorded when
        step t's last stage finished (lets a caller overlap device->host reads of step t).

        Harness glue on the device, so one captured graph serves a whole batch of the reference's
        ``training_step`` / ``_val_test_step_impl`` time loop (modules/detection.py:117-159, :217-255):

        reset_mask: bool / uint8 [B] device tensor = the batch's ``is_first_sample``: the given ``prev_states`` are zeroed IN PLACE
        for those samples before step 0, exactly ``RNNStates.reset`` (modules/utils/detection.py:96-113).

        select: int32 [S] device tensor of ``t * B + b`` for every labelled (step, sample) pair in harness order (t ascending,
        then ``valid_batch_indices``), negative = unused slot.  When given, a third value is returned:
        ``{stage: [S, C, H, W]}`` = what ``BackboneFeatureSelector.get_batched_backbone_features()`` would hold (:24-46);
        refill ``select`` in place before each graph replay and slice ``[:n_selected]``.
        """
        L = len(xs)
        if prev_states is None:
            prev_states = [None] * self.num_stages
        assert len(prev_states) == self.num_stages
        if L == 0:
            return [], list(prev_states)
        dev = xs[0].device
        packed = self._ensure_packed(dev)
        main = torch.cuda.current_stream(dev)
        n = self.num_stages
        if wavefront:
            if getattr(self, '_streams', None) is None or self._streams[0].device != dev:
                # later stages = small grids on the critical recurrence chain: give them scheduling priority so their
                # CTAs are placed as soon as an SM slot frees instead of queueing behind the big early-stage grids
                mode = os.environ.get('RVT_STREAM_PRIO', '2')
                prio = [0] * n
                if mode == '1':
                    prio = [-s for s in range(n)]
                elif mode == '2':
                    prio = [0, 0] + [-1] * (n - 2)
                elif mode == '3':
                    prio = [-(n - 1 - s) for s in range(n)]
                self._streams = [torch.cuda.Stream(dev, priority=p) for p in prio]
            streams = self._streams
            for st_ in streams:
                st_.wait_stream(main)
        else:
            streams = [main] * n
        state = list(prev_states)
        if reset_mask is not None:
            assert reset_mask.device == dev and reset_mask.dtype in (torch.bool, torch.uint8)
            reset_mask = reset_mask.contiguous()
            for s in range(n):
                if state[s] is None:
                    continue
                with torch.cuda.stream(streams[s]):
                    hv, cv = (self._as_nhwc_f32(t_) for t_ in state[s])        # views of the caller's tensors when channels-last fp32
                    ops.state_reset_(hv, cv, reset_mask)
                    state[s] = (hv.permute(0, 3, 1, 2), cv.permute(0, 3, 1, 2))
        seq_feats = None                     # select: every step's h_t of a stage lands in one [L, B, H, W, C] buffer
        outs = [dict() for _ in range(L)]
        feats_prev = [None] * L              # output of stage s-1 per step (channels-last)
        done = [[None] * L for _ in range(n)]
        capturing = torch.cuda.is_current_stream_capturing()
        # Every tensor that crosses streams (the fp16 copy of h_t that stage s+1 reads on ITS stream) is held until the
        # call returns: under CUDA-graph capture record_stream() is not available to defer the allocator's reuse of the
        # block, and a block recycled by stage s at step t+1 while stage s+1 (a full step behind) still reads it would be
        # a silent race in the replayed graph.
        keep_alive = []
        for t in range(L):
            for s in range(n):
                with torch.cuda.stream(streams[s]):
                    if s == 0:
                        if input_ready is not None:
                            streams[0].wait_event(input_ready[t])
                        x_t = self._prep_input(xs[t])      # any dtype / layout conversion runs on stage 0's stream
                    if wavefront and s > 0:
                        streams[s].wait_event(done[s - 1][t])
                    cur, nchw = (x_t, True) if s == 0 else (feats_prev[t], False)
                    tm = token_masks[t] if (token_masks is not None and s == 0) else None
                    h_dst = None
                    if select is not None:
                        if seq_feats is None:
                            seq_feats = [None] * n
                        if seq_feats[s] is None:
                            bsz = x_t.shape[0]
                            vh, vw = self.pad_to_hw if self.pad_to_hw is not None else (x_t.shape[2], x_t.shape[3])
                            seq_feats[s] = torch.empty((L, bsz, vh // self.strides[s], vw // self.strides[s], self.stage_dims[s]),
                                                       dtype=torch.float32, device=dev)
                        h_dst = seq_feats[s][t]
                    timeline = getattr(self, 'debug_timeline', None)      # profiling aid (profiles/wavefront_timeline.py), eager only
                    if timeline is not None:
                        ev0 = torch.cuda.Event(enable_timing=True)
                        ev0.record(streams[s])
                    h_new, c_new, h16 = self._stage_step(s, packed[s], cur, nchw, state[s], tm, h_out=h_dst)
                    if wavefront or s == n - 1 or timeline is not None:
                        ev = torch.cuda.Event(enable_timing=timeline is not None)
                        ev.record(streams[s])
                        done[s][t] = ev
                        if timeline is not None:
                            timeline.append((s, t, ev0, ev))
                    if wavefront and not capturing:
                        # eager mode: tell the caching allocator about the cross-stream consumers
                        if s + 1 < n:
                            (h16 if h16 is not None else h_new).record_stream(streams[s + 1])
                        h_new.record_stream(main)
                        c_new.record_stream(main)
                    feats_prev[t] = h16 if h16 is not None else h_new
                    keep_alive.append((x_t, feats_prev[t], h_new, c_new))
                    state[s] = (h_new.permute(0, 3, 1, 2), c_new.permute(0, 3, 1, 2))
                    outs[t][s + 1] = state[s][0]
        self.last_step_events = done[n - 1]
        selected = None
        if select is not None:
            assert select.device == dev and select.dtype == torch.int32
            selected = {}
            for s in range(n):
                with torch.cuda.stream(streams[s]):
                    buf = seq_feats[s]
                    sel = ops.gather_rows(buf.view((buf.shape[0] * buf.shape[1],) + tuple(buf.shape[2:])), select.contiguous())
                    if wavefront and not capturing:
                        sel.record_stream(main)
                    selected[s + 1] = sel.permute(0, 3, 1, 2)
        if wavefront:
            for st_ in streams:
                main.wait_stream(st_)
        del keep_alive          # all streams have been joined into `main`: same-stream reuse from here on is ordered
        return (outs, state) if select is None else (outs, state, selected)


def build_recurrent_backbone(backbone_cfg):
    """Mirror of models/detection/recurrent_backbone/__init__.py:6-11."""
    if _cfg(backbone_cfg, 'name') == 'MaxViTRNN':
        return RNNDetector(backbone_cfg)
    raise NotImplementedError

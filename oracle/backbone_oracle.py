"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path
(rvt_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
``--impl reference`` legs may use it, and only as the checker / CPU baseline.

CPU fp32 restatement of RVT's per-timestep recurrent backbone forward, written
as plain functions over a flat ``{state_dict key: tensor}`` mapping (the
reference's own key names).  Each function cites the reference
code it restates (paths relative to /root/reference).

Pinning: ``oracle/make_golden.py`` imports the *real* reference in the build
container, checks this restatement against it to <=2e-5 max-abs on every golden
configuration, and commits the reference's outputs under tests/golden/.
``tests/test_oracle_golden.py`` re-checks the restatement against those
fixtures wherever the suite runs.  The reference itself ships no tests or
golden vectors, so these minted vectors are the pin.

``operand_dtype`` (default None = pure fp32) optionally rounds every GEMM /
conv operand to fp16 or bf16 (fp32 accumulate), mimicking the rounding points
of the CUDA path; it is a debugging aid for localising kernel bugs, the parity
bar itself is always the pure-fp32 oracle / the reference goldens.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Optional[Tuple[Tensor, Tensor]]


@dataclass
class BackboneSpec:
    """Static hyper-parameters of the backbone (what the reference reads from
    ``mdl_config``: models/detection/recurrent_backbone/maxvit_rnn.py:28-33,144-160
    and models/layers/maxvit/maxvit.py:157-158,201-213)."""
    input_channels: int = 20
    embed_dim: int = 64
    dim_multiplier: Sequence[int] = (1, 2, 4, 8)
    num_blocks: Sequence[int] = (1, 1, 1, 1)
    patch_size: int = 4
    overlap: bool = True
    partition_size: Tuple[int, int] = (6, 10)
    dim_head: int = 32
    norm_eps: float = 1e-5
    ls_init_value: float = 1e-5
    dws_conv: bool = False
    dws_conv_only_hidden: bool = True
    dws_conv_kernel_size: int = 3
    enable_masking: bool = False
    stage_dims: List[int] = field(init=False)

    def __post_init__(self):
        self.stage_dims = [self.embed_dim * m for m in self.dim_multiplier]


def _round(t: Tensor, operand_dtype) -> Tensor:
    if operand_dtype is None:
        return t
    return t.to(operand_dtype).to(torch.float32)


def _linear(x, w, b, od):
    return F.linear(_round(x, od), _round(w, od), b)


# ---------------------------------------------------------------------------
# models/layers/maxvit/maxvit.py:143-178  ConvDownsampling_Cf2Cl
# ---------------------------------------------------------------------------
def downsample_cf2cl(x_nchw: Tensor, p: Dict[str, Tensor], prefix: str, factor: int,
                     overlap: bool, eps: float = 1e-5, od=None) -> Tensor:
    """Strided conv (no bias; overlapping kernel (f-1)*2+1, pad k//2 — maxvit.py:160-171)
    -> NHWC (maxvit.py:176) -> LayerNorm over C, eps 1e-5 (maxvit.py:172,177)."""
    w = p[prefix + 'conv.weight']
    k = w.shape[-1]
    pad = k // 2 if overlap else 0
    assert k == ((factor - 1) * 2 + 1 if overlap else factor)
    y = F.conv2d(_round(x_nchw, od), _round(w, od), None, stride=factor, padding=pad)
    y = y.permute(0, 2, 3, 1)
    c = y.shape[-1]
    nw, nb = p.get(prefix + 'norm.weight'), p.get(prefix + 'norm.bias')
    return F.layer_norm(y, (c,), nw, nb, eps)


# ---------------------------------------------------------------------------
# models/layers/maxvit/maxvit.py:273-304  window / grid partition + reverse
# ---------------------------------------------------------------------------
def partition_index(h: int, w: int, part: Tuple[int, int], window: bool) -> Tensor:
    """Flat token index (y*w+x) for every (group, position) pair, shape [nG, P].

    window (maxvit.py:273-279): group (gy,gx) holds the contiguous ph x pw block.
    grid   (maxvit.py:290-296): group (gy,gx) in (h/ph, w/pw) holds the dilated
    lattice  y = py*(h/ph)+gy,  x = px*(w/pw)+gx.
    """
    ph, pw = part
    assert h % ph == 0 and w % pw == 0
    ny, nx = h // ph, w // pw
    gy = torch.arange(ny).view(ny, 1, 1, 1)
    gx = torch.arange(nx).view(1, nx, 1, 1)
    py = torch.arange(ph).view(1, 1, ph, 1)
    px = torch.arange(pw).view(1, 1, 1, pw)
    if window:
        y, x = gy * ph + py, gx * pw + px
    else:
        y, x = py * ny + gy, px * nx + gx
    return (y * w + x).reshape(ny * nx, ph * pw)


def self_attention(xp: Tensor, p: Dict[str, Tensor], prefix: str, dim_head: int, od=None) -> Tensor:
    """maxvit.py:343-354 SelfAttentionCl on partitioned tokens xp [G, P, C].
    qkv rows are per-head interleaved [q_h | k_h | v_h] (the view/chunk at :347)."""
    g, pp, c = xp.shape
    nh = c // dim_head
    qkv = _linear(xp, p[prefix + 'qkv.weight'], p.get(prefix + 'qkv.bias'), od)
    qkv = qkv.view(g, pp, nh, 3, dim_head)
    q, k, v = (qkv[:, :, :, i].transpose(1, 2) for i in range(3))  # [G, nh, P, dh]
    s = torch.matmul(_round(q, od), _round(k, od).transpose(-1, -2)) * (dim_head ** -0.5)
    a = torch.softmax(s, dim=-1)
    o = torch.matmul(_round(a, od), _round(v, od))  # [G, nh, P, dh]
    o = o.transpose(1, 2).reshape(g, pp, c)
    return _linear(o, p[prefix + 'proj.weight'], p.get(prefix + 'proj.bias'), od)


def attention_branch(x: Tensor, p, prefix, part, window, dim_head, eps, od=None) -> Tensor:
    """norm1 -> partition -> attention -> reverse (maxvit.py:252-265, :268 sans ls1/residual).
    norm1 is Identity when its parameters are absent (skip_first_norm, maxvit.py:234)."""
    b, h, w, c = x.shape
    if prefix + 'norm1.weight' in p:
        xn = F.layer_norm(x, (c,), p[prefix + 'norm1.weight'], p[prefix + 'norm1.bias'], eps)
    else:
        xn = x
    idx = partition_index(h, w, part, window)            # [nG, P]
    flat = xn.reshape(b, h * w, c)
    xp = flat[:, idx]                                      # [B, nG, P, C]
    ng, pp = idx.shape
    yp = self_attention(xp.reshape(b * ng, pp, c), p, prefix + 'self_attn.', dim_head, od)
    out = torch.empty_like(flat)
    out[:, idx.reshape(-1)] = yp.reshape(b, ng * pp, c).to(out.dtype)   # (dtype cast only matters under autocast)
    return out.reshape(b, h, w, c)


def mlp_branch(x: Tensor, p, prefix, eps, od=None) -> Tensor:
    """norm2 -> Linear(C,4C) -> exact-erf GELU -> Linear(4C,C)  (maxvit.py:110-118,241,269)."""
    c = x.shape[-1]
    xn = F.layer_norm(x, (c,), p[prefix + 'norm2.weight'], p[prefix + 'norm2.bias'], eps)
    hdn = F.gelu(_linear(xn, p[prefix + 'mlp.net.0.0.weight'], p.get(prefix + 'mlp.net.0.0.bias'), od))
    return _linear(hdn, p[prefix + 'mlp.net.2.weight'], p.get(prefix + 'mlp.net.2.bias'), od)


def partition_attention_cl(x: Tensor, p, prefix, part, window, dim_head, eps, od=None,
                           taps: Optional[dict] = None) -> Tensor:
    """maxvit.py:267-270: x + ls1(attn(...)); x + ls2(mlp(norm2(x))).  LayerScale
    (maxvit.py:45-53) is Identity when gamma is absent (ls_init_value <= 0, :238)."""
    a = attention_branch(x, p, prefix, part, window, dim_head, eps, od)
    if taps is not None:
        taps[prefix + 'attn_branch'] = a
    g1 = p.get(prefix + 'ls1.gamma')
    x = x + (a * g1 if g1 is not None else a)
    m = mlp_branch(x, p, prefix, eps, od)
    g2 = p.get(prefix + 'ls2.gamma')
    y = x + (m * g2 if g2 is not None else m)
    if taps is not None:
        taps[prefix + 'mlp_branch'] = m
        taps[prefix + 'x_attn'] = x
        taps[prefix + 'x_mlp'] = y
    return y


# ---------------------------------------------------------------------------
# models/layers/rnn.py:36-69  DWSConvLSTM2d
# ---------------------------------------------------------------------------
def dws_conv_lstm(x_nchw: Tensor, state: State, p, prefix, dws_conv: bool,
                  only_hidden: bool, od=None) -> Tuple[Tensor, Tensor]:
    c = x_nchw.shape[1]
    if state is None:                                   # rnn.py:43-47
        h0 = torch.zeros_like(x_nchw)
        c0 = torch.zeros_like(x_nchw)
    else:
        h0, c0 = state
    if dws_conv:
        dw, db = p[prefix + 'conv3x3_dws.weight'], p[prefix + 'conv3x3_dws.bias']
        pad = dw.shape[-1] // 2
    if dws_conv and only_hidden:                        # rnn.py:50-51
        h0 = F.conv2d(h0, dw, db, padding=pad, groups=h0.shape[1])
    xh = torch.cat((x_nchw, h0), dim=1)                 # rnn.py:52
    if dws_conv and not only_hidden:                    # rnn.py:53-54
        xh = F.conv2d(xh, dw, db, padding=pad, groups=xh.shape[1])
    mix = F.conv2d(_round(xh, od), _round(p[prefix + 'conv1x1.weight'], od),
                   p[prefix + 'conv1x1.bias'])          # rnn.py:55
    f, i, o = (torch.sigmoid(mix[:, j * c:(j + 1) * c]) for j in range(3))  # rnn.py:57-61
    g = torch.tanh(mix[:, 3 * c:])                      # rnn.py:64
    c1 = f * c0 + i * g                                 # rnn.py:66
    h1 = o * torch.tanh(c1)                             # rnn.py:67
    return h1, c1


# ---------------------------------------------------------------------------
# models/detection/recurrent_backbone/maxvit_rnn.py:93-105,169-182
# ---------------------------------------------------------------------------
def stage_forward(x_nchw, state, token_mask, p, s: int, spec: BackboneSpec, od=None, taps=None):
    pre = f'stages.{s}.'
    factor = spec.patch_size if s == 0 else 2
    x = downsample_cf2cl(x_nchw, p, pre + 'downsample_cf2cl.', factor, spec.overlap, 1e-5, od)
    if token_mask is not None:                          # maxvit_rnn.py:174-176
        assert pre + 'mask_token' in p, 'No mask token present in this stage'
        x = x.clone()
        x[token_mask] = p[pre + 'mask_token'].reshape(-1)
    if taps is not None:
        taps[pre + 'downsample'] = x                    # after the mask token, as the CUDA op emits it
    for b in range(spec.num_blocks[s]):
        bp = f'{pre}att_blocks.{b}.'
        x = partition_attention_cl(x, p, bp + 'att_window.', spec.partition_size, True,
                                   spec.dim_head, spec.norm_eps, od, taps)
        x = partition_attention_cl(x, p, bp + 'att_grid.', spec.partition_size, False,
                                   spec.dim_head, spec.norm_eps, od, taps)
    if taps is not None:
        taps[pre + 'pre_lstm'] = x
    h1, c1 = dws_conv_lstm(x.permute(0, 3, 1, 2), state, p, pre + 'lstm.',
                           spec.dws_conv, spec.dws_conv_only_hidden, od)
    return h1, (h1, c1)


def backbone_forward(x: Tensor, prev_states: Optional[List[State]], p: Dict[str, Tensor],
                     spec: BackboneSpec, token_mask: Optional[Tensor] = None,
                     operand_dtype=None, taps: Optional[dict] = None):
    """RNNDetector.forward (maxvit_rnn.py:93-105): -> ({1..4: feat NCHW}, [(h,c)]*4)."""
    n = len(spec.num_blocks)
    if prev_states is None:
        prev_states = [None] * n
    assert len(prev_states) == n
    out, states = {}, []
    x = x.to(torch.float32)
    for s in range(n):
        x, st = stage_forward(x, prev_states[s], token_mask if s == 0 else None, p, s, spec,
                              operand_dtype, taps)
        states.append(st)
        out[s + 1] = x
    return out, states


# ---------------------------------------------------------------------------
# deterministic synthetic parameters / inputs (numpy legacy RandomState: stable
# across numpy versions and platforms, so fixtures only need to store the seed)
# ---------------------------------------------------------------------------
def param_shapes(spec: BackboneSpec) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys and shapes (verified against the reference's
    ``named_parameters()`` by oracle/make_golden.py)."""
    shp: Dict[str, Tuple[int, ...]] = {}
    cin = spec.input_channels
    for s, c in enumerate(spec.stage_dims):
        pre = f'stages.{s}.'
        f = spec.patch_size if s == 0 else 2
        k = (f - 1) * 2 + 1 if spec.overlap else f
        shp[pre + 'downsample_cf2cl.conv.weight'] = (c, cin, k, k)
        shp[pre + 'downsample_cf2cl.norm.weight'] = (c,)
        shp[pre + 'downsample_cf2cl.norm.bias'] = (c,)
        for b in range(spec.num_blocks[s]):
            for kind in ('att_window', 'att_grid'):
                bp = f'{pre}att_blocks.{b}.{kind}.'
                if not (kind == 'att_window' and b == 0):
                    shp[bp + 'norm1.weight'] = (c,)
                    shp[bp + 'norm1.bias'] = (c,)
                shp[bp + 'self_attn.qkv.weight'] = (3 * c, c)
                shp[bp + 'self_attn.qkv.bias'] = (3 * c,)
                shp[bp + 'self_attn.proj.weight'] = (c, c)
                shp[bp + 'self_attn.proj.bias'] = (c,)
                if spec.ls_init_value > 0:
                    shp[bp + 'ls1.gamma'] = (c,)
                shp[bp + 'norm2.weight'] = (c,)
                shp[bp + 'norm2.bias'] = (c,)
                shp[bp + 'mlp.net.0.0.weight'] = (4 * c, c)
                shp[bp + 'mlp.net.0.0.bias'] = (4 * c,)
                shp[bp + 'mlp.net.2.weight'] = (c, 4 * c)
                shp[bp + 'mlp.net.2.bias'] = (c,)
                if spec.ls_init_value > 0:
                    shp[bp + 'ls2.gamma'] = (c,)
        if spec.dws_conv:
            d = c if spec.dws_conv_only_hidden else 2 * c
            kk = spec.dws_conv_kernel_size
            shp[pre + 'lstm.conv3x3_dws.weight'] = (d, 1, kk, kk)
            shp[pre + 'lstm.conv3x3_dws.bias'] = (d,)
        shp[pre + 'lstm.conv1x1.weight'] = (4 * c, 2 * c, 1, 1)
        shp[pre + 'lstm.conv1x1.bias'] = (4 * c,)
        if spec.enable_masking and s == 0:
            shp[pre + 'mask_token'] = (1, 1, 1, c)
        cin = c
    return shp


def synth_params(spec: BackboneSpec, seed: int, gamma_mode: str = 'uniform') -> Dict[str, Tensor]:
    """Random parameters with every branch 'visible' (SURVEY.md D10): LayerScale gamma
    ~U(0.5,1.5) unless gamma_mode == 'init' (then ls_init_value, the reference default)."""
    import numpy as np
    rs = np.random.RandomState(seed)
    out = {}
    for k, shape in param_shapes(spec).items():
        n = int(np.prod(shape))
        if k.endswith('gamma'):
            v = rs.uniform(0.5, 1.5, n) if gamma_mode == 'uniform' else np.full(n, spec.ls_init_value)
        elif 'norm' in k and k.endswith('weight'):
            v = rs.uniform(0.5, 1.5, n)
        elif k.endswith('bias'):
            v = rs.normal(0.0, 0.1, n)
        elif k.endswith('mask_token'):
            v = rs.normal(0.0, 0.5, n)
        else:
            fan_in = int(np.prod(shape[1:]))
            v = rs.normal(0.0, 1.0 / math.sqrt(fan_in), n)
        out[k] = torch.from_numpy(v.astype(np.float32).reshape(shape))
    return out


def synth_events_tensor(seed: int, b: int, c: int, h: int, w: int, density: float = 0.1,
                        cutoff: int = 10):
    """uint8 event-histogram-like input: ~(1-density) zeros, counts in [1, cutoff]
    (SURVEY.md §8d 'Synthetic inputs')."""
    import numpy as np
    rs = np.random.RandomState(seed)
    vals = rs.randint(1, cutoff + 1, size=(b, c, h, w)).astype(np.uint8)
    keep = rs.uniform(size=(b, c, h, w)) < density
    return torch.from_numpy(vals * keep)

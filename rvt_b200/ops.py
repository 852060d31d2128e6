"""Thin tensor-level wrappers over the C-ABI operators (include/rvt_b200.h).  One function per
reference function on the hot path; used by rvt_b200.backbone and by the parity
tests.  Channels-last fp32 tensors [B, H, W, C]; CUDA only."""
import functools
from typing import Optional, Tuple

import torch

from . import _lib


def device_guarded(fn):
    """The C-ABI launches on the CURRENT device (kernel attributes, SM counts and tensor maps are per device): when the
    first tensor argument lives on another GPU, run the call under that device."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        t = next((a for a in args if isinstance(a, torch.Tensor)), None)
        if t is not None and t.is_cuda and t.device.index != torch.cuda.current_device():
            with torch.cuda.device(t.device):
                return fn(*args, **kw)
        return fn(*args, **kw)
    return wrapper

_IN_DTYPES = {torch.float32: 0, torch.uint8: 1, torch.float16: 2}   # channels-last inputs: f32 or f16


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


@device_guarded
def downsample_cf2cl(x: torch.Tensor, x_is_nchw: bool, conv_w_packed: torch.Tensor, cout: int, ksize: int,
                     stride: int, pad: int, ln_w: Optional[torch.Tensor], ln_b: Optional[torch.Tensor],
                     virtual_hw: Optional[Tuple[int, int]] = None, token_mask: Optional[torch.Tensor] = None,
                     mask_token: Optional[torch.Tensor] = None, eps: float = 1e-5,
                     s2d_scratch: Optional[torch.Tensor] = None, stem_mode: int = 0,
                     split_ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ConvDownsampling_Cf2Cl.forward (maxvit.py:174-178) [+ mask token, maxvit_rnn.py:174-176].
    x: [B,Cin,H,W] (f32/u8/f16) if x_is_nchw else [B,H,W,Cin] f32.  -> f32 [B,Hout,Wout,cout]."""
    assert x.is_cuda and x.is_contiguous() and x.dtype in _IN_DTYPES
    if s2d_scratch is not None:
        assert x_is_nchw and s2d_scratch.dtype == torch.float16
    if x_is_nchw:
        b, cin, hin, win = x.shape
    else:
        b, hin, win, cin = x.shape
    vh, vw = virtual_hw if virtual_hw is not None else (hin, win)
    assert vh >= hin and vw >= win, 'input larger than the model resolution'
    hout = (vh + 2 * pad - ksize) // stride + 1
    wout = (vw + 2 * pad - ksize) // stride + 1
    out = torch.empty((b, hout, wout, cout), dtype=torch.float32, device=x.device)
    if token_mask is not None:
        token_mask = token_mask.to(device=x.device, dtype=torch.uint8).contiguous()
        assert tuple(token_mask.shape) == (b, hout, wout) and mask_token is not None
    L = _lib.lib()
    if s2d_scratch is None and not x_is_nchw and cout >= 256:
        # wide stages: split-K workspace (one fp32 partial tile per K slice, summed by the LayerNorm pass)
        splits = L.rvt_conv_split_k(b * hout * wout, cout, cin * ksize * ksize)
        if splits > 1:
            s2d_scratch = split_ws if split_ws is not None else torch.empty(splits * b * hout * wout * cout, dtype=torch.float32,
                                                                           device=x.device)
            assert s2d_scratch.dtype == torch.float32 and s2d_scratch.numel() >= splits * b * hout * wout * cout
    _lib.check(L.rvt_downsample_cf2cl(
        _lib.ptr(x), _IN_DTYPES[x.dtype], int(x_is_nchw), b, cin, hin, win, ksize, stride, pad, hout, wout, cout,
        _lib.ptr(conv_w_packed), _lib.ptr(ln_w), _lib.ptr(ln_b), eps, _lib.ptr(token_mask), _lib.ptr(mask_token),
        _lib.ptr(out), _lib.ptr(s2d_scratch), stem_mode, _stream(x)), 'downsample_cf2cl')
    return out


@device_guarded
def stem_u8_ok(x: torch.Tensor, cin: int, ksize: int, stride: int, pad: int, virtual_hw, cout: int) -> bool:
    """uint8 NCHW input + the default stem geometry -> smem-patch loader (no scratch tensor)."""
    if x.dtype != torch.uint8 or x.data_ptr() % 16:
        return False
    vh, vw = virtual_hw
    hout, wout = (vh + 2 * pad - ksize) // stride + 1, (vw + 2 * pad - ksize) // stride + 1
    return bool(_lib.lib().rvt_stem_u8_ok(cin, ksize, stride, pad, x.shape[3], hout, wout, cout))


def stem_uses_s2d(cin: int, factor: int, ksize: int, pad: int, virtual_w: int) -> bool:
    """The space-to-depth stem path applies to the reference's two stem geometries when the
    grouped channel count is 16-byte aligned."""
    overlap = ksize == 2 * factor - 1 and pad == factor - 1
    patch = ksize == factor and pad == 0
    return (overlap or patch) and 64 % factor == 0 and (factor * cin) % 8 == 0 and virtual_w % factor == 0


def attention_scratch_rows(b, h, w, part) -> int:
    rows = _lib.lib().rvt_attention_scratch_rows(b, h, w, part[0], part[1])
    if rows < 0:
        raise RuntimeError(f'rvt_b200: partition {part} does not tile {h}x{w} or exceeds 128 tokens')
    return rows


@device_guarded
def partition_attention_(x: torch.Tensor, blk: dict, scratch_qkv: torch.Tensor, scratch_o: torch.Tensor,
                         scratch_xn: Optional[torch.Tensor] = None) -> None:
    """In place: x += ls1(proj(attn(partition(norm1(x)))))  (maxvit.py:252-268)."""
    assert x.is_cuda and x.is_contiguous() and x.dtype == torch.float32
    b, h, w, c = x.shape
    ph, pw = blk['part']
    rows = attention_scratch_rows(b, h, w, (ph, pw))
    if not _lib.lib().rvt_attention_is_fused(c, blk['dh']):
        assert scratch_qkv.numel() >= rows * 3 * c and scratch_o.numel() >= rows * c
    L = _lib.lib()
    _lib.check(L.rvt_partition_attention(
        _lib.ptr(x), b, h, w, c, ph, pw, blk['grid'], blk['dh'], _lib.ptr(blk['n1_w']), _lib.ptr(blk['n1_b']),
        blk['eps'], _lib.ptr(blk['wqkv']), _lib.ptr(blk['bqkv']), _lib.ptr(blk['wproj']), _lib.ptr(blk['bproj']),
        _lib.ptr(blk['g1']), _lib.ptr(scratch_qkv), _lib.ptr(scratch_o), _lib.ptr(scratch_xn), _stream(x)),
        'partition_attention')


@device_guarded
def mlp_block_(x: torch.Tensor, blk: dict, scratch_hidden: torch.Tensor,
               scratch_xn: Optional[torch.Tensor] = None) -> None:
    """In place: x += ls2(mlp(norm2(x)))  (maxvit.py:269)."""
    assert x.is_cuda and x.is_contiguous() and x.dtype == torch.float32
    c = x.shape[-1]
    n_tok = x.numel() // c
    hid = blk['hidden']
    assert scratch_hidden.numel() >= ((n_tok + 127) // 128) * 128 * hid
    L = _lib.lib()
    _lib.check(L.rvt_mlp_block(
        _lib.ptr(x), n_tok, c, hid, _lib.ptr(blk['n2_w']), _lib.ptr(blk['n2_b']), blk['eps'], _lib.ptr(blk['w1']),
        _lib.ptr(blk['b1']), _lib.ptr(blk['w2']), _lib.ptr(blk['b2']), _lib.ptr(blk['g2']),
        _lib.ptr(scratch_hidden), _lib.ptr(scratch_xn), _stream(x)), 'mlp_block')


@device_guarded
def dws_conv_lstm(x: torch.Tensor, h_prev: Optional[torch.Tensor], c_prev: Optional[torch.Tensor], pk: dict,
                  dws_ks: int, scratch_xh: Optional[torch.Tensor] = None,
                  h16_out: Optional[torch.Tensor] = None, h_out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """DWSConvLSTM2d.forward (rnn.py:36-69) on channels-last tensors -> (h_t, c_t).  h_out: optional destination of h_t
    (a slice of a whole-sequence feature buffer, see RNNDetector.forward_sequence(select=...))."""
    assert x.is_cuda and x.is_contiguous() and x.dtype == torch.float32
    b, h, w, c = x.shape
    for t in (h_prev, c_prev):
        assert t is None or (t.shape == x.shape and t.is_contiguous() and t.dtype == torch.float32)
    if h_out is not None:
        assert h_out.shape == x.shape and h_out.is_contiguous() and h_out.dtype == torch.float32 and h_out.device == x.device
    h_new, c_new = (h_out if h_out is not None else torch.empty_like(x)), torch.empty_like(x)
    L = _lib.lib()
    _lib.check(L.rvt_dws_conv_lstm(
        _lib.ptr(x), _lib.ptr(h_prev), _lib.ptr(c_prev), b, h, w, c, _lib.ptr(pk['lstm_w']), _lib.ptr(pk['lstm_b']),
        _lib.ptr(pk['dw_w']), _lib.ptr(pk['dw_b']), pk['dws_mode'], dws_ks, _lib.ptr(h_new), _lib.ptr(c_new),
        _lib.ptr(scratch_xh), _lib.ptr(h16_out), _stream(x)), 'dws_conv_lstm')
    return h_new, c_new


# =============================================================================================
# harness glue of modules/utils/detection.py as device-side ops (graph-capturable, no host sync)
# =============================================================================================
@device_guarded
def state_reset_(h: torch.Tensor, c: Optional[torch.Tensor], mask: torch.Tensor) -> None:
    """RNNStates.reset -> recursive_reset (modules/utils/detection.py:96-113): ``state[mask] = 0`` in place for one stage's
    (h, c); h / c contiguous f32 with the batch as the leading dimension, mask bool / uint8 [B] on the same device."""
    assert h.is_cuda and h.is_contiguous() and h.dtype == torch.float32
    assert c is None or (c.shape == h.shape and c.is_contiguous() and c.dtype == torch.float32)
    assert mask.device == h.device and mask.numel() == h.shape[0] and mask.dtype in (torch.bool, torch.uint8) and mask.is_contiguous()
    b = h.shape[0]
    _lib.check(_lib.lib().rvt_state_reset(_lib.ptr(h), _lib.ptr(c), _lib.ptr(mask), b, h.numel() // max(b, 1), _stream(h)),
               'state_reset')


@device_guarded
def gather_rows(src: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """BackboneFeatureSelector (modules/utils/detection.py:24-46) over a whole sequence: out[j] = src[idx[j]] for the leading
    dimension (idx int32 [S] on the device, negative = empty slot -> zeros)."""
    assert src.is_cuda and src.is_contiguous() and src.dtype == torch.float32
    assert idx.device == src.device and idx.dtype == torch.int32 and idx.is_contiguous()
    rows = src.shape[0]
    row_elems = src.numel() // max(rows, 1)
    if out is None:
        out = torch.empty((idx.numel(),) + tuple(src.shape[1:]), dtype=torch.float32, device=src.device)
    assert out.is_contiguous() and out.numel() == idx.numel() * row_elems
    _lib.check(_lib.lib().rvt_gather_rows(_lib.ptr(src), _lib.ptr(idx), idx.numel(), rows, row_elems, _lib.ptr(out), _stream(src)),
               'gather_rows')
    return out


# =============================================================================================
# Training building blocks (include/rvt_b200.h "Training step"); composed in rvt_b200/train.py.
# All tensors CUDA + contiguous; f16 matrices are row-major with ld == number of columns.
# =============================================================================================
import os as _os
TN_MODE = int(_os.environ.get('RVT_TN_MODE', '0'))   # rvt_gemm_tn operand form: 0 = MN-major tiles in place, 1 = transposed copies (K-major)


def round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


@device_guarded
def linear_ex(a: torch.Tensor, m: int, k: int, n: int, w_packed: torch.Tensor, out: torch.Tensor,
              bias: Optional[torch.Tensor] = None, act: int = 0, aux: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = a[:m, :k] @ W^T (+bias, act); a f16 [>=m, k]; out f16 [round_up(m,128), n] or f32 [m, n]."""
    assert a.dtype == torch.float16 and a.numel() >= m * k
    out_f32 = out.dtype == torch.float32
    assert out.numel() >= (m if out_f32 else round_up(m, 128)) * n
    if aux is not None:
        assert aux.dtype == torch.float16 and aux.numel() >= round_up(m, 128) * n
    _lib.check(_lib.lib().rvt_linear_ex(_lib.ptr(a), m, k, n, _lib.ptr(w_packed), _lib.ptr(bias), act, _lib.ptr(aux),
                                        _lib.ptr(out), int(out_f32), _stream(a)), 'linear_ex')
    return out


@device_guarded
def gemm_tn(a1: torch.Tensor, n1: int, a2: torch.Tensor, n2: int, m: int, g: torch.Tensor, transpose_out: bool = False,
            mode: Optional[int] = None, colsum1: Optional[torch.Tensor] = None, colsum2: Optional[torch.Tensor] = None) -> None:
    """g[n1, n2] += a1[:m, :n1]^T @ a2[:m, :n2]  (g f32; transpose_out: g is [n2, n1] and receives the transpose).
    The MMA tile is 128 rows of a1^T x up to 128 of a2^T: pass the wider operand as a1."""
    assert a1.dtype == torch.float16 and a2.dtype == torch.float16 and g.dtype == torch.float32
    assert a1.numel() >= m * n1 and a2.numel() >= m * n2 and g.numel() == n1 * n2 and g.is_contiguous()
    mode = TN_MODE if mode is None else mode
    L = _lib.lib()
    scratch = None
    if mode == 1:
        scratch = torch.empty(L.rvt_gemm_tn_scratch_elems(m, n1, n2), dtype=torch.float16, device=a1.device)
    s_i, s_j = (1, n1) if transpose_out else (n2, 1)
    for cs, n in ((colsum1, n1), (colsum2, n2)):
        assert cs is None or (cs.dtype == torch.float32 and cs.numel() >= n)
    _lib.check(L.rvt_gemm_tn(_lib.ptr(a1), n1, n1, _lib.ptr(a2), n2, n2, m, _lib.ptr(g), s_i, s_j, mode,
                             _lib.ptr(scratch), _lib.ptr(colsum1), _lib.ptr(colsum2), _stream(a1)), 'gemm_tn')


@device_guarded
def ln_rows_f16(x: torch.Tensor, map_mode: int, part, ln_w, ln_b, do_ln: bool, eps: float, out16: torch.Tensor) -> None:
    b, h, w, c = x.shape
    ph, pw = part if part is not None else (1, 1)
    _lib.check(_lib.lib().rvt_ln_rows_f16(_lib.ptr(x), map_mode, b, h, w, c, ph, pw, _lib.ptr(ln_w), _lib.ptr(ln_b),
                                          int(do_ln), eps, _lib.ptr(out16), _stream(x)), 'ln_rows_f16')


@device_guarded
def ln_bwd(x: Optional[torch.Tensor], dy: torch.Tensor, shape, map_mode: int, part, ln_w, do_ln: bool, eps: float,
           dres: Optional[torch.Tensor], dx16: Optional[torch.Tensor], dw_acc, db_acc) -> None:
    b, h, w, c = shape
    ph, pw = part if part is not None else (1, 1)
    _lib.check(_lib.lib().rvt_ln_bwd(_lib.ptr(x), _lib.ptr(dy), int(dy.dtype == torch.float16), map_mode, b, h, w, c, ph, pw,
                                     _lib.ptr(ln_w), int(do_ln), eps, _lib.ptr(dres), _lib.ptr(dx16), _lib.ptr(dw_acc),
                                     _lib.ptr(db_acc), _stream(dy)), 'ln_bwd')


@device_guarded
def gather_cast(dres: torch.Tensor, map_mode: int, part, gamma, d0, d1) -> None:
    b, h, w, c = dres.shape
    ph, pw = part if part is not None else (1, 1)
    _lib.check(_lib.lib().rvt_gather_cast(_lib.ptr(dres), map_mode, b, h, w, c, ph, pw, _lib.ptr(gamma), _lib.ptr(d0),
                                          _lib.ptr(d1), _stream(dres)), 'gather_cast')


@device_guarded
def attn_core_bwd(qkv, o, dout, dqkv, shape, part, dim_head: int) -> None:
    b, h, w, c = shape
    _lib.check(_lib.lib().rvt_attn_core_bwd(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(dout), _lib.ptr(dqkv), b, h, w, c, part[0],
                                            part[1], dim_head, _stream(qkv)), 'attn_core_bwd')


@device_guarded
def lstm_gates_bwd(gates, c_prev, c_new, dh, dc, n_tokens: int, dim: int, dpre, dc_prev) -> None:
    _lib.check(_lib.lib().rvt_lstm_gates_bwd(_lib.ptr(gates), _lib.ptr(c_prev), _lib.ptr(c_new), _lib.ptr(dh), _lib.ptr(dc),
                                             n_tokens, dim, _lib.ptr(dpre), _lib.ptr(dc_prev), _stream(gates)),
               'lstm_gates_bwd')


@device_guarded
def im2col(x: torch.Tensor, x_is_nchw: bool, ksize: int, stride: int, pad: int, hout: int, wout: int, col: torch.Tensor) -> None:
    if x_is_nchw:
        b, cin, hin, win = x.shape
    else:
        b, hin, win, cin = x.shape
    _lib.check(_lib.lib().rvt_im2col(_lib.ptr(x), _IN_DTYPES[x.dtype], int(x_is_nchw), b, cin, hin, win, ksize, stride, pad,
                                     hout, wout, _lib.ptr(col), _stream(x)), 'im2col')


@device_guarded
def col2im(dcol: torch.Tensor, b, cin, hin, win, ksize, stride, pad, hout, wout, d_in: torch.Tensor) -> None:
    _lib.check(_lib.lib().rvt_col2im(_lib.ptr(dcol), b, cin, hin, win, ksize, stride, pad, hout, wout, _lib.ptr(d_in),
                                     _stream(dcol)), 'col2im')


@device_guarded
def colsum(a: torch.Tensor, m: int, n: int, acc: torch.Tensor) -> None:
    assert a.dtype == torch.float16 and acc.dtype == torch.float32 and acc.numel() >= n
    _lib.check(_lib.lib().rvt_colsum(_lib.ptr(a), m, n, n, _lib.ptr(acc), _stream(a)), 'colsum')


@device_guarded
def nchw_to_nhwc_f16(x: torch.Tensor, channels_padded: int, out: torch.Tensor) -> None:
    b, c, h, w = x.shape
    assert out.dtype == torch.float16 and out.numel() >= b * h * w * channels_padded
    _lib.check(_lib.lib().rvt_nchw_to_nhwc_f16(_lib.ptr(x), _IN_DTYPES[x.dtype], b, c, h, w, channels_padded, _lib.ptr(out),
                                               _stream(x)), 'nchw_to_nhwc_f16')

"""ctypes binding of include/rvt_b200.h.  There is NO fallback: if the CUDA library is missing
or a call fails, this raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'librvt_b200.so')

_c = ctypes
_vp, _i, _i64, _f = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float

# name -> (restype, argtypes); mirrors include/rvt_b200.h exactly (tests/test_capi_symbols.py)
SIGNATURES = {
    'rvt_abi_version': (_i, []),
    'rvt_error_string': (_c.c_char_p, [_i]),
    'rvt_tile_n': (_i, [_i, _i]),
    'rvt_attention_is_fused': (_i, [_i, _i]),
    'rvt_mlp_tiles': (_i, [_i, _i, _vp, _vp]),
    'rvt_conv_tile_n': (_i, [_i]),
    'rvt_conv_split_k': (_i, [_i64, _i, _i]),
    'rvt_lstm_cw': (_i, [_i]),
    'rvt_rows_per_group': (_i, [_i]),
    'rvt_attention_scratch_rows': (_i64, [_i, _i, _i, _i, _i]),
    'rvt_stacked_histogram': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'rvt_downsample_cf2cl': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f,
                                   _vp, _vp, _vp, _vp, _i, _vp]),
    'rvt_stem_u8_ok': (_i, [_i, _i, _i, _i, _i, _i, _i, _i]),
    'rvt_partition_attention': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _vp, _vp, _vp]),
    'rvt_mlp_block': (_i, [_vp, _i64, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'rvt_dws_conv_lstm': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'rvt_linear_f16': (_i, [_vp, _i64, _i, _i, _vp, _vp, _i, _vp, _vp]),
    # ---- training step ----
    'rvt_downsample_cf2cl_train': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f,
                                         _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    'rvt_partition_attention_train': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp,
                                            _vp, _vp, _vp, _vp, _vp, _vp]),
    'rvt_mlp_block_train': (_i, [_vp, _vp, _i64, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'rvt_dws_conv_lstm_train': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'rvt_linear_ex': (_i, [_vp, _i64, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    'rvt_gemm_tn': (_i, [_vp, _i, _i, _vp, _i, _i, _i64, _vp, _i64, _i64, _i, _vp, _vp, _vp, _vp]),
    'rvt_nchw_to_nhwc_f16': (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    'rvt_gemm_tn_scratch_elems': (_i64, [_i64, _i, _i]),
    'rvt_ln_rows_f16': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _f, _vp, _vp]),
    'rvt_ln_bwd': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    'rvt_gather_cast': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'rvt_attn_core_bwd': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'rvt_lstm_gates_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp]),
    'rvt_im2col': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    'rvt_col2im': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    'rvt_colsum': (_i, [_vp, _i64, _i, _i, _vp, _vp]),
    # ---- harness glue + preprocessing neighbours ----
    'rvt_state_reset': (_i, [_vp, _vp, _vp, _i, _i64, _vp]),
    'rvt_gather_rows': (_i, [_vp, _vp, _i, _i64, _i64, _vp, _vp]),
    'rvt_downsample2_nearest': (_i, [_vp, _i, _i, _i, _vp, _vp]),
    'rvt_cummax_scratch_elems': (_i64, [_i64]),
    'rvt_cummax_i64': (_i, [_vp, _i64, _i64, _vp, _vp]),
    'rvt_searchsorted_i64': (_i, [_vp, _i64, _vp, _i64, _i, _vp, _vp]),
    'rvt_conv2d_nhwc_f16': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp]),
    'rvt_cast_slice_f16': (_i, [_vp, _i64, _i64, _i64, _i64, _i, _i, _i, _i, _vp, _i, _vp]),
    'rvt_upsample2_slice_f16': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    'rvt_yolox_decode': (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _vp]),
    'rvt_yolox_postprocess': (_i, [_vp, _i, _i, _i, _f, _f, _vp, _vp, _vp]),
    'rvt_debug_set_trace': (_i, [_vp]),
    'rvt_mixed_density_stack': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'rvt_b200: CUDA library not built ({LIB_PATH}). Run `python -m rvt_b200.build` '
                f'(or __graft_entry__.build()). There is no CPU / PyTorch fallback.')
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.rvt_abi_version() != 1:
            raise RuntimeError('rvt_b200: ABI version mismatch, rebuild the library')
        _lib = l
    return _lib


def check(code: int, what: str):
    if code != 0:
        msg = lib().rvt_error_string(code).decode()
        raise RuntimeError(f'rvt_b200: {what} failed: [{code}] {msg}')


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def mlp_tiles(dim: int, hidden: int):
    """(bn_fc1, bn_fc2, fused) — the N-tiles rvt_mlp_block expects its weights packed with."""
    b1, b2 = ctypes.c_int(0), ctypes.c_int(0)
    fused = lib().rvt_mlp_tiles(dim, hidden, ctypes.addressof(b1), ctypes.addressof(b2))
    return b1.value, b2.value, bool(fused)

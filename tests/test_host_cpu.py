"""CPU: host-side logic — C-ABI library exports, weight packing layout, module API parity
(state_dict names/shapes vs the reference's, verified at golden-mint time)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import backbone_oracle as bo
from tests.golden_configs import BACKBONE_CASES, spec_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_cfg(spec: bo.BackboneSpec):
    """Plain-dict mdl_config with the fields the reference reads."""
    return dict(
        name='MaxViTRNN', compile=dict(enable=False, args=dict(mode='reduce-overhead')),
        input_channels=spec.input_channels, enable_masking=spec.enable_masking, partition_split_32=1,
        embed_dim=spec.embed_dim, dim_multiplier=list(spec.dim_multiplier), num_blocks=list(spec.num_blocks),
        T_max_chrono_init=[4, 8, 16, 32], stem=dict(patch_size=spec.patch_size),
        stage=dict(downsample=dict(type='patch', overlap=spec.overlap, norm_affine=True),
                   attention=dict(use_torch_mha=False, partition_size=tuple(spec.partition_size),
                                  dim_head=spec.dim_head, attention_bias=True, mlp_activation='gelu',
                                  mlp_gated=False, mlp_bias=True, mlp_ratio=4, drop_mlp=0, drop_path=0,
                                  ls_init_value=spec.ls_init_value, norm_eps=spec.norm_eps),
                   lstm=dict(dws_conv=spec.dws_conv, dws_conv_only_hidden=spec.dws_conv_only_hidden,
                             dws_conv_kernel_size=spec.dws_conv_kernel_size, drop_cell_update=0)))


@pytest.fixture(scope='module')
def built_lib():
    from rvt_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    """Every function declared in include/rvt_b200.h is exported by the .so and bound in _lib.py."""
    hdr = open(os.path.join(ROOT, 'include', 'rvt_b200.h')).read()
    declared = set(re.findall(r'\b(rvt_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 10
    so = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(so, name), f'{name} declared in the header but not exported'
    from rvt_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_ctypes_signatures_match_header(built_lib):
    """Arity and scalar/pointer kind of every ctypes binding equal the C prototype in include/rvt_b200.h."""
    from rvt_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'rvt_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    protos = re.findall(r'\b([a-z_0-9]+\s*\**)\s*\b(rvt_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', hdr)
    assert len(protos) == len(_lib.SIGNATURES)
    for _ret, name, args in protos:
        params = [a.strip() for a in args.split(',')] if args.strip() != 'void' else []
        _res, argtypes = _lib.SIGNATURES[name]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for prm, at in zip(params, argtypes):
            is_ptr = '*' in prm
            assert is_ptr == (at is ctypes.c_void_p), (name, prm, at)
            if not is_ptr:
                want = ctypes.c_float if prm.startswith('float') else ctypes.c_int64 if prm.startswith('int64_t') else ctypes.c_int
                assert at is want, (name, prm, at)


def test_tiling_helpers(built_lib):
    from rvt_b200 import _lib
    L = _lib.lib()
    assert L.rvt_abi_version() == 1
    assert [L.rvt_tile_n(n, k) for n, k in ((64, 64), (192, 64), (256, 64), (1536, 512), (144, 48), (512, 2048), (768, 256))] == \
        [64, 96, 128, 256, 48, 256, 256]           # narrow stages: one N tile up to 128; wide stages (min(n, k) >= 256): tiles of 256
    assert [L.rvt_lstm_cw(c) for c in (32, 48, 64, 96, 192, 512)] == [32, 48, 64, 48, 64, 64]
    assert L.rvt_rows_per_group(60) == 64 and L.rvt_rows_per_group(80) == 128 and L.rvt_rows_per_group(129) < 0
    assert L.rvt_attention_scratch_rows(8, 96, 160, 6, 10) == 2048 * 64
    assert L.rvt_attention_scratch_rows(1, 8, 10, 8, 10) == 128
    assert L.rvt_error_string(0) == b'ok'


def test_pack_linear_weight_layout():
    """Vectorised packer == scalar statement of the SW128 K-major tile image (csrc/umma.cuh)."""
    from rvt_b200 import packing
    torch.manual_seed(0)
    n, k, bn = 96, 72, 48
    w = torch.randn(n, k)
    got = packing.pack_linear_weight(w, bn).reshape(-1)
    kc = (k + 63) // 64
    ref = torch.zeros(n // bn, kc, bn * 64, dtype=torch.float16)
    w16 = w.to(torch.float16)
    for r_glob in range(n):
        nt, r = divmod(r_glob, bn)
        for kk in range(k):
            c, within = divmod(kk % 64, 8)
            byte = r * 128 + ((c ^ (r & 7)) << 4) + within * 2
            ref[nt, kk // 64, byte // 2] = w16[r_glob, kk]
    assert torch.equal(got, ref.reshape(-1))


def test_stem_weight_u8_k_order():
    """pack_stem_weight_u8: K = (kyi, ci, kx8) with kernel rows in the row-phase order the stem kernels assume
    (csrc/gemm_fused.cuh stem_ky: 0, 4, 1, 5, 2, 6, 3), kx8 = kx + 1 and a zero weight in slot kx8 = 0."""
    from rvt_b200 import packing
    assert packing.STEM_KY_ORDER == tuple((0x3625140 >> (4 * i)) & 7 for i in range(7))      # == stem_ky(i) in the CUDA header
    torch.manual_seed(0)
    co, cin = 32, 3
    w = torch.randn(co, cin, 7, 7)
    got = packing.pack_stem_weight_u8(w).reshape(-1)
    k = 7 * cin * 8
    kc = (k + 63) // 64
    ref = torch.zeros(1, kc, co * 64, dtype=torch.float16)
    w16 = w.to(torch.float16)
    for r in range(co):
        for kyi, ky in enumerate(packing.STEM_KY_ORDER):
            for ci in range(cin):
                for kx in range(7):
                    kk = (kyi * cin + ci) * 8 + kx + 1
                    c, within = divmod(kk % 64, 8)
                    ref[0, kk // 64, (r * 128 + ((c ^ (r & 7)) << 4) + within * 2) // 2] = w16[r, ci, ky, kx]
    assert torch.equal(got, ref.reshape(-1))


def test_lstm_row_order():
    from rvt_b200 import packing
    order = packing.lstm_row_order(8, 4).tolist()
    # tile 0: f[0:4] i[0:4] o[0:4] g[0:4]; tile 1: f[4:8] ...
    assert order[:16] == [0, 1, 2, 3, 8, 9, 10, 11, 16, 17, 18, 19, 24, 25, 26, 27]
    assert sorted(order) == list(range(32))


@pytest.mark.parametrize('name', ['tiny_p6', 'small_dh24', 'dws_hidden', 'dws_xh', 'ls_init_mask'])
def test_state_dict_matches_reference_keys(name):
    import rvt_b200
    spec = spec_of(BACKBONE_CASES[name])
    m = rvt_b200.build_recurrent_backbone(make_cfg(spec))
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert mine == bo.param_shapes(spec)      # param_shapes is asserted == reference at golden-mint time
    m.load_state_dict(bo.synth_params(spec, 1), strict=True)
    assert m.get_stage_dims((2, 3, 4)) == tuple(spec.stage_dims[1:])
    assert m.get_strides((1, 2, 3, 4)) == (4, 8, 16, 32)


def test_no_cpu_fallback():
    import rvt_b200
    spec = spec_of(BACKBONE_CASES['tiny_p6'])
    m = rvt_b200.RNNDetector(make_cfg(spec)).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match='no CPU fallback'):
        m(torch.zeros(1, 20, 64, 96))
    m.train()
    with pytest.raises(RuntimeError, match='no CPU fallback'):       # grad mode = the training path: same rule
        m(torch.zeros(1, 20, 64, 96))
    sh = rvt_b200.StackedHistogram(10, 8, 8, 10)
    z = torch.zeros(0, dtype=torch.int64)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        sh.construct(z, z, z, z)
    assert sh.get_shape() == (20, 8, 8) and sh.dtype == torch.uint8


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under rvt_b200/ may import or execute it (task rule 3)."""
    pkg = os.path.join(ROOT, 'rvt_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn

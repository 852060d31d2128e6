"""GPU: the tcgen05 GEMM building block (rvt_linear_f16) against an fp32 CPU matmul of the same
fp16-rounded operands.  First line of defence for the UMMA descriptor / swizzle / TMEM plumbing:
on failure an identity-weight probe is dumped to gpurun_out/ for offline layout forensics."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_linear(a16, w, bias, act=0):
    from rvt_b200 import _lib, packing
    L = _lib.lib()
    m, k = a16.shape
    n = w.shape[0]
    dev = a16.device
    wp = packing.pack_linear_weight(w.to(dev), L.rvt_tile_n(n, k))
    out = torch.zeros(((m + 127) // 128) * 128, n, dtype=torch.float16, device=dev)
    _lib.check(L.rvt_linear_f16(_lib.ptr(a16), m, k, n, _lib.ptr(wp), _lib.ptr(bias), act, _lib.ptr(out),
                                torch.cuda.current_stream().cuda_stream), 'linear_f16')
    torch.cuda.synchronize()
    return out[:m]


def test_identity_probe():
    """W = I: the output must reproduce A exactly (fp16 -> fp32 acc -> fp16 is lossless)."""
    dev = torch.device('cuda:0')
    m = k = n = 64
    a = (torch.arange(m).view(m, 1) * 1.0 + torch.arange(k).view(1, k) / 64.0).to(torch.float16).to(dev)
    w = torch.eye(n, k)
    out = run_linear(a, w, None)
    if not torch.equal(out, a):
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        np.savez(os.path.join(ROOT, 'gpurun_out', 'diag_identity.npz'), a=a.cpu().numpy(), out=out.cpu().numpy())
    assert torch.equal(out, a)


@pytest.mark.parametrize('m,k,n,act', [(128, 64, 64, 0), (300, 64, 192, 0), (128, 256, 256, 0), (1000, 512, 1536, 0),
                                       (257, 72, 144, 0), (128, 2048, 512, 0), (640, 128, 512, 1), (5, 8, 16, 0)])
def test_linear_matches_fp32_matmul(m, k, n, act):
    torch.manual_seed(m * 7 + k)
    dev = torch.device('cuda:0')
    a = torch.randn(m, k).to(torch.float16)
    w = (torch.randn(n, k) / k ** 0.5)
    b = torch.randn(n) * 0.1
    ref = a.float() @ w.to(torch.float16).float().t() + b
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    out = run_linear(a.to(dev), w, b.to(dev), act).float().cpu()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-3, err      # fp16 output rounding (2^-11 relative) dominates

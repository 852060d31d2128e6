"""CPU numeric study behind the opt-in packed-half GELU (csrc/gemm_fused.cuh gelu_f16x2, RVT_GELU_F16X2=1).
Emulates every fp16 rounding of the device code (HMUL2 / HFMA2 single rounding, tanh.approx.f16x2 modelled as the exact
tanh plus a uniform error of the PTX-documented bound 2^-10.987, result rounded to fp16) and compares with the exact
erf GELU.  The fitted constants are a weighted least-squares fit of atanh(erf(x / sqrt2)) by the odd cubic c1 x + c3 x^3.
usage: python profiles/gelu_f16x2_study.py"""
import math

import numpy as np
import torch


def fit():
    x = np.linspace(-5, 5, 20001)
    t = np.vectorize(math.erf)(x / math.sqrt(2))
    target = np.arctanh(np.clip(t, -1 + 1e-12, 1 - 1e-12))
    w = (1 - t * t) * np.maximum(np.abs(x), 0.2)
    a = np.stack([x, x ** 3], 1)
    return np.linalg.lstsq(a * w[:, None], target * w, rcond=None)[0]


def fma16(a, b, c):
    return (a.float() * b.float() + c.float()).half()


def main():
    c1, c3 = fit()
    print(f'fit: c1 = {c1:.8f}, c3 = {c3:.8f}  (device constants 0.79978222, 0.03487167)')
    torch.manual_seed(0)
    v = torch.cat([torch.randn(2_000_000) * 1.5, torch.randn(1_000_000) * 0.5, torch.linspace(-8, 8, 400001)])
    exact = torch.nn.functional.gelu(v.double()).float()

    def stats(name, out):
        err = out.float() - exact
        print(f'{name:58s} rel-L2 {float(err.norm() / exact.norm()):.3e}  max abs {float(err.abs().max()):.3e}')

    stats('exact erf GELU rounded to fp16 (the default kernel)', exact.half())
    v16 = v.half()
    u = (v16.float() * v16.float()).half()
    q = fma16(torch.tensor(c3).half().expand_as(u), u, torch.tensor(c1).half().expand_as(u))
    p = (q.float() * v16.float()).half()
    hv = (0.5 * v16.float()).half()
    for noise in (0.0, 2 ** -10.987):
        t = torch.tanh(p.float())
        if noise:
            t = t + (torch.rand_like(t) * 2 - 1) * noise
        stats(f'packed-half path, tanh.approx error bound {noise:.2e}', fma16(hv, t.half(), hv))


if __name__ == '__main__':
    main()

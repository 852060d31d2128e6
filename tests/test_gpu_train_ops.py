"""GPU: every building block of the training step (include/rvt_b200.h "Training step") against a plain
PyTorch fp32 reference of the same operator / its autograd gradient.  Tolerances: fp16 operand / output
rounding (2^-11 relative per element) with fp32 accumulation -> rel-L2 <= 2e-3 unless noted."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import backbone_oracle as bo  # noqa: E402


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    return torch.device('cuda:0')


TN_SHAPES = [(1000, 64, 256), (5000, 192, 96), (777, 32, 48), (128, 128, 128), (46080, 256, 64), (640, 512, 1024)]


def _check_gemm_tn(dev, mode, m, n1, n2):
    from rvt_b200 import ops
    g = torch.Generator(device='cpu').manual_seed(m + n1)
    a1 = (torch.randn(m, n1, generator=g) * 0.5).to(dev).half()
    a2 = (torch.randn(m, n2, generator=g) * 0.5).to(dev).half()
    ref = a1.float().t() @ a2.float()
    acc = torch.zeros(n1, n2, device=dev)
    cs1, cs2 = torch.ones(n1, device=dev), torch.ones(n2, device=dev)
    ops.gemm_tn(a1, n1, a2, n2, m, acc, mode=mode, colsum1=cs1, colsum2=cs2)       # column sums ride along
    torch.cuda.synchronize()
    assert rel_l2(acc, ref) < 1e-3, (mode, rel_l2(acc, ref))
    assert rel_l2(cs1 - 1, a1.float().sum(0)) < 1e-4 and rel_l2(cs2 - 1, a2.float().sum(0)) < 1e-4
    ops.gemm_tn(a1, n1, a2, n2, m, acc, mode=mode)          # accumulates
    assert rel_l2(acc, 2 * ref) < 1e-3
    acc_t = torch.zeros(n2, n1, device=dev)
    ops.gemm_tn(a1, n1, a2, n2, m, acc_t, transpose_out=True, mode=mode)
    assert rel_l2(acc_t, ref.t()) < 1e-3


@pytest.mark.parametrize('m,n1,n2', TN_SHAPES)
def test_gemm_tn_mn_major(dev, m, n1, n2):
    """operands consumed in place as MN-major tcgen05 tiles (the product default, ops.TN_MODE = 0)"""
    _check_gemm_tn(dev, 0, m, n1, n2)


@pytest.mark.parametrize('m,n1,n2', TN_SHAPES)
def test_gemm_tn_k_major(dev, m, n1, n2):
    """alternative operand form: transposed copies + K-major tiles"""
    _check_gemm_tn(dev, 1, m, n1, n2)


def test_linear_ex(dev):
    from rvt_b200 import _lib, ops, packing
    m, k, n = 300, 256, 64
    g = torch.Generator(device='cpu').manual_seed(3)
    a = torch.randn(m, k, generator=g).to(dev).half()
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(dev)
    aux = torch.zeros(ops.round_up(m, 128), n, device=dev).half()
    aux[:m] = torch.randn(m, n, generator=g).to(dev).half()
    wp = packing.pack_linear_weight(w, _lib.lib().rvt_tile_n(n, k))
    ref = a.float() @ w.half().float().t()
    out32 = torch.empty(m, n, device=dev)
    ops.linear_ex(a, m, k, n, wp, out32)
    assert rel_l2(out32, ref) < 1e-3
    out16 = torch.empty(ops.round_up(m, 128), n, device=dev).half()
    ops.linear_ex(a, m, k, n, wp, out16, act=2, aux=aux)
    x = aux[:m].float().requires_grad_(True)
    F.gelu(x).sum().backward()
    assert rel_l2(out16[:m].float(), ref * x.grad) < 2e-3


MAPS = [(0, None), (1, (2, 3)), (2, (2, 3)), (1, (8, 10)), (2, (4, 5))]


def _map_index(b, h, w, mode, part):
    """row -> token index table (or -1) mirroring rvt_attention_scratch_rows' layout."""
    from rvt_b200 import _lib
    if mode == 0:
        return torch.arange(b * h * w)
    idx = bo.partition_index(h, w, part, mode == 1)            # [nG, P]
    ng, p = idx.shape
    rpg = _lib.lib().rvt_rows_per_group(p)
    rows = _lib.lib().rvt_attention_scratch_rows(b, h, w, part[0], part[1])
    out = torch.full((rows,), -1, dtype=torch.long)
    for bb in range(b):
        for gi in range(ng):
            r0 = (bb * ng + gi) * rpg
            out[r0:r0 + p] = bb * h * w + idx[gi]
    return out


@pytest.mark.parametrize('mode,part', MAPS)
@pytest.mark.parametrize('c', [32, 48, 256])
@pytest.mark.parametrize('do_ln', [True, False])
def test_ln_rows_and_bwd(dev, mode, part, c, do_ln):
    from rvt_b200 import ops
    b, h, w = 2, 24, 30
    g = torch.Generator(device='cpu').manual_seed(c + mode)
    x = torch.randn(b, h, w, c, generator=g).to(dev) * 1.5 + 0.3
    lw = (torch.rand(c, generator=g) + 0.5).to(dev)
    lb = (torch.randn(c, generator=g) * 0.1).to(dev)
    rmap = _map_index(b, h, w, mode, part).to(dev)
    rows = rmap.numel()
    valid = rmap >= 0
    # forward rows
    out16 = torch.full((rows, c), 7.0, device=dev).half()
    ops.ln_rows_f16(x, mode, part, lw, lb, do_ln, 1e-5, out16)
    xr = x.reshape(-1, c).clone().requires_grad_(True)
    lwr, lbr = lw.clone().requires_grad_(True), lb.clone().requires_grad_(True)
    y = F.layer_norm(xr, (c,), lwr, lbr, 1e-5) if do_ln else xr
    ref_rows = torch.zeros(rows, c, device=dev)
    ref_rows[valid] = y.detach()[rmap[valid]]
    assert rel_l2(out16.float(), ref_rows) < 1e-3
    # backward: dy given in row order
    dy_rows = torch.zeros(rows, c, device=dev)
    dy_rows[valid] = torch.randn(int(valid.sum()), c, generator=g).to(dev)
    dy16 = dy_rows.half()
    dy_tok = torch.zeros(b * h * w, c, device=dev)
    dy_tok[rmap[valid]] = dy16.float()[valid]
    y.backward(dy_tok)
    dres0 = torch.randn(b, h, w, c, generator=g).to(dev)
    dres = dres0.clone()
    dw, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    ops.ln_bwd(x if do_ln else None, dy16, (b, h, w, c), mode, part, lw if do_ln else None, do_ln, 1e-5, dres, None,
               dw if do_ln else None, db if do_ln else None)
    assert rel_l2(dres - dres0, xr.grad.reshape(b, h, w, c)) < 1e-4
    if do_ln:
        assert rel_l2(dw, lwr.grad) < 1e-4 and rel_l2(db, lbr.grad) < 1e-4
    if mode == 0 and do_ln:
        # conv-LN form: fp32 dy in token order -> fp16 dx rows
        dx16 = torch.empty(rows, c, device=dev).half()
        ops.ln_bwd(x, dy_tok.reshape(b, h, w, c).contiguous(), (b, h, w, c), 0, None, lw, True, 1e-5, None, dx16, None, None)
        assert rel_l2(dx16.float(), xr.grad) < 1e-3


@pytest.mark.parametrize('mode,part', MAPS)
def test_gather_cast(dev, mode, part):
    from rvt_b200 import ops
    b, h, w, c = 2, 24, 30, 64
    g = torch.Generator(device='cpu').manual_seed(5)
    dres = torch.randn(b, h, w, c, generator=g).to(dev)
    gamma = (torch.rand(c, generator=g) + 0.5).to(dev)
    rmap = _map_index(b, h, w, mode, part).to(dev)
    rows, valid = rmap.numel(), rmap >= 0
    d0 = torch.full((rows, c), 3.0, device=dev).half()
    d1 = torch.full((rows, c), 3.0, device=dev).half()
    ops.gather_cast(dres, mode, part, gamma, d0, d1)
    ref = torch.zeros(rows, c, device=dev)
    ref[valid] = dres.reshape(-1, c)[rmap[valid]]
    assert torch.equal(d0, ref.half())
    assert torch.equal(d1, (ref * gamma).half())


@pytest.mark.parametrize('part,dh,c', [((2, 3), 32, 64), ((8, 10), 32, 32), ((4, 5), 24, 48), ((6, 10), 32, 128)])
def test_attn_core_bwd(dev, part, dh, c):
    from rvt_b200 import _lib, ops
    b = 2
    h, w = part[0] * 3, part[1] * 2
    p = part[0] * part[1]
    nh = c // dh
    rpg = _lib.lib().rvt_rows_per_group(p)
    ng = b * 3 * 2
    rows = _lib.lib().rvt_attention_scratch_rows(b, h, w, part[0], part[1])
    g = torch.Generator(device='cpu').manual_seed(p)
    qkv = torch.zeros(rows, 3 * c)
    dout = torch.zeros(rows, c)
    for gi in range(ng):
        qkv[gi * rpg:gi * rpg + p] = torch.randn(p, 3 * c, generator=g)
        dout[gi * rpg:gi * rpg + p] = torch.randn(p, c, generator=g)
        qkv[gi * rpg + p:(gi + 1) * rpg] = 0.3           # padding rows hold the qkv bias in the real pipeline
    qkv16, dout16 = qkv.to(dev).half(), dout.to(dev).half()
    # reference: autograd over the valid tokens of every group
    q32 = qkv16.float().requires_grad_(True)
    loss = 0.0
    o_rows = torch.zeros(rows, c, device=dev)
    for gi in range(ng):
        blk = q32[gi * rpg:gi * rpg + p].view(p, nh, 3, dh)
        q, k, v = (blk[:, :, i].transpose(0, 1) for i in range(3))            # [nh, P, dh]
        a = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1)
        o = (a @ v).transpose(0, 1).reshape(p, c)
        o_rows[gi * rpg:gi * rpg + p] = o.detach()
        loss = loss + (o * dout16.float()[gi * rpg:gi * rpg + p]).sum()
    loss.backward()
    dqkv = torch.full((rows, 3 * c), 9.0, device=dev).half()
    ops.attn_core_bwd(qkv16, o_rows.half(), dout16, dqkv, (b, h, w, c), part, dh)
    n_valid_rows = ng * rpg
    assert rel_l2(dqkv.float()[:n_valid_rows], q32.grad[:n_valid_rows]) < 3e-3
    for gi in range(ng):
        assert float(dqkv[gi * rpg + p:(gi + 1) * rpg].float().abs().max()) == 0.0


def test_lstm_gates_bwd(dev):
    from rvt_b200 import ops
    n, c = 1000, 64
    g = torch.Generator(device='cpu').manual_seed(9)
    pre = torch.randn(n, 4 * c, generator=g).to(dev).requires_grad_(True)
    cp = torch.randn(n, c, generator=g).to(dev).requires_grad_(True)
    f, i, o = (torch.sigmoid(pre[:, j * c:(j + 1) * c]) for j in range(3))
    gg = torch.tanh(pre[:, 3 * c:])
    gates16 = torch.cat([f, i, o, gg], 1).detach().half()
    f, i, o, gg = (t.detach().half().float() + (t - t.detach()) for t in (f, i, o, gg))   # value = fp16-rounded gate, grad = identity
    cn = f * cp + i * gg
    hn = o * torch.tanh(cn)
    dh = torch.randn(n, c, generator=g).to(dev)
    dc = torch.randn(n, c, generator=g).to(dev)
    (hn * dh + cn * dc).sum().backward()
    dpre = torch.empty(ops.round_up(n, 128), 4 * c, device=dev).half()
    dcp = torch.empty(n, c, device=dev)
    ops.lstm_gates_bwd(gates16, cp.detach(), cn.detach(), dh, dc, n, c, dpre, dcp)
    assert rel_l2(dcp, cp.grad) < 1e-5
    assert rel_l2(dpre[:n].float(), pre.grad) < 2e-3
    # zero state / no dc
    dpre2 = torch.empty_like(dpre)
    ops.lstm_gates_bwd(gates16, None, cn.detach(), dh, None, n, c, dpre2, None)
    assert torch.isfinite(dpre2[:n].float()).all()


@pytest.mark.parametrize('nchw,dtype,cin,ks,stride', [(True, torch.uint8, 20, 7, 4), (True, torch.float32, 20, 7, 4),
                                                      (False, torch.float32, 32, 3, 2), (False, torch.float32, 48, 3, 2)])
def test_im2col_col2im(dev, nchw, dtype, cin, ks, stride):
    from rvt_b200 import ops
    b, hin, win = 2, 24, 40
    pad = ks // 2
    hout, wout = (hin + 2 * pad - ks) // stride + 1, (win + 2 * pad - ks) // stride + 1
    g = torch.Generator(device='cpu').manual_seed(cin)
    xn = torch.randint(0, 11, (b, cin, hin, win), generator=g).float()
    if dtype != torch.uint8:
        xn = xn + torch.randn(b, cin, hin, win, generator=g)
    x = (xn.to(dtype) if nchw else xn.permute(0, 2, 3, 1).contiguous().to(dtype)).to(dev)
    k = ks * ks * cin
    ldc = ops.round_up(k, 8)
    col = torch.full((b * hout * wout, ldc), 5.0, device=dev).half()
    ops.im2col(x, nchw, ks, stride, pad, hout, wout, col)
    unf = F.unfold(xn.to(dtype).float().to(dev), ks, padding=pad, stride=stride)          # [B, cin*ks*ks, L]  (ci, ky, kx)
    if nchw:
        ref = unf.permute(0, 2, 1).reshape(b * hout * wout, k)                                      # (ci, ky, kx): the stem's order
    else:
        ref = unf.view(b, cin, ks * ks, hout * wout).permute(0, 3, 2, 1).reshape(b * hout * wout, k)   # (ky, kx, ci)
    assert rel_l2(col[:, :k].float(), ref.half().float()) == 0.0
    assert float(col[:, k:].float().abs().max() if ldc > k else 0.0) == 0.0
    if not nchw:
        dcol = torch.randn(b * hout * wout, ldc, generator=g).to(dev).half()
        d_in = torch.empty(b, hin, win, cin, device=dev)
        ops.col2im(dcol, b, cin, hin, win, ks, stride, pad, hout, wout, d_in)
        cols = dcol[:, :k].float().view(b, hout * wout, ks * ks, cin).permute(0, 3, 2, 1).reshape(b, cin * ks * ks, hout * wout)
        ref_in = F.fold(cols, (hin, win), ks, padding=pad, stride=stride).permute(0, 2, 3, 1)
        assert rel_l2(d_in, ref_in) < 1e-5


def test_colsum(dev):
    from rvt_b200 import ops
    m, n = 3000, 192
    a = torch.randn(m + 50, n, device=dev).half()
    acc = torch.ones(n, device=dev)
    ops.colsum(a, m, n, acc)
    assert rel_l2(acc - 1, a[:m].float().sum(0)) < 1e-4


def test_nchw_to_nhwc_f16(dev):
    from rvt_b200 import ops
    x = torch.randint(0, 11, (2, 20, 13, 36), dtype=torch.uint8, device=dev)
    out = torch.full((2, 13, 36, 24), 5.0, device=dev).half()
    ops.nchw_to_nhwc_f16(x, 24, out)
    assert torch.equal(out[..., :20], x.permute(0, 2, 3, 1).half()) and float(out[..., 20:].abs().max()) == 0.0

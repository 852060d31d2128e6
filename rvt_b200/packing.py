"""Host-side weight packing into the shared-memory tile images the tcgen05 kernels consume.

A GEMM weight W[N, K] is cut into n-tiles of BN rows (BN = rvt_tile_n(N) unless given) and
K-chunks of 64; each (n-tile, k-chunk) block is stored contiguously as the exact smem image of
a 'SW128 K-major' UMMA operand tile (csrc/umma.cuh): row r at byte r*128, its eight 16-byte
chunks XOR-permuted with (r & 7).  One 1-D bulk async copy then lands a tile in smem.
Layout: fp16 [n_tiles][KC][BN][64].
"""
import torch


def _swizzle_tiles(w2: torch.Tensor, bn: int) -> torch.Tensor:
    n, k = w2.shape
    assert n % bn == 0 and bn % 8 == 0
    kc = (k + 63) // 64
    wp = torch.zeros(n, kc * 64, dtype=torch.float16, device=w2.device)
    wp[:, :k] = w2.to(torch.float16)
    t = wp.view(n // bn, bn, kc, 8, 8).permute(0, 2, 1, 3, 4)          # [nt, kc, r, chunk, 8]
    r = torch.arange(bn, device=w2.device).view(bn, 1)
    p = torch.arange(8, device=w2.device).view(1, 8)
    src = (p ^ (r & 7)).view(1, 1, bn, 8, 1).expand(n // bn, kc, bn, 8, 8)
    return torch.gather(t, 3, src).contiguous()                        # out[.., r, p, :] = in[.., r, p^(r&7), :]


def pack_linear_weight(w: torch.Tensor, bn: int) -> torch.Tensor:
    """nn.Linear weight [N, K] -> tile images; n-tile j holds rows [j*bn, (j+1)*bn)."""
    assert w.dim() == 2
    return _swizzle_tiles(w.detach().float(), bn)


def pack_conv_weight(w: torch.Tensor, channels_last_input: bool, bn: int = 0) -> torch.Tensor:
    """Conv2d weight [Cout, Cin, KS, KS] -> n-tiles of bn rows (default: one tile of Cout rows).
    K order matches the im2col loader: (ci, ky, kx) for NCHW input, (ky, kx, ci) for NHWC."""
    assert w.dim() == 4
    co = w.shape[0]
    w2 = w.detach().float().permute(0, 2, 3, 1).reshape(co, -1) if channels_last_input \
        else w.detach().float().reshape(co, -1)
    return _swizzle_tiles(w2, bn or co)


def lstm_row_order(dim: int, cw: int, device=None) -> torch.Tensor:
    """Row permutation of the [4C, 2C] gate weight: tile j = [f|i|o|g] x channels [j*cw,(j+1)*cw).
    Built on `device` directly (no host->device copy, so weight re-packing can be captured in a CUDA graph)."""
    j = torch.arange(dim // cw, device=device).view(-1, 1, 1)
    g = torch.arange(4, device=device).view(1, 4, 1)
    c = torch.arange(cw, device=device).view(1, 1, cw)
    return (g * dim + j * cw + c).reshape(-1)


def pack_lstm_weight(w: torch.Tensor, b: torch.Tensor, cw: int):
    """conv1x1 weight [4C, 2C, 1, 1] + bias [4C] -> (tile images with BN = 4*cw, tiled bias)."""
    dim = w.shape[0] // 4
    w2 = w.detach().float().reshape(4 * dim, 2 * dim)
    order = lstm_row_order(dim, cw, w.device)
    return _swizzle_tiles(w2[order], 4 * cw), b.detach().float()[order].contiguous()


def pack_dw_weight(w: torch.Tensor) -> torch.Tensor:
    """depthwise weight [D, 1, ks, ks] -> tap-major [ks*ks, D] fp32."""
    d, _, ks, _ = w.shape
    return w.detach().float().reshape(d, ks * ks).t().contiguous()


def pack_stem_weight_s2d(w: torch.Tensor, factor: int) -> torch.Tensor:
    """Stem conv weight [Cout, Cin, KS, KS] for the space-to-depth input layout
    [B, H, W/f, f*Cin] (csrc/gemm_fused.cuh stem_s2d_kernel): K order (ky, t, sub, ci) where the
    x-tap t in {0,1} selects pixel group ox-1+t and sub the pixel inside the group.
    Overlapping stem (KS = 2f-1, pad f-1): kx = sub-1 for t=0 (sub>=1), kx = sub+f-1 for t=1.
    Patch stem (KS = f, pad 0): single x-tap, kx = sub."""
    co, cin, ks, _ = w.shape
    f = factor
    wf = w.detach().float()
    if ks == 2 * f - 1:
        w2 = torch.zeros(co, ks, 2, f, cin, device=w.device)
        for sub in range(1, f):
            w2[:, :, 0, sub, :] = wf[:, :, :, sub - 1].permute(0, 2, 1)
        for sub in range(f):
            w2[:, :, 1, sub, :] = wf[:, :, :, sub + f - 1].permute(0, 2, 1)
    elif ks == f:
        w2 = wf.permute(0, 2, 3, 1).reshape(co, ks, 1, f, cin)
    else:
        raise ValueError('unsupported stem geometry')
    return _swizzle_tiles(w2.reshape(co, -1), co)


def pack_qkv_weight(w: torch.Tensor, b, dim_head: int, dhp: int = 32):
    """qkv Linear weight [3C, C] (+ bias [3C] or None) for the fused attention kernel
    (csrc/attn_fused.cuh): one N-tile of 3*dhp rows per head, [q_h | k_h | v_h] each zero-padded
    from dim_head to dhp rows (the reference's per-head interleaved layout, maxvit.py:347).
    Returns (tile images [nh][KC][3*dhp x 64], padded bias [nh * 3*dhp] fp32)."""
    c3, c = w.shape
    nh = c // dim_head
    assert c3 == 3 * c and dim_head <= dhp
    wf = w.detach().float().view(nh, 3, dim_head, c)
    wpad = torch.zeros(nh, 3, dhp, c, device=w.device)
    wpad[:, :, :dim_head] = wf
    bpad = torch.zeros(nh, 3, dhp, device=w.device)
    if b is not None:
        bpad[:, :, :dim_head] = b.detach().float().view(nh, 3, dim_head)
    return _swizzle_tiles(wpad.reshape(nh * 3 * dhp, c), 3 * dhp), bpad.reshape(-1).contiguous()


STEM_KY_ORDER = (0, 4, 1, 5, 2, 6, 3)


def pack_stem_weight_u8(w: torch.Tensor) -> torch.Tensor:
    """7x7 / stride-4 stem weight [Cout, Cin, 7, 7] for the uint8 smem-patch loader (LD_STEM):
    K order (kyi, ci, kx8), ky = STEM_KY_ORDER[kyi], where kx8 indexes the 8 input bytes [4*ox-4, 4*ox+4) of a row, i.e. kx8 = kx + 1
    and kx8 = 0 carries a zero weight."""
    co, cin, ks, _ = w.shape
    assert ks == 7
    w2 = torch.zeros(co, ks, cin, 8, device=w.device)
    w2[:, :, :, 1:] = w.detach().float().permute(0, 2, 1, 3)
    # kernel rows in the order 0, 4, 1, 5, 2, 6, 3 (csrc/gemm_fused.cuh stem_ky): rows that read the same input-row phase are adjacent
    w2 = torch.stack([w2[:, k] for k in STEM_KY_ORDER], dim=1)      # no index tensor: re-packing must stay capturable in a CUDA graph
    return _swizzle_tiles(w2.reshape(co, -1), co)

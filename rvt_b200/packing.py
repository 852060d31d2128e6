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


def pack_conv_weight(w: torch.Tensor, channels_last_input: bool) -> torch.Tensor:
    """Conv2d weight [Cout, Cin, KS, KS] -> one n-tile of Cout rows.
    K order matches the im2col loader: (ci, ky, kx) for NCHW input, (ky, kx, ci) for NHWC."""
    assert w.dim() == 4
    co = w.shape[0]
    w2 = w.detach().float().permute(0, 2, 3, 1).reshape(co, -1) if channels_last_input \
        else w.detach().float().reshape(co, -1)
    return _swizzle_tiles(w2, co)


def lstm_row_order(dim: int, cw: int) -> torch.Tensor:
    """Row permutation of the [4C, 2C] gate weight: tile j = [f|i|o|g] x channels [j*cw,(j+1)*cw)."""
    j = torch.arange(dim // cw).view(-1, 1, 1)
    g = torch.arange(4).view(1, 4, 1)
    c = torch.arange(cw).view(1, 1, cw)
    return (g * dim + j * cw + c).reshape(-1)


def pack_lstm_weight(w: torch.Tensor, b: torch.Tensor, cw: int):
    """conv1x1 weight [4C, 2C, 1, 1] + bias [4C] -> (tile images with BN = 4*cw, tiled bias)."""
    dim = w.shape[0] // 4
    w2 = w.detach().float().reshape(4 * dim, 2 * dim)
    order = lstm_row_order(dim, cw).to(w.device)
    return _swizzle_tiles(w2[order], 4 * cw), b.detach().float()[order].contiguous()


def pack_dw_weight(w: torch.Tensor) -> torch.Tensor:
    """depthwise weight [D, 1, ks, ks] -> tap-major [ks*ks, D] fp32."""
    d, _, ks, _ = w.shape
    return w.detach().float().reshape(d, ks * ks).t().contiguous()

"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md); never imported by rvt_b200/.

numpy restatement of ``StackedHistogram.construct``
(/root/reference/data/utils/representations.py:76-121).  Integer/byte work: the
parity bar against the CUDA voxelizer is bit-exact.

Semantics restated:
  * time bin: torch true-divides two int64 tensors -> both operands are converted to
    float32 FIRST (round-to-nearest-even), then an IEEE fp32 divide, an fp32 multiply by
    ``bins``, floor, clamp(max=bins-1)                       (representations.py:102-109)
  * flat index = x + W*y + H*W*t_idx + bins*H*W*pol  (polarity-major) (:111-114)
  * fastmode: accumulate in uint8 with wrap-around mod 256, then clamp to count_cutoff;
    else accumulate in int16 (wraps mod 2^16 as two's complement), clamp to [0, cutoff]
    and cast to uint8                                          (:84,115-119)
  * empty input -> zeros                                       (:89-93)
Pinned against the imported reference by oracle/make_golden.py -> tests/golden/voxel_*.npz.
"""
import numpy as np


def time_bin_index(t: np.ndarray, bins: int) -> np.ndarray:
    t = np.asarray(t, dtype=np.int64)
    t0, t1 = t[0], t[-1]
    assert t1 >= t0
    denom = np.float32(max(int(t1 - t0), 1))
    tn = (t - t0).astype(np.float32) / denom            # fp32 divide
    tn = tn * np.float32(bins)                          # fp32 multiply
    ti = np.floor(tn)
    ti = np.minimum(ti, np.float32(bins - 1))
    return ti.astype(np.int64)


def stacked_histogram(x, y, pol, t, bins: int, height: int, width: int,
                      count_cutoff=None, fastmode: bool = True) -> np.ndarray:
    cutoff = 255 if count_cutoff is None else min(int(count_cutoff), 255)
    n_out = 2 * bins * height * width
    x = np.asarray(x, dtype=np.int64)
    if x.size == 0:
        return np.zeros((2 * bins, height, width), np.uint8)
    y = np.asarray(y, dtype=np.int64)
    pol = np.asarray(pol, dtype=np.int64)
    assert pol.min() >= 0 and pol.max() <= 1
    ti = time_bin_index(t, bins)
    idx = x + width * y + height * width * ti + bins * height * width * pol
    counts = np.bincount(idx, minlength=n_out).astype(np.int64)
    if fastmode:
        rep = (counts & 0xFF).astype(np.int64)          # uint8 wrap-around
        rep = np.minimum(rep, cutoff)
    else:
        rep = ((counts + 0x8000) & 0xFFFF) - 0x8000     # int16 two's-complement wrap
        rep = np.clip(rep, 0, cutoff)
    return rep.astype(np.uint8).reshape(2 * bins, height, width)


def synth_events(seed: int, n: int, height: int, width: int, t_span: int = 50000,
                 hot_fraction: float = 0.0, hot_pixels: int = 16):
    """x~U[0,W) y~U[0,H) p~Bern(.5), t sorted U[0,t_span) int64;
    optional hot-pixel variant (a fraction of events on a few pixels)."""
    rs = np.random.RandomState(seed)
    x = rs.randint(0, width, n).astype(np.int64)
    y = rs.randint(0, height, n).astype(np.int64)
    p = rs.randint(0, 2, n).astype(np.int64)
    t = np.sort(rs.randint(0, t_span, n).astype(np.int64))
    if hot_fraction > 0:
        m = rs.uniform(size=n) < hot_fraction
        hx = rs.randint(0, width, hot_pixels)
        hy = rs.randint(0, height, hot_pixels)
        sel = rs.randint(0, hot_pixels, n)
        x = np.where(m, hx[sel], x)
        y = np.where(m, hy[sel], y)
    return x, y, p, t

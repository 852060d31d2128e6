"""ORACLE — TEST INFRASTRUCTURE ONLY; never imported by rvt_b200/.

CPU restatement of the hot path's neighbours in the reference, pinned against the reference's own code
by oracle/make_golden.py (tests/golden/neigh_*.npz):

  downsample_ev_repr     scripts/genx/preprocess_dataset.py:467-477   (F.interpolate, scale 0.5, 'nearest-exact')
  correct_time           scripts/genx/preprocess_dataset.py:163-172   (H5Reader._correct_time)
  event_window_indices   scripts/genx/preprocess_dataset.py:511-516
  mixed_density_stack    data/utils/representations.py:130-218        (MixedDensityEventStack.construct)
  reset_states / select_features   modules/utils/detection.py:96-113 / :24-46 (RNNStates.reset, BackboneFeatureSelector)

numpy for the integer work; the one floating-point expression (the logarithmic time bin of MixedDensityEventStack) is evaluated
with torch CPU fp32 ops because the reference's arithmetic IS those ops (bit-exact parity needs the same log)."""
import math

import numpy as np
import torch


def downsample_ev_repr(x: np.ndarray) -> np.ndarray:
    """[..., H, W] -> [..., H//2, W//2]; 'nearest-exact' with scale 0.5 reads source pixel floor((dst + 0.5) * 2) = 2*dst + 1."""
    h, w = x.shape[-2], x.shape[-1]
    return np.ascontiguousarray(x[..., 1:2 * (h // 2):2, 1:2 * (w // 2):2])


def correct_time(t: np.ndarray) -> np.ndarray:
    """time_last = 0; every timestamp below the running maximum is replaced by it (preprocess_dataset.py:165-172)."""
    t = np.asarray(t, dtype=np.int64)
    assert t.size == 0 or t[0] >= 0
    return np.maximum.accumulate(np.maximum(t, 0))


def event_window_indices(ev_ts_us, ev_repr_timestamps_us, ev_repr_num_events=None, ev_repr_delta_ts_ms=None):
    end = np.searchsorted(ev_ts_us, ev_repr_timestamps_us, side='right')
    if ev_repr_num_events is not None:
        start = np.maximum(end - ev_repr_num_events, 0)
    else:
        start = np.searchsorted(ev_ts_us, ev_repr_timestamps_us - ev_repr_delta_ts_ms * 1000, side='left')
    return start, end


def mixed_density_time_bin(t: np.ndarray, bins: int) -> np.ndarray:
    """representations.py:178-203 in torch CPU fp32 (int64 / int64 true-divide -> float32, clamp, log, floor)."""
    tt = torch.from_numpy(np.asarray(t, dtype=np.int64))
    t0, t1 = tt[0], tt[-1]
    t_norm = (tt - t0) / max((t1 - t0), 1)
    t_norm = torch.clamp(t_norm, min=1e-6, max=1 - 1e-6)
    bin_float = bins - torch.log(t_norm) / math.log(1 / 2)
    bin_float = torch.clamp(bin_float, min=0)
    return bin_float.floor().long().numpy()


def mixed_density_stack(x, y, pol, t, bins: int, height: int, width: int, count_cutoff=None) -> np.ndarray:
    x = np.asarray(x, dtype=np.int64)
    if x.size == 0:
        return np.zeros((bins, height, width), np.int8)
    y, pol = np.asarray(y, dtype=np.int64), np.asarray(pol, dtype=np.int64)
    assert pol.min() >= 0 and pol.max() <= 1
    ti = mixed_density_time_bin(t, bins)
    idx = x + width * y + height * width * ti
    n_out = bins * height * width
    raw = (np.bincount(idx, weights=None, minlength=n_out)[:n_out] * 0).astype(np.int64)
    np.add.at(raw, idx, pol * 2 - 1)                      # put_(accumulate=True) of +-1 (int8 wrap applied below: mod-256 ring)
    rep = raw.reshape(bins, height * width)
    cum = np.cumsum(rep, axis=0)                          # x[i] = sum(x[:i+1]) on the ORIGINAL channels (reversed loop, :122-125)
    cum = ((cum + 128) & 0xFF) - 128                      # int64 -> int8 cast wraps
    if count_cutoff is not None:
        cum = np.clip(cum, -count_cutoff, count_cutoff)
    return cum.astype(np.int8).reshape(bins, height, width)


def reset_states(states, mask: np.ndarray):
    """state[mask] = 0 for every (h, c) of every stage (RNNStates.recursive_reset with a bool tensor)."""
    out = []
    for st in states:
        if st is None:
            out.append(None)
            continue
        h, c = (a.copy() for a in st)
        h[mask] = 0
        c[mask] = 0
        out.append((h, c))
    return out


def select_features(feats_per_step, selected_per_step):
    """BackboneFeatureSelector: per step with labels append v[selected_indices]; finally cat along dim 0."""
    acc = {}
    for feats, sel in zip(feats_per_step, selected_per_step):
        if sel is None or len(sel) == 0:
            continue
        for k, v in feats.items():
            acc.setdefault(k, []).append(v[sel])
    return {k: np.concatenate(v, axis=0) for k, v in acc.items()} if acc else None

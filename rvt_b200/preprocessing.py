"""GPU versions of the voxelizer's neighbours in the reference's offline preprocessing
(``scripts/genx/preprocess_dataset.py``), so raw events -> model tensor can stay on the device:

  downsample_ev_repr   :467-477, used at :525-528   nearest-exact x0.5 of a [.., C, H, W] uint8 / int8 representation
  correct_time_        :163-172 (H5Reader._correct_time, numba there)   running maximum, in place
  event_window_indices :511-516   [idx_start, idx_end) of every representation window (np.searchsorted)

CUDA tensors only (C-ABI kernels in csrc/neighbours.cuh); no CPU fallback."""
from typing import Optional, Tuple

import torch

from . import _lib


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f'rvt_b200.preprocessing.{what} runs on CUDA (sm_100a) only; there is no CPU fallback')


def downsample_ev_repr(x: torch.Tensor, scale_factor: float = 0.5) -> torch.Tensor:
    """``downsample_ev_repr(x, scale_factor)`` of the reference for its only call site value 0.5: x is [N, C, H, W] (or [C, H, W])
    uint8 / int8; returns the same rank with H//2, W//2 (F.interpolate 'nearest-exact': out[y, x] = in[2y+1, 2x+1])."""
    _need_cuda(x, 'downsample_ev_repr')
    if scale_factor != 0.5:
        raise NotImplementedError('only scale_factor=0.5 (the reference\'s downsample_by_2) is built')
    assert x.dtype in (torch.uint8, torch.int8) and x.dim() in (3, 4)
    shape = x.shape
    h, w = shape[-2], shape[-1]
    c = x.numel() // (h * w)
    xc = x.contiguous()
    out = torch.empty(shape[:-2] + (h // 2, w // 2), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().rvt_downsample2_nearest(_lib.ptr(xc), c, h, w, _lib.ptr(out), _stream(x)), 'downsample2_nearest')
    return out


def correct_time_(time_array: torch.Tensor) -> torch.Tensor:
    """In place ``H5Reader._correct_time``: timestamps become non-decreasing (each replaced by the running maximum, floor 0)."""
    _need_cuda(time_array, 'correct_time_')
    assert time_array.dtype == torch.int64 and time_array.dim() == 1 and time_array.is_contiguous()
    n = time_array.numel()
    if n == 0:
        return time_array
    L = _lib.lib()
    scratch = torch.empty(L.rvt_cummax_scratch_elems(n), dtype=torch.int64, device=time_array.device)
    with torch.cuda.device(time_array.device):
        _lib.check(L.rvt_cummax_i64(_lib.ptr(time_array), n, 0, _lib.ptr(scratch), _stream(time_array)), 'cummax_i64')
    return time_array


def searchsorted(sorted_ts: torch.Tensor, queries: torch.Tensor, side: str = 'left') -> torch.Tensor:
    _need_cuda(sorted_ts, 'searchsorted')
    assert sorted_ts.dtype == torch.int64 and queries.dtype == torch.int64 and side in ('left', 'right')
    a, q = sorted_ts.contiguous(), queries.contiguous().to(sorted_ts.device)
    out = torch.empty(q.shape, dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().rvt_searchsorted_i64(_lib.ptr(a), a.numel(), _lib.ptr(q), q.numel(), int(side == 'right'),
                                                   _lib.ptr(out), _stream(a)), 'searchsorted_i64')
    return out


def event_window_indices(ev_ts_us: torch.Tensor, ev_repr_timestamps_us: torch.Tensor, ev_repr_num_events: Optional[int] = None,
                         ev_repr_delta_ts_ms: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(start_indices, end_indices) of ``write_event_representations`` (preprocess_dataset.py:511-516)."""
    end = searchsorted(ev_ts_us, ev_repr_timestamps_us, side='right')
    if ev_repr_num_events is not None:
        start = torch.clamp(end - ev_repr_num_events, min=0)
    else:
        assert ev_repr_delta_ts_ms is not None
        start = searchsorted(ev_ts_us, ev_repr_timestamps_us.to(ev_ts_us.device) - ev_repr_delta_ts_ms * 1000, side='left')
    return start, end

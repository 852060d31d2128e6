"""Drop-in mirror of ``data/utils/representations.py::StackedHistogram`` (reference :37-121):
same constructor, ``construct(x, y, pol, time) -> uint8 [2*bins, H, W]``, ``get_shape``,
``get_numpy_dtype`` / ``get_torch_dtype`` / ``dtype``.  The histogram is built by the sm_100a
warp-aggregated atomic kernel behind the C-ABI (``rvt_stacked_histogram``); CUDA tensors only,
no CPU fallback.  Output is bit-exact with the reference (tests/test_gpu_voxel.py)."""
from typing import Optional, Tuple

import numpy as np
import torch as th

from . import _lib


class StackedHistogram:
    def __init__(self, bins: int, height: int, width: int, count_cutoff: Optional[int] = None,
                 fastmode: bool = True, validate: bool = True):
        assert bins >= 1
        self.bins = bins
        assert height >= 1
        self.height = height
        assert width >= 1
        self.width = width
        self.count_cutoff = count_cutoff
        if self.count_cutoff is None:
            self.count_cutoff = 255
        else:
            assert count_cutoff >= 1
            self.count_cutoff = min(count_cutoff, 255)
        self.fastmode = fastmode
        self.channels = 2
        # validate=True reproduces the reference's asserts (pol in {0,1}, sorted time) at the price
        # of one 4-byte device->host read per call; streaming pipelines may pass False.
        self.validate = validate
        self._counts = None
        self._err = None

    @staticmethod
    def get_numpy_dtype() -> np.dtype:
        return np.dtype('uint8')

    @staticmethod
    def get_torch_dtype() -> th.dtype:
        return th.uint8

    @property
    def dtype(self) -> th.dtype:
        return self.get_torch_dtype()

    def get_shape(self) -> Tuple[int, int, int]:
        return 2 * self.bins, self.height, self.width

    @staticmethod
    def _is_int_tensor(tensor: th.Tensor) -> bool:
        return not th.is_floating_point(tensor) and not th.is_complex(tensor)

    @staticmethod
    def _aligned16(t: th.Tensor) -> th.Tensor:
        return t.clone() if (t.numel() and t.data_ptr() % 16) else t

    def _buffers(self, device):
        n_out = 2 * self.bins * self.height * self.width
        if self._counts is None or self._counts.device != device:
            self._counts = th.zeros(n_out, dtype=th.int32, device=device)   # u32 scratch, kept zero between calls
            self._err = th.zeros(1, dtype=th.int32, device=device)
        return self._counts, self._err

    def construct(self, x: th.Tensor, y: th.Tensor, pol: th.Tensor, time: th.Tensor,
                  out: Optional[th.Tensor] = None) -> th.Tensor:
        device = x.device
        assert y.device == pol.device == time.device == device
        if device.type != 'cuda':
            raise RuntimeError('rvt_b200.StackedHistogram runs on CUDA (sm_100a) only; there is no CPU fallback')
        for t in (x, y, pol, time):
            assert self._is_int_tensor(t)
        assert x.numel() == y.numel() == pol.numel() == time.numel()
        # the kernel reads the event arrays as 16-byte vectors: a slice such as x[1:] is only 8-byte aligned -> copy it
        x, y, pol, time = (self._aligned16(t.to(th.int64).contiguous()) for t in (x, y, pol, time))
        if device.index is not None and device.index != th.cuda.current_device():
            with th.cuda.device(device):          # the C-ABI launches on the current device
                return self.construct(x, y, pol, time, out)
        counts, err = self._buffers(device)
        if out is None:
            out = th.empty(self.get_shape(), dtype=th.uint8, device=device)
        else:
            assert out.dtype == th.uint8 and out.is_contiguous() and tuple(out.shape) == self.get_shape()
        L = _lib.lib()
        stream = th.cuda.current_stream(device).cuda_stream
        _lib.check(L.rvt_stacked_histogram(
            _lib.ptr(x), _lib.ptr(y), _lib.ptr(pol), _lib.ptr(time), x.numel(), self.bins, self.height, self.width,
            self.count_cutoff, int(self.fastmode), _lib.ptr(counts), _lib.ptr(out), _lib.ptr(err), stream),
            'stacked_histogram')
        if self.validate:
            code = int(err.item())
            if code:
                err.zero_()
                counts.zero_()
                raise AssertionError(f'StackedHistogram.construct: invalid events (flags {code}: '
                                     f'1=time not sorted, 2=pol not in {{0,1}}, 4=coordinate outside frame)')
        return out

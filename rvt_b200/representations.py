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


class MixedDensityEventStack:
    """Drop-in mirror of ``data/utils/representations.py::MixedDensityEventStack`` (reference :130-218): same constructor,
    ``construct(x, y, pol, time) -> int8 [bins, H, W]``, ``get_shape`` / dtype accessors.  Built by the CUDA kernels behind
    ``rvt_mixed_density_stack``; CUDA tensors only, no CPU fallback.

    Bit-exactness of the logarithmic time binning: the reference computes ``floor(clamp(bins - log(t_norm) / log(1/2), 0))``
    in torch fp32 on the CPU.  Instead of re-deriving a device ``logf`` that rounds like torch's, the ``bins - 1`` fp32 thresholds
    at which that very expression steps to the next bin are located once per instance by bisection over float32 bit patterns,
    evaluating the reference's own expression with torch CPU ops (host-side constants, like a lookup table); the kernel then
    only compares ``t_norm`` against them."""

    def __init__(self, bins: int, height: int, width: int, count_cutoff: Optional[int] = None,
                 allow_compilation: bool = False, validate: bool = True):
        assert bins >= 1
        self.bins = bins
        assert height >= 1
        self.height = height
        assert width >= 1
        self.width = width
        self.count_cutoff = count_cutoff
        if self.count_cutoff is not None:
            assert isinstance(count_cutoff, int)
            assert 0 <= self.count_cutoff <= 2 ** 7 - 1
        self.validate = validate
        self._lo = float(th.tensor(1e-6, dtype=th.float32))             # th.clamp(t_norm, min=1e-6, max=1 - 1e-6) in fp32
        self._hi = float(th.tensor(1 - 1e-6, dtype=th.float32))
        self._thr_host = self._bin_thresholds(bins, self._lo, self._hi)
        self._thr = None
        self._counts = None
        self._err = None

    @staticmethod
    def _bin_of(t: th.Tensor, bins: int) -> th.Tensor:
        """the reference's expression (representations.py:186-203), torch CPU fp32"""
        import math
        bin_float = bins - th.log(t) / math.log(1 / 2)
        bin_float = th.clamp(bin_float, min=0)
        return bin_float.floor()

    @classmethod
    def _bin_thresholds(cls, bins: int, lo: float, hi: float) -> th.Tensor:
        lo_bits = int(th.tensor(lo, dtype=th.float32).view(th.int32))
        hi_bits = int(th.tensor(hi, dtype=th.float32).view(th.int32))
        f = lambda bits: float(cls._bin_of(th.tensor(bits, dtype=th.int32).view(th.float32), bins))
        thr = []
        for k in range(1, bins):
            if f(hi_bits) < k:                       # bin k never reached inside the clamp range
                thr.append(float('inf'))
                continue
            a, b = lo_bits, hi_bits                  # smallest bit pattern in [lo, hi] with bin >= k (positive floats: monotone in bits)
            if f(a) >= k:
                b = a
            while a < b:
                m = (a + b) // 2
                if f(m) >= k:
                    b = m
                else:
                    a = m + 1
            # the located step must be a step of a monotone function in its neighbourhood
            near = th.arange(max(lo_bits, b - 64), min(hi_bits, b + 64) + 1, dtype=th.int32)
            vals = cls._bin_of(near.view(th.float32), bins)
            assert bool((vals[1:] >= vals[:-1]).all()), 'reference time binning is not monotone near a bin edge'
            thr.append(float(th.tensor(b, dtype=th.int32).view(th.float32)))
        return th.tensor(thr, dtype=th.float32)

    @staticmethod
    def get_numpy_dtype() -> np.dtype:
        return np.dtype('int8')

    @staticmethod
    def get_torch_dtype() -> th.dtype:
        return th.int8

    @property
    def dtype(self) -> th.dtype:
        return self.get_torch_dtype()

    def get_shape(self) -> Tuple[int, int, int]:
        return self.bins, self.height, self.width

    def construct(self, x: th.Tensor, y: th.Tensor, pol: th.Tensor, time: th.Tensor) -> th.Tensor:
        device = x.device
        assert y.device == pol.device == time.device == device
        if device.type != 'cuda':
            raise RuntimeError('rvt_b200.MixedDensityEventStack runs on CUDA (sm_100a) only; there is no CPU fallback')
        for t in (x, y, pol, time):
            assert StackedHistogram._is_int_tensor(t)
        assert x.numel() == y.numel() == pol.numel() == time.numel()
        x, y, pol, time = (t.to(th.int64).contiguous() for t in (x, y, pol, time))
        if device.index is not None and device.index != th.cuda.current_device():
            with th.cuda.device(device):
                return self.construct(x, y, pol, time)
        n_out = self.bins * self.height * self.width
        if self._counts is None or self._counts.device != device:
            self._counts = th.zeros(n_out, dtype=th.int32, device=device)
            self._err = th.zeros(1, dtype=th.int32, device=device)
            self._thr = self._thr_host.to(device) if self.bins > 1 else None
        out = th.empty(self.get_shape(), dtype=th.int8, device=device)
        L = _lib.lib()
        _lib.check(L.rvt_mixed_density_stack(
            _lib.ptr(x), _lib.ptr(y), _lib.ptr(pol), _lib.ptr(time), x.numel(), self.bins, self.height, self.width,
            -1 if self.count_cutoff is None else self.count_cutoff, self._lo, self._hi, _lib.ptr(self._thr),
            _lib.ptr(self._counts), _lib.ptr(out), _lib.ptr(self._err), th.cuda.current_stream(device).cuda_stream),
            'mixed_density_stack')
        if self.validate:
            code = int(self._err.item())
            if code:
                self._err.zero_()
                self._counts.zero_()
                raise AssertionError(f'MixedDensityEventStack.construct: invalid events (flags {code}: '
                                     f'1=time not sorted, 2=pol not in {{0,1}}, 4=coordinate outside frame)')
        return out

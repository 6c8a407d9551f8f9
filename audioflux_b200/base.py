"""Common plumbing of the host-side mirror classes (reference: python/audioflux/base.py:4-8)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _libmod

MEM_HOST, MEM_DEVICE = 0, 1


class Base(object):
    def __init__(self, _lib=None):
        self._lib = _libmod.get_lib() if _lib is None else _lib
        self._obj = C.c_void_p()
        self._is_created = False
        self._is_product = _lib is None or hasattr(self._lib, "afb200_version")

    def _require_ext(self, name):
        if not hasattr(self._lib, name):
            raise AttributeError(f"library does not export the additive entry point {name} "
                                 f"(include/afb200_ext.h); it is not libaudioflux_b200")
        return getattr(self._lib, name)


class BandAxis:
    """Plot-axis helper of the banded transforms (reference: the y_coords of bft.py / cqt.py / cwt.py / pwt.py / wsst.py):
    the band frequencies with the lower edge in front."""

    def y_coords(self):
        return np.concatenate(([self.low_fre], self.get_fre_band_arr()))


class FrameAxis:
    """Time axis of a framed transform: data_length / samplate seconds cut into cal_time_length frames."""
    _needs_full_frame = True

    def x_coords(self, data_length):
        if self._needs_full_frame and data_length < self.fft_length:
            raise ValueError(f"radix2_exp={self.radix2_exp}(fft_length={self.fft_length}) is too large for data_length={data_length}")
        return np.linspace(0, data_length / self.samplate, self.cal_time_length(data_length) + 1)


class SampleAxis:
    """Time axis of a per-sample transform (CWT family): one column per sample of the 2**radix2_exp window."""

    def x_coords(self):
        return np.linspace(0, self.fft_length / self.samplate, self.fft_length + 1)


def as_f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def split_batch(x):
    """-> (flat2d, lead_shape, mem_kind, ptr, stream, alloc) for numpy or torch inputs."""
    if is_torch(x):
        import torch
        if not x.is_cuda:
            raise ValueError("torch inputs must live on a CUDA device; pass numpy arrays for host data")
        x = x.contiguous().float()
        lead = tuple(x.shape[:-1])
        x2 = x.reshape(-1, x.shape[-1])
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)

        def alloc(*shape):
            return torch.empty(shape, dtype=torch.float32, device=x.device)

        def ptr(t):
            return C.c_void_p(t.data_ptr())
        return x2, lead, MEM_DEVICE, ptr, stream, alloc
    x = as_f32(x)
    lead = tuple(x.shape[:-1])
    x2 = x.reshape(-1, x.shape[-1])

    def alloc(*shape):
        return np.empty(shape, dtype=np.float32)
    return x2, lead, MEM_HOST, np_ptr, C.c_void_p(None), alloc


def swap_last2(a):
    if is_torch(a):
        return a.transpose(-1, -2).contiguous()
    return np.ascontiguousarray(np.swapaxes(a, -1, -2))

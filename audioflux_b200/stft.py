"""STFT object (reference binding: python/audioflux/stft.py:14-300; C: src/stft_algorithm.c)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .base import Base, as_f32, np_ptr, split_batch
from .capi import opt_int, opt_float
from .lib import check
from .types import WindowType, PaddingPositionType, PaddingModeType, enum_value


class STFT(Base):
    def __init__(self, radix2_exp=12, window_type=WindowType.RECT, slide_length=1024, is_continue=False, _lib=None):
        super().__init__(_lib)
        self.radix2_exp = radix2_exp
        self.fft_length = 1 << radix2_exp
        self.window_type = window_type
        self.slide_length = slide_length
        status = self._lib.stftObj_new(C.byref(self._obj), radix2_exp, opt_int(enum_value(window_type)),
                                       opt_int(slide_length), opt_int(int(is_continue)))
        self.is_continue = is_continue
        if status != 0 or not self._obj:
            raise ValueError(f"stftObj_new failed with status {status}")
        self._is_created = True

    def set_slide_length(self, slide_length):
        self._lib.stftObj_setSlideLength(self._obj, slide_length)
        self.slide_length = slide_length

    def enable_continue(self, flag=False):
        """streaming mode (src/stft_algorithm.c:180-183, 474-599): successive stft() calls continue one signal"""
        self._lib.stftObj_enableContinue(self._obj, int(flag))
        self.is_continue = bool(flag)

    def enable_padding(self, flag=False):
        self._lib.stftObj_enablePadding(self._obj, int(flag))

    def set_padding(self, position_type=PaddingPositionType.CENTER, mode_type=PaddingModeType.CONSTANT,
                    value1=0.0, value2=0.0):
        self._lib.stftObj_setPadding(self._obj, opt_int(enum_value(position_type)),
                                     opt_int(enum_value(mode_type)), opt_float(value1), opt_float(value2))

    def use_window_data_arr(self, data_arr):
        w = as_f32(data_arr)
        if w.shape != (self.fft_length,):
            raise ValueError("window must have fft_length samples")
        self._lib.stftObj_useWindowDataArr(self._obj, np_ptr(w))

    def get_window_data_arr(self):
        p = self._lib.stftObj_getWindowDataArr(self._obj)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(self.fft_length,)).copy()

    def cal_time_length(self, data_length):
        return self._lib.stftObj_calTimeLength(self._obj, data_length)

    def cal_data_length(self, time_length):
        return self._lib.stftObj_calDataLength(self._obj, time_length)

    def stft_planes(self, data_arr):
        """Raw C layout: (re, im) each [T, fft_length] (full mirrored spectrum), one clip."""
        x = as_f32(data_arr)
        T = self.cal_time_length(x.shape[-1])
        re = np.zeros((T, self.fft_length), np.float32)
        im = np.zeros((T, self.fft_length), np.float32)
        self._lib.stftObj_stft(self._obj, np_ptr(x), x.shape[-1], np_ptr(re), np_ptr(im))
        return re, im

    def stft(self, data_arr):
        """-> complex [..., fft_length//2+1, T] like the reference wrapper (stft.py:259-300)."""
        x = as_f32(data_arr)
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        outs = []
        for i in range(x2.shape[0]):
            re, im = self.stft_planes(x2[i])
            outs.append((re + 1j * im).T[: self.fft_length // 2 + 1])
        out = np.stack(outs).reshape(*lead, self.fft_length // 2 + 1, -1)
        return np.ascontiguousarray(out)

    def stft_batch(self, data):
        """Additive batched entry point: data [B, L] (numpy host or torch cuda) ->
        (re, im) each [B, T, fft_length//2+1]."""
        fn = self._require_ext("stftObj_stftBatch")
        x2, lead, kind, ptr, stream, alloc = split_batch(data)
        B, L = x2.shape
        T = self.cal_time_length(L)
        re = alloc(B, T, self.fft_length // 2 + 1)
        im = alloc(B, T, self.fft_length // 2 + 1)
        check(fn(self._obj, ptr(x2), L, B, ptr(re), ptr(im), kind, stream), "stftObj_stftBatch")
        return re.reshape(*lead, T, -1), im.reshape(*lead, T, -1)

    def cal_data_length(self, time_length):
        return self._lib.stftObj_calDataLength(self._obj, int(time_length))

    def y_coords(self, samplate=32000):
        """Bin frequencies 0 .. samplate//2 with a leading 0 (plot axis, stft.py of the reference)."""
        return np.concatenate(([0.0], np.linspace(0, samplate // 2, self.fft_length // 2 + 1)))

    def x_coords(self, data_length, samplate=32000):
        if data_length < self.fft_length:
            raise ValueError(f"radix2_exp={self.radix2_exp}(fft_length={self.fft_length}) is too large for data_length={data_length}")
        return np.linspace(0, data_length / samplate, self.cal_time_length(data_length) + 1)

    def istft_planes(self, re, im, method_type=0):
        """Raw C layout: planes [T, fft_length] (full mirrored spectrum) -> data [(T-1)*slide + fft_length]."""
        re, im = as_f32(re), as_f32(im)
        out = np.zeros(self.cal_data_length(re.shape[0]), np.float32)
        self._lib.stftObj_istft(self._obj, np_ptr(re), np_ptr(im), re.shape[0], int(method_type), np_ptr(out))
        return out

    def istft(self, m_data_arr, method_type=0):
        """complex [..., fft_length//2+1, T] -> [..., data_length] like the reference wrapper (stft.py:302-361):
        method_type 0 'weight', 1 'overlap-add'."""
        z = np.asarray(m_data_arr)
        if not np.iscomplexobj(z):
            raise ValueError("m_data_arr must be of type np.complex")
        if z.ndim < 2:
            raise ValueError("m_data_arr's dimensions must be greater than 1")
        mirror = np.conj(z[..., ::-1, :][..., 1:-1, :])
        full = np.swapaxes(np.concatenate([z, mirror], axis=-2), -1, -2)          # [..., T, fft_length]
        lead = full.shape[:-2]
        f2 = full.reshape((-1,) + full.shape[-2:])
        outs = [self.istft_planes(f2[i].real, f2[i].imag, method_type) for i in range(f2.shape[0])]
        return np.stack(outs).reshape(*lead, -1)

    def istft_batch(self, re, im, method_type=0):
        """Additive: planes [B, T, W] with W = fft_length//2+1 (as stft_batch returns them) or fft_length
        (numpy host | torch cuda) -> data [B, (T-1)*slide + fft_length]."""
        fn = self._require_ext("stftObj_istftBatch")
        r2, lead, kind, ptr, stream, alloc = split_batch(re)
        i2 = split_batch(im)[0]
        if len(lead) < 1:
            raise ValueError("planes must be [..., T, W]")
        T, W = lead[-1], r2.shape[-1]
        B = r2.shape[0] // T
        out = alloc(B, self.cal_data_length(T))
        if kind == 0:
            out[...] = 0
        else:
            out.zero_()
        check(fn(self._obj, ptr(r2), ptr(i2), T, B, W, int(method_type), ptr(out), kind, stream), "stftObj_istftBatch")
        return out.reshape(*lead[:-1], -1)

    def __del__(self):
        if getattr(self, "_is_created", False):
            self._lib.stftObj_free(self._obj)
            self._is_created = False

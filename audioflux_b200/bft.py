"""BFT object: STFT -> power/magnitude -> mel/bark/erb/... filter bank
(reference binding: python/audioflux/bft.py:16-390; C: src/bft_algorithm.c)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .base import Base, BandAxis, FrameAxis, as_f32, np_ptr, split_batch, swap_last2
from .capi import opt_int, opt_float
from .lib import check
from .types import (WindowType, SpectralFilterBankScaleType, SpectralFilterBankStyleType,
                    SpectralFilterBankNormalType, SpectralDataType, CepstralRectifyType, enum_value)


class BFT(BandAxis, FrameAxis, Base):
    def __init__(self, num, radix2_exp=12, samplate=32000, low_fre=None, high_fre=None,
                 bin_per_octave=12, window_type=WindowType.HANN, slide_length=None,
                 scale_type=SpectralFilterBankScaleType.LINEAR,
                 style_type=SpectralFilterBankStyleType.SLANEY,
                 normal_type=SpectralFilterBankNormalType.NONE,
                 data_type=SpectralDataType.MAG, is_reassign=False, is_temporal=False, _lib=None):
        super().__init__(_lib)
        self.fft_length = fft_length = 1 << radix2_exp
        if num > fft_length // 2 + 1:
            raise ValueError(f"num={num} is too large")
        if low_fre is None:
            low_fre = 32.703196 if enum_value(scale_type) in (5, 6) else 0.0
        if high_fre is None:
            high_fre = samplate / 2
        if slide_length is None:
            slide_length = fft_length // 4
        self.num, self.radix2_exp, self.samplate = num, radix2_exp, samplate
        self.low_fre, self.high_fre, self.bin_per_octave = low_fre, high_fre, bin_per_octave
        self.window_type, self.slide_length = window_type, slide_length
        self.scale_type, self.style_type, self.normal_type = scale_type, style_type, normal_type
        self.data_type = data_type
        self.result_type = 0
        self.is_reassign, self.is_temporal = is_reassign, is_temporal
        status = self._lib.bftObj_new(
            C.byref(self._obj), num, radix2_exp, opt_int(samplate), opt_float(low_fre),
            opt_float(high_fre), opt_int(bin_per_octave), opt_int(enum_value(window_type)),
            opt_int(slide_length), opt_int(enum_value(scale_type)), opt_int(enum_value(style_type)),
            opt_int(enum_value(normal_type)), opt_int(enum_value(data_type)),
            opt_int(int(is_reassign)), opt_int(int(is_temporal)))
        if status != 0 or not self._obj:
            raise ValueError(f"bftObj_new failed with status {status}")
        self._is_created = True

    def cal_time_length(self, data_length):
        return self._lib.bftObj_calTimeLength(self._obj, data_length)

    def get_fre_band_arr(self):
        p = self._lib.bftObj_getFreBandArr(self._obj)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(self.num,)).copy()

    def get_bin_band_arr(self):
        p = self._lib.bftObj_getBinBandArr(self._obj)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int)), shape=(self.num,)).copy()

    def get_filter_bank_arr(self):
        """Additive: dense bank [num, fft_length//2+1] the device kernels consume."""
        fn = self._require_ext("bftObj_getFilterBankArr")
        out = np.zeros((self.num, self.fft_length // 2 + 1), np.float32)
        check(fn(self._obj, np_ptr(out)), "bftObj_getFilterBankArr")
        return out

    def set_result_type(self, result_type):
        self._lib.bftObj_setResultType(self._obj, int(result_type))
        self.result_type = int(result_type)

    def set_data_norm_value(self, norm_value):
        self._lib.bftObj_setDataNormValue(self._obj, C.c_float(norm_value))

    def bft_planes(self, data_arr, result_type=0):
        """Raw C layout: (re, im) each [T, num] for one clip."""
        x = as_f32(data_arr)
        if result_type != self.result_type:
            self.set_result_type(result_type)
        T = self.cal_time_length(x.shape[-1])
        re = np.zeros((T, self.num), np.float32)
        im = np.zeros((T, self.num), np.float32)
        self._lib.bftObj_bft(self._obj, np_ptr(x), x.shape[-1], np_ptr(re), np_ptr(im))
        return re, im

    def get_temporal_data(self, data_length):
        """(energy, rms, zero-crossing rate) of the frames of the LAST `bft` call, each [T] (bft.py:391-417,
        bftObj_getTemporalData); needs is_temporal=True."""
        if not self.is_temporal:
            raise ValueError("Please set the parameter is_temporal=True when creating the BFT object")
        T = self.cal_time_length(data_length)
        ptrs = [C.POINTER(C.c_float)() for _ in range(3)]
        self._lib.bftObj_getTemporalData(self._obj, *[C.byref(q) for q in ptrs])
        if not all(bool(q) for q in ptrs):
            raise ValueError("Please call the `BFT.bft()` method before calling this method")
        return tuple(np.ctypeslib.as_array(q, shape=(T,)).copy() for q in ptrs)

    def bft(self, data_arr, result_type=0):
        """-> [..., num, T] complex (result_type 0) or float32 (1), as bft.py:310-389."""
        x = as_f32(data_arr)
        if x.shape[-1] < self.fft_length:
            raise ValueError(f"radix2_exp={self.radix2_exp} is too large for data length {x.shape[-1]}")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        outs = []
        for i in range(x2.shape[0]):
            re, im = self.bft_planes(x2[i], result_type)
            outs.append(re if result_type else re + 1j * im)
        out = np.stack(outs).reshape(*lead, -1, self.num)
        return swap_last2(out)

    # ---- additive batched / device-pointer entry points (include/afb200_ext.h) ----
    def bft_batch(self, data, result_type=1):
        """data [B, L] (numpy host | torch cuda) -> [B, T, num] real, or (re, im) for result_type 0."""
        fn = self._require_ext("bftObj_bftBatch")
        if result_type != self.result_type:
            self.set_result_type(result_type)
        x2, lead, kind, ptr, stream, alloc = split_batch(data)
        B, L = x2.shape
        T = self.cal_time_length(L)
        re = alloc(B, T, self.num)
        im = alloc(B, T, self.num) if result_type == 0 else None
        check(fn(self._obj, ptr(x2), L, B, ptr(re), ptr(im) if im is not None else C.c_void_p(None),
                 kind, stream), "bftObj_bftBatch")
        re = re.reshape(*lead, T, self.num)
        return re if result_type else (re, im.reshape(*lead, T, self.num))

    def mfcc_batch(self, data, cc_num=13, rectify_type=CepstralRectifyType.LOG, out=None):
        """Fused STFT -> |.|^2 (or |.|) -> bank -> log10/cbrt -> DCT-II(ortho) -> first cc_num.
        data [B, L] -> [B, T, cc_num].  Equals bft(result_type=1) followed by XXCC.xxcc.
        Host arrays go through the library's chunked copy-in / transform / copy-out pipeline; pass page-locked arrays
        (and a page-locked `out`) to run it at PCIe speed."""
        fn = self._require_ext("bftObj_mfccBatch")
        x2, lead, kind, ptr, stream, alloc = split_batch(data)
        B, L = x2.shape
        T = self.cal_time_length(L)
        if out is None:
            out = alloc(B, T, cc_num)
        elif tuple(out.shape) != (B, T, cc_num) or not (out.flags["C_CONTIGUOUS"] if hasattr(out, "flags") else out.is_contiguous()):
            raise ValueError(f"out must be a contiguous float32 array of shape {(B, T, cc_num)}")
        check(fn(self._obj, ptr(x2), L, B, cc_num, enum_value(rectify_type), ptr(out), kind, stream),
              "bftObj_mfccBatch")
        return out.reshape(*lead, T, cc_num)

    def __del__(self):
        if getattr(self, "_is_created", False):
            self._lib.bftObj_free(self._obj)
            self._is_created = False

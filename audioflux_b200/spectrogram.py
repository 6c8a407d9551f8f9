"""Spectrogram front door: STFT -> power/magnitude -> Linear slice or mel/bark/erb/... bank, plus the
cepstral calls (reference binding: python/audioflux/spectrogram.py:31-503, 1771-2270;
C: src/spectrogram_algorithm.c).  Chroma / Deep bank types and the spectral descriptors are outside the path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .base import Base, BandAxis, as_f32, np_ptr, split_batch, swap_last2
from .capi import opt_int, opt_float
from .lib import check
from .types import (WindowType, SpectralFilterBankScaleType, SpectralFilterBankStyleType,
                    SpectralFilterBankNormalType, SpectralDataType, CepstralRectifyType, enum_value)

_LOG_LIKE = (5, 6)


class Spectrogram(BandAxis, Base):
    def __init__(self, num=0, samplate=32000, low_fre=None, high_fre=None, bin_per_octave=12, radix2_exp=12,
                 window_type=None, slide_length=None, data_type=SpectralDataType.POWER,
                 filter_bank_type=SpectralFilterBankScaleType.LINEAR,
                 style_type=SpectralFilterBankStyleType.SLANEY,
                 normal_type=SpectralFilterBankNormalType.NONE, is_continue=False, _lib=None):
        super().__init__(_lib)
        scale = enum_value(filter_bank_type)
        if scale == 5:
            if bin_per_octave not in (12, 24, 36):
                raise ValueError(f"bin_per_octave={bin_per_octave} must be 12, 24 or 36")
            if num % bin_per_octave != 0:
                raise ValueError(f"num={num} must be an integer multiple of bin_per_octave={bin_per_octave}")
        if low_fre is None:
            low_fre = 32.703196 if scale in _LOG_LIKE else 0.0
        if high_fre is None:
            high_fre = samplate / 2
        if window_type is None:
            window_type = WindowType.HANN
        if scale in _LOG_LIKE and low_fre < 32.703:
            raise ValueError(f"low_fre={low_fre} must be greater than or equal to 32.703")
        if low_fre < 0:
            raise ValueError(f"low_fre={low_fre} must be a non-negative number")
        self.fft_length = 1 << radix2_exp
        if slide_length is None:
            slide_length = self.fft_length // 4
        self.samplate, self.low_fre, self.high_fre = samplate, low_fre, high_fre
        self.bin_per_octave, self.radix2_exp, self.window_type = bin_per_octave, radix2_exp, window_type
        self.slide_length, self.is_continue, self.data_type = slide_length, is_continue, data_type
        self.filter_bank_type, self.style_type, self.normal_type = filter_bank_type, style_type, normal_type
        status = self._lib.spectrogramObj_new(
            C.byref(self._obj), int(num), opt_int(samplate), opt_float(low_fre), opt_float(high_fre),
            opt_int(bin_per_octave), opt_int(radix2_exp), opt_int(enum_value(window_type)), opt_int(slide_length),
            opt_int(int(is_continue)), opt_int(enum_value(data_type)), opt_int(scale),
            opt_int(enum_value(style_type)), opt_int(enum_value(normal_type)))
        if status != 0 or not self._obj:
            raise ValueError(f"spectrogramObj_new failed with status {status}")
        self._is_created = True
        self.num = self.get_band_num()

    def set_data_norm_value(self, norm_value):
        self._lib.spectrogramObj_setDataNormValue(self._obj, C.c_float(norm_value))

    def cal_time_length(self, data_length):
        return self._lib.spectrogramObj_calTimeLength(self._obj, data_length)

    def get_band_num(self):
        return self._lib.spectrogramObj_getBandNum(self._obj)

    def get_bin_band_length(self):
        return self._lib.spectrogramObj_getBinBandLength(self._obj)

    def get_fre_band_arr(self):
        p = self._lib.spectrogramObj_getFreBandArr(self._obj)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(self.get_bin_band_length(),)).copy()

    def get_bin_band_arr(self):
        p = self._lib.spectrogramObj_getBinBandArr(self._obj)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int)), shape=(self.get_bin_band_length(),)).copy()

    def spectrogram_planes(self, data_arr, is_phase_arr=False):
        """Raw C layout: one clip -> [T, num] (and phase [T, num], Linear bank only)."""
        x = as_f32(data_arr)
        T = self.cal_time_length(x.shape[-1])
        spec = np.zeros((T, self.num), np.float32)
        phase = np.zeros((T, self.num), np.float32) if is_phase_arr else None
        self._lib.spectrogramObj_spectrogram(self._obj, np_ptr(x), x.shape[-1], np_ptr(spec),
                                             np_ptr(phase) if is_phase_arr else None)
        return (spec, phase) if is_phase_arr else spec

    def spectrogram(self, data_arr, is_phase_arr=False):
        """data [..., L] -> [..., num, T] (and phase) as spectrogram.py:239-326."""
        if is_phase_arr and enum_value(self.filter_bank_type) != 0:
            raise ValueError("Only LINEAR bank type has phase arr")
        x = as_f32(data_arr)
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        res = [self.spectrogram_planes(x2[i], is_phase_arr) for i in range(x2.shape[0])]
        if is_phase_arr:
            spec = np.stack([r[0] for r in res]).reshape(*lead, -1, self.num)
            phase = np.stack([r[1] for r in res]).reshape(*lead, -1, self.num)
            return swap_last2(spec), swap_last2(phase)
        return swap_last2(np.stack(res).reshape(*lead, -1, self.num))

    def spectrogram_batch(self, data, is_phase_arr=False):
        """Additive: data [B, L] (numpy host | torch cuda) -> [B, T, num] (time-major; + phase for LINEAR)."""
        fn = self._require_ext("spectrogramObj_spectrogramBatch")
        x2, lead, kind, ptr, stream, alloc = split_batch(data)
        B, L = x2.shape
        T = self.cal_time_length(L)
        spec = alloc(B, T, self.num)
        phase = alloc(B, T, self.num) if is_phase_arr else None
        check(fn(self._obj, ptr(x2), L, B, ptr(spec), ptr(phase) if is_phase_arr else None, kind, stream),
              "spectrogramObj_spectrogramBatch")
        spec = spec.reshape(*lead, T, self.num)
        return (spec, phase.reshape(*lead, T, self.num)) if is_phase_arr else spec

    def mfcc_batch(self, data, cc_num=13, rectify_type=CepstralRectifyType.LOG):
        """Additive: the fused STFT -> bank -> log -> DCT kernel. data [B, L] -> [B, T, cc_num]."""
        fn = self._require_ext("spectrogramObj_mfccBatch")
        x2, lead, kind, ptr, stream, alloc = split_batch(data)
        B, L = x2.shape
        T = self.cal_time_length(L)
        out = alloc(B, T, cc_num)
        check(fn(self._obj, ptr(x2), L, B, cc_num, enum_value(rectify_type), ptr(out), kind, stream),
              "spectrogramObj_mfccBatch")
        return out.reshape(*lead, T, cc_num)

    def _cc(self, fn_name, m_data_arr, cc_num, rectify_type=None):
        """[num, T] of the LAST spectrogram call -> [cc_num, T]."""
        m = as_f32(m_data_arr)
        if m.ndim != 2:
            raise ValueError("cepstral calls work on the [num, T] result of the preceding spectrogram() call")
        if cc_num > self.num:
            raise ValueError("cc_num must be <= num")
        mt = np.ascontiguousarray(m.T)
        out = np.zeros((mt.shape[0], cc_num), np.float32)
        fn = getattr(self._lib, fn_name)
        if fn_name == "spectrogramObj_xxcc":
            fn(self._obj, np_ptr(mt), cc_num, opt_int(enum_value(rectify_type)), np_ptr(out))
        else:
            fn(self._obj, np_ptr(mt), cc_num, np_ptr(out))
        return np.ascontiguousarray(out.T)

    def xxcc(self, m_data_arr, cc_num=13, rectify_type=CepstralRectifyType.LOG):
        return self._cc("spectrogramObj_xxcc", m_data_arr, cc_num, rectify_type)

    def mfcc(self, m_data_arr, cc_num=13):
        if enum_value(self.filter_bank_type) != 2:
            raise ValueError("mfcc needs the MEL bank type")
        return self._cc("spectrogramObj_mfcc", m_data_arr, cc_num)

    def bfcc(self, m_data_arr, cc_num=13):
        if enum_value(self.filter_bank_type) != 3:
            raise ValueError("bfcc needs the BARK bank type")
        return self._cc("spectrogramObj_bfcc", m_data_arr, cc_num)

    def gtcc(self, m_data_arr, cc_num=13):
        if enum_value(self.style_type) != 2:
            raise ValueError("gtcc needs the GAMMATONE style")
        return self._cc("spectrogramObj_gtcc", m_data_arr, cc_num)

    def x_coords(self, data_length):
        """Frame start times in seconds: slide_length / samplate apart (plot axis of the reference's spectrogram classes)."""
        if data_length < self.fft_length:
            raise ValueError(f"radix2_exp={self.radix2_exp}(fft_length={self.fft_length}) is too large for data_length={data_length}")
        return np.arange(self.cal_time_length(data_length) + 1) * (self.slide_length / self.samplate)

    def deconv(self, m_data_arr):
        """[num, T] of the LAST spectrogram call -> (tone, pitch), each [num, T] (spectrogram.py:328-362 of the reference)."""
        m = as_f32(m_data_arr)
        if m.ndim != 2:
            raise ValueError("deconv works on the [num, T] result of the preceding spectrogram() call")
        mt = np.ascontiguousarray(m.T)
        tone, pitch = np.zeros_like(mt), np.zeros_like(mt)
        self._lib.spectrogramObj_deconv(self._obj, np_ptr(mt), np_ptr(tone), np_ptr(pitch))
        return np.ascontiguousarray(tone.T), np.ascontiguousarray(pitch.T)

    def deconv_batch(self, m_tn):
        """Additive: [..., T, num] (numpy host | torch cuda, time-major as spectrogram_batch returns it) -> (tone, pitch)."""
        fn = self._require_ext("spectrogramObj_deconvBatch")
        x2, lead, kind, ptr, stream, alloc = split_batch(m_tn)
        if x2.shape[-1] != self.num:
            raise ValueError(f"last dimension must be num={self.num}")
        tone, pitch = alloc(*x2.shape), alloc(*x2.shape)
        check(fn(self._obj, ptr(x2), x2.shape[0], ptr(tone), ptr(pitch), kind, stream), "spectrogramObj_deconvBatch")
        return tone.reshape(*lead, self.num), pitch.reshape(*lead, self.num)

    def __del__(self):
        if getattr(self, "_is_created", False):
            self._lib.spectrogramObj_free(self._obj)
            self._is_created = False


def _scaled(scale, default_num):
    class _S(Spectrogram):
        def __init__(self, num=0, samplate=32000, low_fre=None, high_fre=None, radix2_exp=12,
                     window_type=WindowType.HANN, slide_length=None,
                     style_type=SpectralFilterBankStyleType.SLANEY,
                     normal_type=SpectralFilterBankNormalType.NONE,
                     data_type=SpectralDataType.POWER, is_continue=False, _lib=None):
            super().__init__(num=num or default_num, samplate=samplate, low_fre=low_fre, high_fre=high_fre,
                             radix2_exp=radix2_exp, window_type=window_type, slide_length=slide_length,
                             data_type=data_type, filter_bank_type=scale, style_type=style_type,
                             normal_type=normal_type, is_continue=is_continue, _lib=_lib)
    return _S


MelSpectrogram = _scaled(SpectralFilterBankScaleType.MEL, 128)     # spectrogram.py:1948-2054
MelSpectrogram.__name__ = "MelSpectrogram"
BarkSpectrogram = _scaled(SpectralFilterBankScaleType.BARK, 128)   # spectrogram.py:2056-2162
BarkSpectrogram.__name__ = "BarkSpectrogram"
ErbSpectrogram = _scaled(SpectralFilterBankScaleType.ERB, 128)     # spectrogram.py:2164-2270
ErbSpectrogram.__name__ = "ErbSpectrogram"

"""XXCC cepstral coefficients: rectify (log10 | cube root) -> DCT-II ortho -> first cc_num
(reference binding: python/audioflux/feature/xxcc.py:14-136; C: src/feature/xxcc_algorithm.c)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .base import Base, as_f32, np_ptr, split_batch, swap_last2, is_torch
from .capi import opt_int
from .lib import check
from .types import CepstralRectifyType, enum_value


class XXCC(Base):
    def __init__(self, num, _lib=None):
        super().__init__(_lib)
        if num < 2:
            raise ValueError("num must be >= 2")
        self.num = num
        self.time_length = 0
        status = self._lib.xxccObj_new(C.byref(self._obj), num)
        if status != 0 or not self._obj:
            raise ValueError(f"xxccObj_new failed with status {status}")
        self._is_created = True

    def set_time_length(self, time_length):
        self._lib.xxccObj_setTimeLength(self._obj, int(time_length))
        self.time_length = int(time_length)

    def xxcc_planes(self, m_tn, cc_num=13, rectify_type=CepstralRectifyType.LOG):
        """Raw C layout: in [T, num] -> out [T, cc_num]."""
        m = as_f32(m_tn)
        if m.shape[0] != self.time_length:
            self.set_time_length(m.shape[0])
        out = np.zeros((m.shape[0], cc_num), np.float32)
        self._lib.xxccObj_xxcc(self._obj, np_ptr(m), cc_num, opt_int(enum_value(rectify_type)), np_ptr(out))
        return out

    def xxcc(self, m_data_arr, cc_num=13, rectify_type=CepstralRectifyType.LOG):
        """m_data_arr [..., num, T] -> [..., cc_num, T] as feature/xxcc.py:90-136."""
        m = np.asarray(m_data_arr)
        if np.iscomplexobj(m):
            m = np.abs(m)
        m = as_f32(m)
        if cc_num > self.num:
            raise ValueError("cc_num must be <= num")
        lead = m.shape[:-2]
        m2 = np.ascontiguousarray(np.swapaxes(m, -1, -2)).reshape(-1, m.shape[-1], m.shape[-2])
        outs = [self.xxcc_planes(m2[i], cc_num, rectify_type) for i in range(m2.shape[0])]
        out = np.stack(outs).reshape(*lead, m.shape[-1], cc_num)
        return swap_last2(out)

    def xxcc_batch(self, m_tn, cc_num=13, rectify_type=CepstralRectifyType.LOG):
        """Additive: m_tn [..., T, num] time-major (numpy host | torch cuda) -> [..., T, cc_num]."""
        fn = self._require_ext("xxccObj_xxccBatch")
        x2, lead, kind, ptr, stream, alloc = split_batch(m_tn)
        rows = x2.shape[0]
        out = alloc(rows, cc_num)
        check(fn(self._obj, ptr(x2), rows, cc_num, enum_value(rectify_type), ptr(out), kind, stream),
              "xxccObj_xxccBatch")
        return out.reshape(*lead, cc_num)

    def __del__(self):
        if getattr(self, "_is_created", False):
            self._lib.xxccObj_free(self._obj)
            self._is_created = False

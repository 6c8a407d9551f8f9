"""Synchrosqueezing (VERDICT r1 missing #1): wsstObj_wsst / synsqObj_synsq on the GPU against the numpy oracle and the
reference build.

Parity bar (stated statistically): the squeezing scatters every CWT cell into the row
round(log2f(|f_inst|) ...), an INTEGER outcome of float32 transcendental math -- a cell whose value sits within a few ulp
of a rounding boundary lands one row up or down depending on the libm / GPU rounding of log2f, atan2f and the divide.
Between the numpy oracle and the reference build itself 0.02 % (wsst) / 0.2 % (synsq) of the time columns differ.  The
test therefore demands (i) >= 98 % of the time columns identical to 1e-5 relative, (ii) a relative Frobenius error
<= 1e-2 of the whole matrix, (iii) the plain CWT planes returned beside it within the usual 1e-4."""
import numpy as np
import pytest

from conftest import rel_max
from oracle import af_oracle as O

pytestmark = pytest.mark.gpu


def _signal(n, sr, seed):
    t = np.arange(n) / sr
    rng = np.random.default_rng(seed)
    return (0.5 * np.sin(2 * np.pi * (300 + 2000 * t) * t) + 0.2 * np.sin(2 * np.pi * 2500 * t) + 0.01 * rng.standard_normal(n)).astype(np.float32)


def _agree(a, b):
    scale = np.abs(b).max()
    cols = (np.abs(a - b) > 1e-5 * scale).any(axis=0)
    return 1.0 - cols.mean(), float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("radix,is_pad,scale,wavelet", [(12, False, O.SCALE_OCTAVE, O.WAVE_MORLET), (12, True, O.SCALE_OCTAVE, O.WAVE_MORLET),
                                                          (13, False, O.SCALE_LOG, O.WAVE_MORSE), (12, False, O.SCALE_LINEAR, O.WAVE_MORLET),
                                                          (12, False, O.SCALE_MEL, O.WAVE_BUMP)])
def test_wsst_vs_oracle(cuda_device, radix, is_pad, scale, wavelet):
    import audioflux_b200 as af
    sr, num = 32000, 84
    x = _signal(1 << radix, sr, radix)
    w = af.WSST(num, radix, sr, wavelet_type=af.WaveletContinueType(wavelet), scale_type=af.SpectralFilterBankScaleType(scale),
                is_padding=is_pad)
    re, im, cr, ci = w.wsst_planes(x)
    o_re, o_im, w_re, w_im = O.wsst(x, num, radix, sr, wavelet=wavelet, scale=scale, is_pad=is_pad,
                                    low=w.low_fre, high=w.high_fre)
    assert rel_max(cr, w_re) < 1e-4 and rel_max(ci, w_im) < 1e-4
    for got, want in ((re, o_re), (im, o_im)):
        same, fro = _agree(got, want)
        assert same >= 0.98 and fro <= 1e-2, (same, fro)
    assert np.abs(re).max() > 0


def test_wsst_vs_reference_build(cuda_device, ref_lib):
    import audioflux_b200 as af
    sr, num, radix = 32000, 84, 12
    x = _signal(1 << radix, sr, 3)
    got = af.WSST(num, radix, sr, is_padding=False).wsst_planes(x)
    want = af.WSST(num, radix, sr, is_padding=False, _lib=ref_lib).wsst_planes(x)
    assert rel_max(got[2], want[2]) < 1e-4 and rel_max(got[3], want[3]) < 1e-4
    for g, w in zip(got[:2], want[:2]):
        same, fro = _agree(g, w)
        assert same >= 0.98 and fro <= 1e-2, (same, fro)


def test_wsst_accumulates_into_the_callers_planes_and_refuses_order_2(cuda_device):
    import audioflux_b200 as af
    from audioflux_b200.base import np_ptr
    sr, num, radix = 32000, 84, 12
    x = _signal(1 << radix, sr, 5)
    w = af.WSST(num, radix, sr, is_padding=False)
    re, im, _, _ = w.wsst_planes(x)
    a = np.full((num, 1 << radix), 2.0, np.float32)
    b = np.full((num, 1 << radix), -1.0, np.float32)
    w._lib.wsstObj_wsst(w._obj, np_ptr(x), np_ptr(a), np_ptr(b), None, None)
    # (the additions start from the caller's value, so the sums round differently from 0 + ... : compare to float32 accuracy)
    assert np.allclose(a, re + np.float32(2.0), rtol=0, atol=4e-6) and np.allclose(b, im - np.float32(1.0), rtol=0, atol=4e-6)
    assert np.abs(a - 2.0).max() > 1e-3
    w.set_order(2)
    assert "order" in af.lib.last_error()


@pytest.mark.parametrize("scale", [O.SCALE_OCTAVE, O.SCALE_LINEAR, O.SCALE_BARK])
def test_synsq_vs_oracle_and_reference(cuda_device, ref_lib, scale):
    import audioflux_b200 as af
    sr, num, radix = 32000, 84, 12
    x = _signal(1 << radix, sr, 7)
    w_re, w_im = O.cwt(x, num, radix, sr, wavelet=O.WAVE_MORLET, scale=scale, is_pad=False)
    _, fre = O.cwt_filterbank(num, 1 << radix, sr, O.WAVE_MORLET, scale, None, None, 12, None, None, 0)
    fre = np.ascontiguousarray(fre, np.float32)
    st = af.SpectralFilterBankScaleType(scale)
    got = af.Synsq(num, radix, sr).synsq_planes(fre, st, w_re, w_im)
    want_o = O.synsq(fre, w_re, w_im, sr, scale)
    want_r = af.Synsq(num, radix, sr, _lib=ref_lib).synsq_planes(fre, st, w_re, w_im)
    for g, wo, wr in zip(got, want_o, want_r):
        for want in (wo, wr):
            same, fro = _agree(g, want)
            assert same >= 0.98 and fro <= 2e-2, (same, fro)
    assert np.abs(got[0]).max() > 0

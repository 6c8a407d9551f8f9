"""cqtObj_cqhc / cqtObj_deconv (VERDICT r1 missing #4): the numpy restatement pinned to the reference build on the CPU."""
import numpy as np
import pytest

from conftest import rel_max
from oracle import af_oracle as O


def spectra(T, num, seed):
    rng = np.random.default_rng(seed)
    k = np.arange(num)
    base = np.exp(-((k[None, :] - rng.uniform(5, num - 5, (T, 1))) / 6.0) ** 2) + 0.3 * np.exp(-k[None, :] / 30.0)
    return (base * rng.uniform(0.5, 2.0, (T, 1)) + 0.01 * rng.random((T, num))).astype(np.float32)


@pytest.mark.parametrize("num,bpo,hc", [(84, 12, 20), (96, 24, 30), (36, 12, 8)])
def test_oracle_deconv_cqhc_vs_reference_build(ref_lib, num, bpo, hc):
    import audioflux_b200 as af
    x = (0.1 * np.random.default_rng(0).standard_normal(24000)).astype(np.float32)
    c = af.CQT(num, 32000, bin_per_octave=bpo, _lib=ref_lib)
    z = c.cqt(x)                                      # sets the object's timeLength
    m = np.ascontiguousarray(np.abs(z).T.astype(np.float32))
    tone, pitch = c.deconv_planes(m)
    o_tone, o_pitch = O.cq_deconv(m, bpo)
    assert rel_max(o_tone, tone) < 1e-5 and rel_max(o_pitch, pitch) < 1e-4
    got = c.cqhc_planes(m, hc)
    assert rel_max(O.cqhc(m, hc, bpo), got) < 1e-5


@pytest.mark.parametrize("num,scale,r", [(128, "MEL", 11), (40, "BARK", 10), (257, "LINEAR", 9)])
def test_oracle_deconv_vs_reference_spectrogram_deconv(ref_lib, num, scale, r):
    """spectrogramObj_deconv (spectrogram_algorithm.c:1545-1612) is the same per-frame transform as cqtObj_deconv"""
    import audioflux_b200 as af
    x = (0.1 * np.random.default_rng(2).standard_normal(12000)).astype(np.float32)
    s = af.Spectrogram(num, radix2_exp=r, samplate=16000, filter_bank_type=getattr(af.SpectralFilterBankScaleType, scale), _lib=ref_lib)
    spec = s.spectrogram(x)                           # [num, T]; sets the object's timeLength
    assert spec.shape[0] == num
    tone, pitch = s.deconv(spec)
    o_tone, o_pitch = O.cq_deconv(np.ascontiguousarray(spec.T))
    assert tone.shape == spec.shape
    assert rel_max(o_tone, tone.T) < 1e-5 and rel_max(o_pitch, pitch.T) < 1e-4

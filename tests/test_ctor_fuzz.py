"""Differential fuzz of the object constructors against the reference build (CPU only, no compute calls): random
parameter sets go through `bftObj_new`, `spectrogramObj_new`, `cqtObj_newWith`, `cwtObj_new`, `pwtObj_new` of both
libraries; status codes, band counts, frequency / bin band arrays and time lengths must agree.  Every call runs in a
forked child because the reference corrupts its heap on some legal parameter sets (those cases are skipped and counted).
Deliberate deviations, asserted as such:
  * cqtObj_newWith returns -1 where slideLength cannot be halved octaveNum - 1 times (the reference builds the object
    and then frames with hop 0 inside cqtObj_cqt);
  * pwtObj_new refuses (-2, message) the Gammatone style -- the reference writes the rows of that pseudo bank
    fftLength/2+1 apart into rows that are fftLength long, and with the Linear scale its `__auditory_linearFilterBank`
    (auditory_filterBank.c:339-365) also overruns the band array -- and banks whose edges lie beyond samplate/2 (Log /
    Linspace scales with highFre at Nyquist: the reference keeps those weights on negative-frequency bins);
  * cwtObj_new refuses (-2) a bump wavelet with beta > gamma (support on negative frequencies, same reason)."""
import ctypes as C
import os
import pickle

import numpy as np
import pytest

from audioflux_b200.capi import opt_int, opt_float


def _arr(p, n, t):
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(t)), shape=(n,)).copy() if n > 0 else np.zeros(0)


def _bands(lib, o, prefix, n):
    f, b = getattr(lib, prefix + "_getFreBandArr"), getattr(lib, prefix + "_getBinBandArr", None)
    f.restype = C.c_void_p
    fa = _arr(f(o), n, C.c_float)
    if b is None:
        return fa, np.zeros(0, np.int32)
    b.restype = C.c_void_p
    return fa, _arr(b(o), n, C.c_int)


def run_bft(lib, a):
    o = C.c_void_p()
    st = lib.bftObj_new(C.byref(o), a["num"], a["r"], opt_int(a["sr"]), opt_float(a["lo"]), opt_float(a["hi"]), opt_int(a["bpo"]),
                        opt_int(a["win"]), opt_int(a["hop"]), opt_int(a["scale"]), opt_int(a["style"]), opt_int(a["norm"]),
                        opt_int(a["dt"]), opt_int(0), opt_int(0))
    if st != 0:
        return (st,)
    return (st,) + _bands(lib, o, "bftObj", a["num"]) + ((lib.bftObj_calTimeLength(o, 20000),),)


def run_spec(lib, a):
    o = C.c_void_p()
    st = lib.spectrogramObj_new(C.byref(o), a["num"], opt_int(a["sr"]), opt_float(a["lo"]), opt_float(a["hi"]), opt_int(a["bpo"]),
                                opt_int(a["r"]), opt_int(a["win"]), opt_int(a["hop"]), opt_int(0), opt_int(a["dt"]),
                                opt_int(a["scale"]), opt_int(a["style"]), opt_int(a["norm"]))
    if st != 0:
        return (st,)
    n, m = lib.spectrogramObj_getBandNum(o), lib.spectrogramObj_getBinBandLength(o)
    return (st,) + _bands(lib, o, "spectrogramObj", m) + ((n, m, lib.spectrogramObj_calTimeLength(o, 20000)),)


def run_cqt(lib, a):
    o = C.c_void_p()
    st = lib.cqtObj_newWith(C.byref(o), a["num"], opt_int(a["sr"]), opt_float(a["lo"]), opt_int(a["bpo"]), opt_float(a["factor"]),
                            opt_float(a["beta"]), opt_float(a["thresh"]), opt_int(a["win"]), opt_int(a["hop"]), opt_int(0),
                            opt_int(a["norm"]), opt_int(a["scale"]))
    if st != 0:
        return (st,)
    f, b = _bands(lib, o, "cqtObj", a["num"])
    return (st, f, b, (lib.cqtObj_getFFTLength(o), lib.cqtObj_calTimeLength(o, 30000), lib.cqtObj_calTimeLength(o, 5)))


def run_cwt(lib, a):
    o = C.c_void_p()
    st = lib.cwtObj_new(C.byref(o), a["num"], a["r"], opt_int(a["sr"]), opt_float(a["lo"]), opt_float(a["hi"]), opt_int(a["bpo"]),
                        opt_int(a["wave"]), opt_int(a["scale"]), opt_float(a["gamma"]), opt_float(a["beta"]), opt_int(a["pad"]))
    if st != 0:
        return (st,)
    return (st,) + _bands(lib, o, "cwtObj", a["num"]) + ((),)


def run_pwt(lib, a):
    o = C.c_void_p()
    st = lib.pwtObj_new(C.byref(o), a["num"], a["r"], opt_int(a["sr"]), opt_float(a["lo"]), opt_float(a["hi"]), opt_int(a["bpo"]),
                        opt_int(a["scale"]), opt_int(a["style"]), opt_int(a["norm"]), opt_int(a["pad"]))
    if st != 0:
        return (st,)
    return (st,) + _bands(lib, o, "pwtObj", a["num"]) + ((),)


def in_child(fn):
    """fn() in a forked child with stdout / stderr silenced; ('crash',) when the child died before answering"""
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        os.close(r)
        try:
            import faulthandler
            faulthandler.disable()                             # a reference crash in the child is an expected outcome, not a report
            devnull = os.open(os.devnull, os.O_WRONLY)
            os.dup2(devnull, 1)
            os.dup2(devnull, 2)
            os.write(w, pickle.dumps(fn()))
        finally:
            os._exit(0)
    os.close(w)
    data = b""
    while True:
        c = os.read(r, 1 << 20)
        if not c:
            break
        data += c
    os.close(r)
    os.waitpid(pid, 0)
    return pickle.loads(data) if data else ("crash",)


SR = [8000, 16000, 22050, 32000, 44100, 48000]


def gen(kind, rng):
    sr = int(rng.choice(SR))
    r = int(rng.integers(6, 13))
    n = 1 << r
    scale = int(rng.integers(0, 7))
    lo = float(rng.choice([0.0, 20.0, 32.703196, 100.0, 55.0, -5.0]))
    hi = float(rng.choice([sr / 2, sr / 4, 4000.0, sr / 2 + 100, 10.0]))
    if kind in ("bft", "spec"):
        if kind == "bft" and scale in (5, 6) and lo < 32.703:
            lo = 32.703196
        return dict(num=int(rng.integers(2, min(n // 2 + 1, 200))), r=r, sr=sr, lo=lo, hi=hi, bpo=int(rng.choice([12, 24, 36, 7])),
                    win=int(rng.integers(0, 14)), hop=int(rng.choice([n // 4, n // 2, 100, 1])), scale=scale,
                    style=int(rng.integers(0, 6)), norm=int(rng.integers(0, 3)), dt=int(rng.integers(0, 2)))
    if kind == "cqt":
        return dict(num=int(rng.choice([12, 24, 36, 48, 60, 84, 96, 120, 13, 7])), sr=sr, lo=float(rng.choice([32.703196, 27.5, 55.0, 65.4, 20.0])),
                    bpo=int(rng.choice([12, 24, 36, 7])), factor=float(rng.choice([1.0, 0.5, 2.0])), beta=float(rng.choice([0.0, 0.0, 5.0, 20.0])),
                    thresh=float(rng.choice([0.01, 0.001, 0.1])), win=int(rng.integers(0, 14)), hop=int(rng.choice([0, 128, 256, 512, 100])),
                    norm=int(rng.integers(0, 3)), scale=int(rng.integers(0, 2)))
    r = int(rng.integers(6, 14))
    n = 1 << r
    lo = float(rng.choice([0.0, 20.0, 32.703196, 100.0]))
    if scale in (5, 6) and lo < 32.703:
        lo = 32.703196
    base = dict(num=int(rng.integers(2, min(n // 2 + 1, 120))), r=r, sr=sr, lo=lo, hi=float(rng.choice([sr / 2, sr / 4, 4000.0])),
                bpo=int(rng.choice([12, 24, 36])), scale=scale, pad=int(rng.integers(0, 2)))
    if kind == "cwt":
        base.update(wave=int(rng.integers(0, 8)), gamma=float(rng.choice([3.0, 6.0, 5.0, 0.0])), beta=float(rng.choice([20.0, 2.0, 0.6, 0.0])))
    else:
        base.update(style=int(rng.integers(0, 6)), norm=int(rng.integers(0, 3)))
    return base


RUN = {"bft": run_bft, "spec": run_spec, "cqt": run_cqt, "cwt": run_cwt, "pwt": run_pwt}


@pytest.mark.parametrize("kind,seed,cases", [("bft", 11, 120), ("spec", 12, 120), ("cqt", 13, 90), ("cwt", 14, 90), ("pwt", 15, 90)])
def test_constructors_agree_with_the_reference_build(ref_lib, product_lib, kind, seed, cases):
    rng = np.random.default_rng(seed)
    compared = crashed = 0
    for _ in range(cases):
        a = gen(kind, rng)
        want = in_child(lambda: RUN[kind](ref_lib, a))
        got = in_child(lambda: RUN[kind](product_lib, a))
        assert got[0] != "crash", (kind, a)
        if want[0] == "crash":
            crashed += 1
            continue
        if kind == "cqt" and want[0] == 0 and got[0] == -1:
            # the documented refusal: hop 0 at the lowest octave
            assert RUN[kind](product_lib, a)[0] == -1 and b"cannot be halved" in product_lib.afb200_lastError(), a
            continue
        if kind in ("cwt", "pwt") and want[0] == 0 and got[0] == -2:
            RUN[kind](product_lib, a)                          # (the child's error text does not reach this process)
            msg = product_lib.afb200_lastError()
            assert (b"Gammatone" in msg or b"above the Nyquist bin" in msg or b"negative frequencies" in msg or b"power-of-two" in msg), (a, msg)
            if kind == "pwt" and b"Gammatone" in msg:
                assert a["style"] == 2, (a, msg)
            if kind == "pwt" and b"above the Nyquist bin" in msg:       # Log / Linspace band edges at / beyond Nyquist
                assert a["scale"] in (1, 6), (a, msg)
            if kind == "cwt" and b"negative frequencies" in msg:
                assert a["wave"] == 2, (a, msg)
            continue
        assert want[0] == got[0], (kind, a, want[0], got[0])
        if want[0] != 0:
            continue
        compared += 1
        assert want[3] == got[3], (kind, a, want[3], got[3])
        assert np.allclose(want[1], got[1], rtol=1e-6, atol=1e-6), (kind, a)
        assert np.array_equal(want[2], got[2]), (kind, a, want[2][:6], got[2][:6])
    assert compared >= cases // 4, (kind, compared, crashed)


def test_documented_refusals_are_loud(product_lib):
    """the three parameter regions where the product returns -2 with a message instead of the reference's result"""
    import audioflux_b200 as af
    S, ST, W = af.SpectralFilterBankScaleType, af.SpectralFilterBankStyleType, af.WaveletContinueType
    for make, text in ((lambda: af.PWT(40, 10, 16000, style_type=ST.GAMMATONE), "Gammatone"),
                       (lambda: af.PWT(40, 10, 16000, low_fre=32.703196, high_fre=8000.0, scale_type=S.LOG), "above the Nyquist bin"),
                       (lambda: af.CWT(40, 10, 16000, wavelet_type=W.BUMP, gamma=4.0, beta=20.0), "negative frequencies")):
        with pytest.raises(ValueError, match="status -2"):
            make()
        assert text in af.lib.last_error()
    af.CWT(40, 10, 16000, wavelet_type=W.BUMP)                               # defaults (gamma 5, beta 0.6) are fine
    af.PWT(40, 10, 16000, low_fre=32.703196, high_fre=4000.0, scale_type=S.LOG)

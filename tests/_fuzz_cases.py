"""Random configurations of the hot-path objects and the computation each one stands for, shared by the reference-side
worker (tests/_fuzz_ref_worker.py, CPU, fork-isolated) and the GPU fuzz test (tests/test_gpu_fuzz.py).  `compute(kind,
a, lib)` drives the repo's mirror classes bound to `lib` (None = the product library) and returns a dict of arrays."""
import numpy as np


def _sig(seed, n, sr):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    f = rng.uniform(60.0, 0.4 * sr, 3)
    x = 0.2 * np.sin(2 * np.pi * f[0] * t) + 0.1 * np.sin(2 * np.pi * f[1] * t + 1.0) + 0.05 * np.sin(2 * np.pi * f[2] * t)
    return (x + 0.05 * rng.standard_normal(n)).astype(np.float32)


SR = [8000, 16000, 22050, 32000, 44100, 48000]


def gen(kind, rng):
    sr = int(rng.choice(SR))
    seed = int(rng.integers(1, 1 << 30))
    if kind in ("bft", "spec"):
        r = int(rng.integers(6, 12))
        n = 1 << r
        scale = int(rng.integers(0, 7))
        lo = float(rng.choice([0.0, 20.0, 32.703196, 100.0, 55.0]))
        if scale in (5, 6) and lo < 32.703:
            lo = 32.703196
        num = int(rng.integers(2, min(n // 2 + 1, 160)))
        return dict(num=num, r=r, sr=sr, lo=lo, hi=float(rng.choice([sr / 2, sr / 4, 4000.0])), bpo=int(rng.choice([12, 24, 36])),
                    win=int(rng.integers(0, 14)), hop=int(rng.choice([n // 4, n // 2, n // 8 + 1, 100])), scale=scale,
                    style=int(rng.integers(0, 6)), norm=int(rng.integers(0, 3)), dt=int(rng.integers(0, 2)),
                    L=int(rng.integers(n, 6 * n + 7)), seed=seed, cc=int(rng.integers(1, 41)), rect=int(rng.integers(0, 2)))
    if kind == "stft":
        r = int(rng.integers(2, 12))
        n = 1 << r
        return dict(r=r, sr=sr, win=int(rng.integers(0, 14)), hop=int(rng.choice([max(1, n // 4), max(1, n // 2), n, n // 3 + 1, 2 * n])),
                    pad=int(rng.integers(0, 2)), pos=int(rng.integers(0, 3)), mode=int(rng.integers(0, 3)),
                    v1=float(rng.choice([0.0, 0.5, -1.0])), v2=float(rng.choice([0.0, 2.0])), L=int(rng.integers(n, 8 * n + 5)),
                    seed=seed, method=int(rng.integers(0, 2)))
    if kind == "cqt":
        return dict(num=int(rng.choice([12, 24, 36, 48, 60, 72, 84, 96])), sr=sr, lo=float(rng.choice([32.703196, 27.5, 55.0, 65.4])),
                    bpo=int(rng.choice([12, 12, 24, 36])), factor=float(rng.choice([1.0, 1.0, 0.5, 2.0])),
                    beta=float(rng.choice([0.0, 0.0, 0.0, 5.0, 20.0])), thresh=float(rng.choice([0.01, 0.001, 0.05])),
                    win=int(rng.choice([1, 1, 2, 3, 0, 6])), hop=int(rng.choice([0, 0, 128, 256, 512])), norm=int(rng.integers(0, 3)),
                    is_scale=int(rng.integers(0, 2)), L=int(rng.integers(3000, 24000)), seed=seed)
    r = int(rng.integers(6, 13))
    n = 1 << r
    scale = int(rng.integers(0, 7))
    lo = float(rng.choice([0.0, 20.0, 32.703196, 100.0]))
    if scale in (5, 6) and lo < 32.703:
        lo = 32.703196
    a = dict(num=int(rng.integers(2, min(n // 2 + 1, 100))), r=r, sr=sr, lo=lo, hi=float(rng.choice([sr / 2, sr / 4, 4000.0])),
             bpo=int(rng.choice([12, 24, 36])), scale=scale, pad=int(rng.integers(0, 2)), L=n, seed=seed)
    if kind == "cwt":
        a.update(wave=int(rng.integers(0, 8)), gamma=None, beta=None, det=int(rng.integers(0, 2)))
        if rng.integers(0, 3) == 0:
            a.update(gamma=float(rng.choice([3.0, 6.0, 5.0, 4.0])), beta=float(rng.choice([20.0, 2.0, 0.6, 1.0])))
    else:
        a.update(style=int(rng.integers(0, 6)), norm=int(rng.integers(0, 3)))
    return a


def compute(kind, a, lib=None):
    """-> dict name -> ndarray; ValueError from a constructor is reported as {'error': text}"""
    import audioflux_b200 as af
    T = af.types if hasattr(af, "types") else None
    x = _sig(a["seed"], a["L"], a["sr"])
    kw = {} if lib is None else {"_lib": lib}
    S, ST, NT, DT, W = (af.SpectralFilterBankScaleType, af.SpectralFilterBankStyleType, af.SpectralFilterBankNormalType,
                        af.SpectralDataType, af.WindowType)
    try:
        if kind == "bft":
            b = af.BFT(a["num"], a["r"], a["sr"], a["lo"], a["hi"], a["bpo"], W(a["win"]), a["hop"], S(a["scale"]), ST(a["style"]),
                       NT(a["norm"]), DT(a["dt"]), **kw)
            mel = b.bft(x, result_type=1)
            z = b.bft(x, result_type=0)
            cc = min(a["cc"], a["num"])
            xx = af.XXCC(a["num"], **kw)
            out = {"mel": mel, "zre": np.ascontiguousarray(z.real), "zim": np.ascontiguousarray(z.imag),
                   "cc": xx.xxcc(mel, cc, af.CepstralRectifyType(a["rect"]))}
            return out
        if kind == "spec":
            s = af.Spectrogram(a["num"], a["sr"], a["lo"], a["hi"], a["bpo"], a["r"], W(a["win"]), a["hop"], DT(a["dt"]), S(a["scale"]),
                               ST(a["style"]), NT(a["norm"]), **kw)
            spec = s.spectrogram(x)
            out = {"spec": spec}
            cc = min(a["cc"], spec.shape[0])
            out["cc"] = s.xxcc(spec, cc, af.CepstralRectifyType(a["rect"]))
            return out
        if kind == "stft":
            s = af.STFT(a["r"], W(a["win"]), a["hop"], **kw)
            if a["pad"]:
                s.enable_padding(True)
                s.set_padding(af.PaddingPositionType(a["pos"]), af.PaddingModeType(a["mode"]), a["v1"], a["v2"])
            re, im = s.stft_planes(x)
            return {"re": re, "im": im}
        if kind == "cqt":
            c = af.CQT(a["num"], a["sr"], a["lo"], a["bpo"], a["factor"], a["beta"], a["thresh"], W(a["win"]),
                       a["hop"] if a["hop"] else None, NT(a["norm"]), bool(a["is_scale"]), **kw)
            z = c.cqt(x)
            return {"re": np.ascontiguousarray(z.real), "im": np.ascontiguousarray(z.imag)}
        if kind == "cwt":
            c = af.CWT(a["num"], a["r"], a["sr"], a["lo"], a["hi"], a["bpo"], af.WaveletContinueType(a["wave"]), S(a["scale"]),
                       a["gamma"], a["beta"], bool(a["pad"]), **kw)
            z = c.cwt(x)
            out = {"re": np.ascontiguousarray(z.real), "im": np.ascontiguousarray(z.imag)}
            if a["det"]:
                c.enable_det(True)
                dr, di = c.cwt_det_planes(x)
                out["dre"], out["dim"] = dr, di
            return out
        if kind == "pwt":
            p = af.PWT(a["num"], a["r"], a["sr"], a["lo"], a["hi"], a["bpo"], S(a["scale"]), ST(a["style"]), NT(a["norm"]),
                       bool(a["pad"]), **kw)
            z = p.pwt(x)
            return {"re": np.ascontiguousarray(z.real), "im": np.ascontiguousarray(z.imag)}
    except ValueError as e:
        return {"error": str(e)}
    raise KeyError(kind)

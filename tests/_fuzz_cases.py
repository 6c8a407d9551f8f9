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
    if kind == "mfcc":
        # the fused kernels' configuration space: fftLength 2048, every bank family, odd clip lengths and hops
        scale = int(rng.choice([2, 2, 2, 3, 4, 1, 5, 6]))
        lo = float(rng.choice([0.0, 20.0, 100.0])) if scale not in (5, 6) else 32.703196
        num = int(rng.choice([128, 128, 80, 64, 40, 26, 13, 5]))
        if scale == 5:
            num = int(rng.choice([84, 60, 36]))
        hop = int(rng.choice([512, 512, 256, 1024, 160, 400, 441, 2048, 3000, 64]))
        return dict(num=num, r=11, sr=sr, lo=lo, hi=float(rng.choice([sr / 2, sr / 4, 0.45 * sr])), bpo=12, win=int(rng.choice([1, 1, 2, 0, 3, 8])),
                    hop=hop, scale=scale, style=int(rng.choice([0, 0, 0, 1, 4, 5])), norm=int(rng.integers(0, 3)), dt=int(rng.integers(0, 2)),
                    L=int(rng.choice([2048, 2049, 2560, 6000, 8192, 12345, 20000, 48000])), seed=seed, cc=int(rng.choice([13, 20, 40, 5, 64, 1])),
                    rect=int(rng.integers(0, 2)), clips=int(rng.integers(1, 4)))
    if kind == "istft":
        r = int(rng.integers(3, 12))
        n = 1 << r
        hop = int(rng.choice([max(1, n // 4), max(1, n // 2), max(1, n // 8), n // 3 + 1]))
        return dict(r=r, sr=sr, win=int(rng.integers(0, 14)), hop=hop, T=int(rng.integers(1, 12)), method=int(rng.integers(0, 2)),
                    half=int(rng.integers(0, 2)), seed=seed, L=0)
    if kind == "xxccstd":
        num = int(rng.integers(8, 129))
        return dict(num=num, T=int(rng.integers(1, 60)), cc=int(rng.integers(2, min(num, 40) + 1)), win=int(rng.choice([3, 5, 7, 9, 11, 2])),
                    etype=int(rng.integers(0, 3)), rect=int(rng.integers(0, 2)), seed=seed, sr=sr, L=0)
    if kind == "cqtpost":
        num = int(rng.choice([12, 24, 36, 48, 84, 96]))
        bpo = int(rng.choice([12, 12, 24]))
        if num % bpo:
            bpo = 12
        return dict(num=num, bpo=bpo, sr=sr, L=int(rng.integers(3000, 12000)), seed=seed, chroma=int(rng.choice([12, 12, 24])),
                    dt=int(rng.integers(0, 2)), cnorm=int(rng.integers(0, 5)), cc=int(rng.integers(1, min(num, 20) + 1)),
                    rect=int(rng.integers(0, 2)), hc=int(rng.integers(1, 30)))
    if kind == "stream":
        r = int(rng.integers(5, 11))
        n = 1 << r
        which = int(rng.integers(0, 3))                       # 0 STFT, 1 Spectrogram (mel), 2 CQT
        hop = int(rng.choice([n // 4, n // 2, n, n // 3 + 1, 2 * n])) if which != 2 else 0
        chunks = [int(c) for c in rng.integers(1, 4 * n, size=int(rng.integers(2, 7)))]
        if which == 2:
            chunks = [int(c) for c in rng.integers(1200, 6000, size=int(rng.integers(2, 5)))]
        return dict(which=which, r=r, sr=sr if which != 2 else 32000, hop=hop, chunks=chunks, seed=seed, L=sum(chunks))
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
    x = _sig(a["seed"], a["L"], a["sr"]) if a["L"] else None
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
        if kind == "mfcc":
            xs = np.stack([_sig(a["seed"] + i, a["L"], a["sr"]) for i in range(a["clips"])])
            b = af.BFT(a["num"], a["r"], a["sr"], a["lo"], a["hi"], a["bpo"], W(a["win"]), a["hop"], S(a["scale"]), ST(a["style"]),
                       NT(a["norm"]), DT(a["dt"]), **kw)
            cc = min(a["cc"], a["num"])
            rect = af.CepstralRectifyType(a["rect"])
            if lib is None:                                    # product: the batched entry points (fused kernels where they apply)
                return {"mel": np.swapaxes(b.bft_batch(xs, result_type=1), -1, -2), "cc": np.swapaxes(b.mfcc_batch(xs, cc, rect), -1, -2)}
            mel = b.bft(xs, result_type=1)                     # reference: clip by clip, BFT then XXCC
            return {"mel": mel, "cc": af.XXCC(a["num"], **kw).xxcc(mel, cc, rect)}
        if kind == "istft":
            n = 1 << a["r"]
            rng = np.random.default_rng(a["seed"])
            s = af.STFT(a["r"], W(a["win"]), a["hop"], **kw)
            xx = (0.1 * rng.standard_normal((a["T"] - 1) * a["hop"] + n)).astype(np.float32)
            w = s.get_window_data_arr()
            fr = np.stack([xx[t * a["hop"]:t * a["hop"] + n] * w for t in range(a["T"])])
            Z = np.fft.fft(fr, axis=1)                         # the same planes for both libraries, whatever their STFT does
            re, im = np.ascontiguousarray(Z.real, dtype=np.float32), np.ascontiguousarray(Z.imag, dtype=np.float32)
            y = s.istft_planes(re, im, a["method"])
            norm = np.zeros(len(y))
            for t in range(a["T"]):
                norm[t * a["hop"]:t * a["hop"] + n] += np.asarray(w, np.float64) ** (2 if a["method"] == 0 else 1)
            return {"y": np.where(norm > 1e-2, y, 0.0).astype(np.float32)}       # ill-conditioned samples masked
        if kind == "xxccstd":
            rng = np.random.default_rng(a["seed"])
            m = (rng.random((a["T"], a["num"])) ** 2 + 1e-3).astype(np.float32)
            e = (rng.random(a["T"]) + 0.1).astype(np.float32)
            xo = af.XXCC(a["num"], **kw)
            c0, c1, c2 = xo.xxcc_standard_planes(m, e, a["cc"], a["win"], af.CepstralEnergyType(a["etype"]), af.CepstralRectifyType(a["rect"]))
            return {"coe": c0, "d1": c1, "d2": c2}
        if kind == "cqtpost":
            c = af.CQT(a["num"], a["sr"], bin_per_octave=a["bpo"], **kw)
            T = c.cqt(x).shape[-1]                             # (the legacy post-processing calls use the object's last timeLength)
            rng = np.random.default_rng(a["seed"])
            z = ((rng.standard_normal((a["num"], T)) + 1j * rng.standard_normal((a["num"], T))) *
                 np.exp(-np.arange(a["num"])[:, None] / 30.0)).astype(np.complex64)      # the SAME spectrum for both libraries
            tone, pitch = c.deconv(z)
            return {"chroma": c.chroma(z, a["chroma"], DT(a["dt"]), af.ChromaDataNormalType(a["cnorm"])),
                    "cqcc": c.cqcc(z, a["cc"], af.CepstralRectifyType(a["rect"])), "cqhc": c.cqhc(z, a["hc"]),
                    "tone": tone, "pitch": pitch}
        if kind == "stream":
            n = 1 << a["r"]
            pos, outs = 0, []
            if a["which"] == 0:
                o = af.STFT(a["r"], W.HANN, a["hop"], is_continue=True, **kw)
                step = lambda piece: o.stft_planes(piece)[0]
            elif a["which"] == 1:
                o = af.Spectrogram(24, a["sr"], radix2_exp=a["r"], slide_length=a["hop"], filter_bank_type=S.MEL, is_continue=True, **kw)
                step = lambda piece: o.spectrogram_planes(piece)
            else:
                o = af.CQT(48, a["sr"], is_continue=True, **kw)
                step = lambda piece: np.abs(o.cqt(piece)).T.astype(np.float32)
            for cn in a["chunks"]:
                res = step(x[pos:pos + cn])
                pos += cn
                res = np.asarray(res, np.float32)
                outs.append(res.reshape(res.shape[0], int(np.prod(res.shape[1:]))))
            cols = max(o_.shape[1] for o_ in outs)
            return {"frames": np.concatenate([o_ for o_ in outs if o_.shape[0]] or [np.zeros((0, cols), np.float32)]),
                    "counts": np.array([o_.shape[0] for o_ in outs], np.float32)}
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

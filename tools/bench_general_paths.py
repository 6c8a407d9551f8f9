#!/usr/bin/env python3
"""Device-timed throughput of the GENERAL (non-fused) paths: MFCC at other FFT lengths, mel spectrogram, STFT, ISTFT.
usage: python tools/bench_general_paths.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import torch
import audioflux_b200 as af

S, D, W = af.SpectralFilterBankScaleType, af.SpectralDataType, af.WindowType

def timed(fn, steps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

out = {}
B, L = 256, 240000
x = 0.1 * torch.randn((B, L), device="cuda")
for r, hop in ((9, 128), (10, 256), (11, 512), (12, 1024)):
    b = af.BFT(128 if r > 9 else 64, r, 48000, slide_length=hop, scale_type=S.MEL, data_type=D.POWER)
    T = b.cal_time_length(L)
    ms = timed(lambda: b.mfcc_batch(x, 20))
    out[f"mfcc_n{1 << r}_hop{hop}"] = {"ms": ms, "frames_per_s": B * T / ms * 1e3, "fused": int(af.lib.get_lib().bftObj_mfccPlanMode(b._obj) >= 0)}
    ms = timed(lambda: b.bft_batch(x))
    out[f"mel_n{1 << r}_hop{hop}"] = {"ms": ms, "frames_per_s": B * T / ms * 1e3}
s = af.STFT(11, W.HANN, 512)
T = s.cal_time_length(L)
ms = timed(lambda: s.stft_batch(x))
out["stft_n2048_hop512"] = {"ms": ms, "frames_per_s": B * T / ms * 1e3}
re, im = s.stft_batch(x)
ms = timed(lambda: s.istft_batch(re, im))
out["istft_n2048_hop512"] = {"ms": ms, "frames_per_s": B * T / ms * 1e3}
print(json.dumps(out))

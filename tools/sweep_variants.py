#!/usr/bin/env python3
"""Time prebuilt library variants (variants/lib_*.so, linked on the build host from one differing object each) on config 2
in ONE process: every variant is loaded as a separate ctypes library (RTLD_LOCAL), checked bit-for-bit against the
default library's result, then timed with CUDA events.  usage: sweep_variants.py [reps] [name ...]"""
import glob, os, sys
import torch
sys.path.insert(0, '.')
import audioflux_b200 as af
from audioflux_b200.lib import load_library

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
names = sys.argv[2:] or sorted(os.path.basename(p)[4:-3] for p in glob.glob('variants/lib_*.so'))
S, D = af.SpectralFilterBankScaleType, af.SpectralDataType
x = 0.1 * torch.randn((1024, 240000), device='cuda')
ref = af.BFT(128, 11, 48000, slide_length=512, scale_type=S.MEL, data_type=D.POWER).mfcc_batch(x, 40).clone()
res = {}
for rnd in range(2):                                           # two rounds: the order of the variants must not matter
    for name in names:
        lib, _ = load_library(os.path.abspath(f'variants/lib_{name}.so'))
        b = af.BFT(128, 11, 48000, slide_length=512, scale_type=S.MEL, data_type=D.POWER, _lib=lib)
        out = b.mfcc_batch(x, 40)
        same = bool(torch.equal(out, ref))
        for _ in range(3): b.mfcc_batch(x, 40, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): b.mfcc_batch(x, 40, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res.setdefault(name, []).append(ms)
        print(f'round {rnd} variant {name:10s} {ms:8.4f} ms  bit-identical to default: {same}', flush=True)
print({k: round(min(v), 4) for k, v in res.items()})

#!/usr/bin/env python3
"""Device-timed throughput of the CQT (BASELINE config 3) and CWT (config 4) paths on one GPU,
with the reference CPU time per clip next to it.  Not the driver's bench (that is bench.py / MFCC).
usage: python tools/bench_cqt_cwt.py [--cqt-batch 1024] [--cwt-batch 32] [--cpu]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import torch
import audioflux_b200 as af

def timed(fn, steps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cqt-batch", type=int, default=1024)
    ap.add_argument("--cwt-batch", type=int, default=32)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    out = {}
    g = torch.Generator(device="cuda").manual_seed(1237)
    # ---- C3: CQT 84 bins, 48 kHz, 5 s clips
    B, L = a.cqt_batch, 240000
    x = 0.1 * torch.randn((B, L), generator=g, device="cuda")
    c = af.CQT(84, 48000)
    T = c.cal_time_length(L)
    ms = timed(lambda: c.cqt_batch(x))
    bytes_clip = 4 * L + 2 * 4 * T * 84
    out["cqt"] = {"batch": B, "ms": ms, "clips_per_s": B / ms * 1e3, "cqt_frames_per_s": B * T / ms * 1e3,
                  "compulsory_GBs": B * bytes_clip / ms / 1e6, "frac_hbm": B * bytes_clip / ms / 1e6 / peaks["hbm_gbs"],
                  "fp32_tflops": B * 7 * T * 512 * 24 * 2 / ms / 1e9}
    del x
    # ---- C4: CWT morlet 84 scales, N = 2^19 (10 s zero-padded), isPad=0
    B2, N = a.cwt_batch, 1 << 19
    xw = torch.zeros((B2, N), device="cuda")
    xw[:, :480000] = 0.1 * torch.randn((B2, 480000), generator=g, device="cuda")
    w = af.CWT(84, 19, 48000, wavelet_type=af.WaveletContinueType.MORLET, is_padding=False)
    re = torch.empty((B2, 84, N), device="cuda"); im = torch.empty_like(re)
    from audioflux_b200.lib import check
    import ctypes as C
    fn = w._lib.cwtObj_cwtBatch
    def run_cwt():
        check(fn(w._obj, C.c_void_p(xw.data_ptr()), B2, C.c_void_p(re.data_ptr()), C.c_void_p(im.data_ptr()), 1,
                 C.c_void_p(torch.cuda.current_stream().cuda_stream)), "cwtObj_cwtBatch")
    ms2 = timed(run_cwt, steps=3, warm=1)
    bytes_clip2 = 4 * 480000 + 2 * 4 * 84 * N
    out["cwt"] = {"batch": B2, "ms": ms2, "clips_per_s": B2 / ms2 * 1e3, "compulsory_GBs": B2 * bytes_clip2 / ms2 / 1e6,
                  "frac_hbm": B2 * bytes_clip2 / ms2 / 1e6 / peaks["hbm_gbs"], "fp32_tflops": B2 * 85 * 5 * N * 19 / ms2 / 1e9}
    if a.cpu:
        from oracle.ref_lib import get_ref_lib
        ref = get_ref_lib()
        xc = (0.1 * np.random.default_rng(0).standard_normal(L)).astype(np.float32)
        rc = af.CQT(84, 48000, _lib=ref); rc.cqt_planes(xc)
        t0 = time.perf_counter(); rc.cqt_planes(xc); out["cqt"]["ref_cpu_s_per_clip_1core"] = time.perf_counter() - t0
        xw1 = (0.1 * np.random.default_rng(0).standard_normal(1 << 16)).astype(np.float32)
        rw = af.CWT(84, 16, 48000, wavelet_type=af.WaveletContinueType.MORLET, is_padding=False, _lib=ref)
        t0 = time.perf_counter(); rw.cwt_planes(xw1); out["cwt"]["ref_cpu_s_per_clip_1core_N65536"] = time.perf_counter() - t0
    print(json.dumps(out))

if __name__ == "__main__":
    main()

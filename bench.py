#!/usr/bin/env python3
"""bench.py -- MFCC frames/s (48 kHz, n_fft=2048, hop=512, 128 mel, 40 coefficients) on N B200s.

Contract (see the task statement / DESIGN.md section 6):
  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the reference's own CPU path, oracle/_ref)
One JSON line on stdout from rank 0.

A "step" is one pass of the fused STFT->mel->log->DCT path over one synthetic batch:
  N=1 : BASELINE config 2 = 1024 clips x 5 s (983 MB of samples, larger than the 126 MB L2).
  N>1 : the same 1024 clips PER GPU (weak scaling, config 5 at N=8) with the all-gather of the (1024, 465, 40)
        result blocks -- the path's only exchange step -- fused into the kernel epilogue as NVLink P2P stores
        (--gather peer, default) or done by NCCL on a side stream (--gather nccl).
`value` is device-timed (CUDA events on the launching stream) with inputs resident in HBM;
`e2e` goes through the public call with pinned HOST buffers, H2D and D2H inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.realpath(__file__))
_REAL_STDOUT = None


def emit(obj):
    """The one JSON line of the contract, on the process's original stdout."""
    text = json.dumps(obj) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text.encode())

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR, RADIX, NFFT, HOP, NMEL, NCC = 48000, 11, 2048, 512, 128, 40
CLIP_SECONDS = 5
L = SR * CLIP_SECONDS
T = (L - NFFT) // HOP + 1
BATCH_PER_GPU = 1024
# algorithmic (compulsory) HBM bytes per clip: every sample read once + result written once
BYTES_PER_CLIP = 4 * L + 4 * T * NCC
BYTES_PER_FRAME_READ_MODEL = 4 * NFFT + 4 * NCC     # north-star "each frame reads its window" accounting
FLOP_PER_FRAME = 75e3


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons polled every ~5 ms through NVML (nvidia-smi every 200 ms as a fallback) in a side
    thread; `mark()` brackets the timed region and the reported median / reasons come from the samples inside it."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.nvml = index, [], None, None
        self.t0 = self.t1 = None
        self.stop_flag = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self, begin):
        if begin:
            self.t0 = time.perf_counter()
        else:
            self.t1 = time.perf_counter()

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.perf_counter(), mhz, self.max_mhz, mask))
            except Exception:
                pass
            time.sleep(0.004)

    def _read(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                mask = sum(bit for (name, bit), v in zip((("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40),
                                                           ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)), r[5:9])
                           if v.lower().startswith("active"))
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]), mask))
            except Exception:
                continue

    def stop(self):
        if self.nvml is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
        else:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        inside = [r for r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        use = inside if inside else self.rows
        reasons = sorted(name for name, bit in self.BITS.items() if any(r[3] & bit for r in use))
        return {"sm_mhz": float(np.median([r[1] for r in use])) if use else None,
                "sm_max_mhz": max(r[2] for r in use) if use else None,
                "samples": len(use), "samples_in_timed_region": len(inside), "samples_total": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi", "reasons": reasons}


# --------------------------------------------------------------------------- reference / CPU arm
def _cpu_worker(args):
    lib_path, seed, clips, warm = args
    import ctypes
    sys.path.insert(0, ROOT)
    import audioflux_b200 as af
    from audioflux_b200 import capi
    lib = ctypes.CDLL(lib_path)
    capi.bind(lib)
    S, D = af.SpectralFilterBankScaleType, af.SpectralDataType
    b = af.BFT(NMEL, RADIX, SR, slide_length=HOP, scale_type=S.MEL, data_type=D.POWER, _lib=lib)
    xx = af.XXCC(NMEL, _lib=lib)
    x = (0.1 * np.random.default_rng(seed).standard_normal((clips, L))).astype(np.float32)
    for i in range(warm):
        mel, _ = b.bft_planes(x[i % clips], 1)
        xx.xxcc_planes(mel, NCC)
    t0 = time.perf_counter()
    for i in range(clips):
        mel, _ = b.bft_planes(x[i], 1)
        xx.xxcc_planes(mel, NCC)
    return time.perf_counter() - t0


def cpu_reference_rate(clips_per_worker, workers, warm=1):
    """The reference's own C path (oracle/_ref: built-in radix-2 FFT + naive double-accumulated dot; no
    FFTW/MKL/BLAS exist in this image) on `workers` independent processes, one object each, disjoint
    clips -- the fair all-cores figure, since the reference's OpenMP only splits frames inside a clip."""
    from oracle import ref_lib as R
    import multiprocessing as mp
    if not R.available():
        return None
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(R.REF_PATH, 1234 + i, clips_per_worker, warm) for i in range(workers)])
        wall = time.perf_counter() - t0
        # wall includes process start + warm-up; time a second, steady pass for the rate
        t0 = time.perf_counter()
        times = pool.map(_cpu_worker, [(R.REF_PATH, 4321 + i, clips_per_worker, 0) for i in range(workers)])
        wall2 = time.perf_counter() - t0
    busy = max(times)
    return {"frames_per_s": workers * clips_per_worker * T / busy, "busy_s": busy, "wall_s": wall2, "first_wall_s": wall}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per = max(2, int(os.environ.get("AFB200_CPU_CLIPS_PER_WORKER", "4")))
    vals = []
    t_all = time.perf_counter()
    for _ in range(max(1, args.warmup // 3)):
        cpu_reference_rate(1, cores, warm=0)
    for _ in range(max(1, min(args.steps, 3))):
        r = cpu_reference_rate(per, cores)
        if r is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref/libaudioflux_ref.so not built"})
            return 0
        vals.append(r["frames_per_s"])
    v = float(np.median(vals))
    frames_step = per * cores * T
    line = {
        "metric": "mfcc_frames_per_s", "value": v, "unit": "frames/s", "impl": "reference", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * frames_step / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the B200 arm's workload (same clips, same transform); each reference step is a bounded sample of it
        "config": {"workload": f"BASELINE config {'2' if args.gpus == 1 else '5-shaped'}: batch={args.batch} x 5 s 48 kHz clips per GPU, "
                               "STFT(2048,hop 512,hann)->mel128(slaney)->log10->DCT MFCC(40)",
                   "batch_per_gpu": args.batch, "clip_samples": L, "frames_per_clip": T,
                   "implementation": "reference CPU path: bftObj_bft(resultType=1) + xxccObj_xxcc of oracle/_ref (unmodified reference sources)",
                   "sample": f"{per} clips x {cores} worker processes per step ({per * cores} of the {args.batch} clips)"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "reference",
                         "sample": f"{per * cores} clips ({frames_step} frames) per step, {cores} processes x 1 object; "
                                   "reference built with gcc -O3, built-in radix-2 FFT + naive dot (no FFTW/MKL/BLAS in the image)"},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------- B200 arm
def run_b200_arm(args):
    import torch
    import audioflux_b200 as af
    from audioflux_b200 import lib as L_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line, but NCCL prints "NCCL version ..." there when the box sets NCCL_DEBUG:
        # from here on file descriptor 1 is stderr, and the JSON line is written to the saved real stdout (emit())
        global _REAL_STDOUT
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = L_.get_lib()

    B = args.batch
    S, D = af.SpectralFilterBankScaleType, af.SpectralDataType
    bft = af.BFT(NMEL, RADIX, SR, slide_length=HOP, scale_type=S.MEL, data_type=D.POWER)
    g = torch.Generator(device=dev).manual_seed(1234 + 2 + rank)
    x = 0.1 * torch.randn((B, L), generator=g, device=dev, dtype=torch.float32)
    # N > 1: the gather of the per-rank results.  Default "peer": ONE kernel per step, whose epilogue stores every
    # finished tile into all ranks' gathered arrays over NVLink (audioflux_b200/dist.py:PeerScatter), then a 4-byte
    # all-reduce as the cross-rank fence.  Fallback / comparison "nccl": chunked compute with chunk k's
    # all_gather_into_tensor on a side stream while chunk k+1 computes (OverlappedGather).
    overlap = scatter = None
    gather_mode = "none"
    if world > 1:
        from audioflux_b200.dist import OverlappedGather, PeerScatter
        gather_mode = args.gather
        if gather_mode == "peer":
            try:
                scatter = PeerScatter(bft, B, L, NCC)
            except RuntimeError as e:
                if rank == 0:
                    print(f"[bench] peer scatter unavailable, using NCCL all-gather: {e}", file=sys.stderr)
                gather_mode = "nccl"
        if gather_mode == "nccl":
            overlap = OverlappedGather(chunks=args.gather_chunks)

    def step(inp=None):
        inp = x if inp is None else inp
        if scatter is not None:
            out = scatter(inp)
            scatter.fence()
            return out
        if overlap is not None:
            return overlap(lambda c: bft.mfcc_batch(c, NCC), inp)
        return bft.mfcc_batch(inp, NCC)

    # parity gate on this very configuration before any timing counts (clip 0 vs the numpy oracle)
    parity = None
    if rank == 0:
        from oracle import af_oracle as O
        out = bft.mfcc_batch(x, NCC)          # no collective here: only rank 0 runs the gate
        torch.cuda.synchronize()
        want = O.mfcc(x[0].cpu().numpy(), SR, RADIX, HOP, NMEL, NCC)
        parity = float(np.abs(out[0].cpu().numpy() - want).max() / np.abs(want).max())
        if not parity < 1e-4:
            emit({"error": f"parity gate failed: rel err {parity}"})
            return 1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    # gather gate (N > 1): rank 0 regenerates the LAST rank's seeded shard, transforms it locally and compares with
    # what arrived in its gathered array -- bit for bit (results do not depend on which GPU computed them)
    gather_ok = None
    if world > 1 and rank == 0:
        gl = torch.Generator(device=dev).manual_seed(1234 + 2 + (world - 1))
        xl = 0.1 * torch.randn((B, L), generator=gl, device=dev, dtype=torch.float32)
        want_l = bft.mfcc_batch(xl[:8], NCC)
        res = step()
        torch.cuda.synchronize()
        got_l = res[world - 1][:8] if scatter is not None else res[0][(world - 1) * res[0].shape[0] // world:][:8]
        gather_ok = bool(torch.equal(got_l, want_l))
        del xl
    elif world > 1:
        step()
        torch.cuda.synchronize()
    if dist:
        dist.barrier()
    if gather_ok is False:
        emit({"error": "gather gate failed: the last rank's block did not arrive intact on rank 0"})
        return 1

    launches0 = lib.afb200_kernelLaunchCount()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kern_ms = []
    torch.cuda.synchronize()
    sampler.mark(True)
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    sampler.mark(False)
    if dist:
        dist.barrier()
    launches = lib.afb200_kernelLaunchCount() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    if world > 1:                      # kernel-only duration for the roofline: one untimed-by-the-metric extra pass
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(3):
            bft.mfcc_batch(x, NCC)
        k1.record()
        torch.cuda.synchronize()
        kernel_ms = k0.elapsed_time(k1) / 3
    else:
        kernel_ms = float(np.mean(per_step))
    t = torch.tensor([total_ms], device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_step = total_ms / args.steps
    value = world * B * T / (ms_step * 1e-3)

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    xh = torch.empty((B, L), dtype=torch.float32).pin_memory()
    xh.copy_(x.cpu())
    oh = torch.empty((B, T, NCC), dtype=torch.float32).pin_memory()
    xd = torch.empty_like(x)

    xh_np, oh_np = xh.numpy(), oh.numpy()           # views of the page-locked buffers for the C-ABI host-pointer call
    h2d_stream = torch.cuda.Stream()
    h2d_done = [torch.cuda.Event() for _ in range(8)]

    def e2e_step():
        if world == 1:
            # the call a user of the C ABI makes: bftObj_mfccBatch with HOST pointers (memKind 0); inside, the library
            # pipelines copy-in / transform / copy-out over chunks and returns when `oh` is complete
            bft.mfcc_batch(xh_np, NCC, out=oh_np)
            return
        if scatter is not None:
            # stream the shard in chunk by chunk: copy-in of chunk k+1 on a side stream overlaps the fused transform +
            # scatter of chunk k; then the cross-rank fence and this rank's own slot of the gathered array -> host
            cur = torch.cuda.current_stream()
            nchunk = 8
            per = (B + nchunk - 1) // nchunk
            for k in range(nchunk):
                lo, hi = k * per, min(B, (k + 1) * per)
                if lo >= hi:
                    break
                with torch.cuda.stream(h2d_stream):
                    if k == 0:
                        h2d_stream.wait_stream(cur)                  # xd is free again (previous step's kernels done)
                    xd[lo:hi].copy_(xh[lo:hi], non_blocking=True)
                    h2d_done[k].record(h2d_stream)
                cur.wait_event(h2d_done[k])
                scatter(xd[lo:hi], clip_offset=lo)
            scatter.fence()
            oh.copy_(scatter.gathered[rank], non_blocking=True)
            return
        xd.copy_(xh, non_blocking=True)            # --gather nccl: copy-in, chunked compute + overlapped all-gather
        res = step(xd)
        per = res[0].shape[0] // world
        for k, o in enumerate(res):                # this rank's own rows of every gathered chunk -> host
            oh[k * per:(k + 1) * per].copy_(o[rank * per:(rank + 1) * per], non_blocking=True)

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(n_e2e):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    te = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * T / (float(te.item()) / n_e2e * 1e-3)

    if scatter is not None:
        torch.cuda.synchronize()
        dist.barrier()                 # nobody unmaps while a peer may still be storing
        scatter.close()
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0

    peaks, peak_kind = measured_peaks()
    achieved = B * BYTES_PER_CLIP / (kernel_ms * 1e-3) / 1e9
    cores = os.cpu_count() or 1
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_rate(max(2, int(os.environ.get("AFB200_CPU_CLIPS_PER_WORKER", "4"))), cores)
    line = {
        "metric": "mfcc_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE config {'2' if world == 1 else '5-shaped'}: batch={B} x 5 s 48 kHz clips per GPU, "
                               "STFT(2048,hop 512,hann)->mel128(slaney)->log10->DCT MFCC(40), fused kernel",
                   "batch_per_gpu": B, "clip_samples": L, "frames_per_clip": T,
                   "l2": "inputs (983 MB per GPU) exceed the 126 MB L2; no explicit flush needed",
                   "collective": {"none": "none",
                                  "peer": "fused: the kernel epilogue stores each tile into every rank's gathered (world,B,T,40) array "
                                          "over NVLink P2P (cudaIpc-mapped), then a 4-byte NCCL all-reduce as the cross-rank fence",
                                  "nccl": f"nccl all_gather of (B,T,40) per step in {args.gather_chunks} chunks overlapped with compute"}[gather_mode],
                   "gather_gate_bitexact": gather_ok,
                   "parity_rel_err_clip0": parity},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": B * L * 4, "d2h_bytes_per_step": B * T * NCC * 4,
                "steps": n_e2e},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"],
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel at this configuration from the
                     # committed ncu --set full capture (profiles/r1_final_mfcc_fused_ncu_summary.txt): 981.77 + 72.74 MB
                     "traffic": (981766144 + 72740864) if (B == 1024 and L == 240000) else None,
                     "traffic_source": "profiles/r1_final_mfcc_fused_ncu_summary.txt", "peak_kind": peak_kind,
                     "kernel": "k_mfcc_fused<5>", "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": B * BYTES_PER_CLIP,
                     "frame_read_model_frac": (B * T * BYTES_PER_FRAME_READ_MODEL / (kernel_ms * 1e-3) / 1e9) / peaks["hbm_gbs"],
                     "fp32_tflops": B * T * FLOP_PER_FRAME / (kernel_ms * 1e-3) / 1e12},
        "per_step_ms": per_step,
    }
    if cpu:
        line["cpu_baseline"] = {"value": cpu["frames_per_s"], "unit": "frames/s", "cores": cores, "kind": "reference",
                                "sample": f"{cores} processes x {max(2, int(os.environ.get('AFB200_CPU_CLIPS_PER_WORKER', '4')))} clips, "
                                          "oracle/_ref (gcc -O3, built-in radix-2 FFT + naive dot)"}
    emit(line)
    if dist:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="clips per GPU (default: BASELINE config 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather-chunks", type=int, default=4, help="N>1, --gather nccl: chunks per step for compute/all-gather overlap")
    ap.add_argument("--gather", choices=("peer", "nccl"), default="peer",
                    help="N>1: 'peer' = all-gather fused into the kernel epilogue (NVLink P2P stores); 'nccl' = overlapped NCCL all-gather")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())

#!/bin/bash
# builds the fused kernel with different numbers of frame warps on the GPU box and benches each
mkdir -p gpurun_out
for W in "$@"; do
  touch audioflux_b200/csrc/kernels/mfcc_fused.cu
  make -s -C audioflux_b200/csrc EXTRA_NVFLAGS=-DAF_FRAME_WARPS=$W > /dev/null 2>&1
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/sweep_$W.json 2> gpurun_out/sweep_$W.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/sweep_$W.json')); print('W=$W', round(d['value']/1e6,1), 'Mframes/s', round(d['ms_per_step'],3), 'ms parity', d['config']['parity_rel_err_clip0'])
except Exception as e: print('W=$W failed', e)"
done

#!/bin/bash
# builds the fused kernel on the GPU box with different (frame warps, epilogue warps, CTAs per SM) and benches each
# usage: tools/sweep_frame_warps.sh 12,2,1 5,1,2 ...
mkdir -p gpurun_out
for cfg in "$@"; do
  IFS=, read W E C <<< "$cfg"
  touch audioflux_b200/csrc/kernels/mfcc_fused.cu
  make -s -C audioflux_b200/csrc EXTRA_NVFLAGS="-DAF_FRAME_WARPS=$W -DAF_EPI_WARPS=$E -DAF_CTAS_PER_SM=$C" > /dev/null 2>&1
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/sweep_$cfg.json 2> gpurun_out/sweep_$cfg.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/sweep_$cfg.json')); print('cfg=$cfg', round(d['value']/1e6,1), 'Mframes/s', round(d['ms_per_step'],3), 'ms parity', d['config']['parity_rel_err_clip0'])
except Exception as e: print('cfg=$cfg failed', e, open('gpurun_out/sweep_$cfg.err').read()[-300:])"
done

#!/bin/bash
# builds the v2 fused kernel on the GPU box with different (frame warps, helper warps, ablation mask) and times config 2
# usage: tools/sweep_mfcc2.sh 13,3,3,0 12,2,1,0 13,3,3,1 ...   (ablated builds compute wrong results on purpose)
mkdir -p gpurun_out
for cfg in "$@"; do
  IFS=, read W E D A <<< "$cfg"
  touch audioflux_b200/csrc/kernels/mfcc_fused2.cu
  make -s -C audioflux_b200/csrc EXTRA_NVFLAGS="-DAF2_FRAME_WARPS=$W -DAF2_BANK_WARPS=$E -DAF2_DCT_WARPS=$D -DAF2_ABLATE=$A" > /dev/null 2>&1
  timeout 120 python - <<PY
import torch, sys
sys.path.insert(0,'.')
import audioflux_b200 as af
S,D=af.SpectralFilterBankScaleType, af.SpectralDataType
b=af.BFT(128,11,48000,slide_length=512,scale_type=S.MEL,data_type=D.POWER)
x=0.1*torch.randn((1024,240000),device='cuda')
for _ in range(3): b.mfcc_batch(x,40)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8): b.mfcc_batch(x,40)
e1.record(); torch.cuda.synchronize()
print('cfg=$cfg (frame, bank, dct warps, ablate)', round(e0.elapsed_time(e1)/8,4),'ms', flush=True)
PY
done
touch audioflux_b200/csrc/kernels/mfcc_fused2.cu; make -s -C audioflux_b200/csrc > /dev/null 2>&1

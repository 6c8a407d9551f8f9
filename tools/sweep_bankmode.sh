#!/bin/bash
# diagnostic: fused MFCC kernel time per bank-loop mode (AFB200_MFCC_BANK_MODE) and ablation mask
# usage: tools/sweep_bankmode.sh "mode:ablate" ...   e.g. 0:0 1:0 1:32 1:64 1:96 1:1
mkdir -p gpurun_out
for cfg in "$@"; do
  mode=${cfg%%:*}; a=${cfg##*:}
  touch audioflux_b200/csrc/kernels/mfcc_fused.cu
  make -s -C audioflux_b200/csrc EXTRA_NVFLAGS="-DAF_ABLATE=$a" > /dev/null 2>&1
  AFB200_MFCC_BANK_MODE=$mode python - <<PY
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
for _ in range(10): b.mfcc_batch(x,40)
e1.record(); torch.cuda.synchronize()
print('mode=$mode ablate=$a plan', af.lib.get_lib().bftObj_mfccPlanMode(b._obj), round(e0.elapsed_time(e1)/10,4),'ms')
PY
done
touch audioflux_b200/csrc/kernels/mfcc_fused.cu; make -s -C audioflux_b200/csrc > /dev/null 2>&1

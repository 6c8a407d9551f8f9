set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/debug_mfcc2.py > gpurun_out/dbg1.log 2>&1; tail -4 gpurun_out/dbg1.log
DBG_B=4 DBG_L=240000 timeout 300 python tools/debug_mfcc2.py > gpurun_out/dbg2.log 2>&1; tail -4 gpurun_out/dbg2.log
bash tools/sweep_mfcc2.sh 13,2,0 13,2,1 12,3,0 14,1,0 > gpurun_out/sweep3.log 2>&1; cat gpurun_out/sweep3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mfcc_fused2 -c 1 -f -o gpurun_out/r2_mfcc2c python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/debug_cqt.py > gpurun_out/dbgcqt.log 2>&1; tail -30 gpurun_out/dbgcqt.log
bash tools/sweep_mfcc2.sh 13,3,3,2 13,3,3,3 13,3,3,31 > gpurun_out/sweep5.log 2>&1; cat gpurun_out/sweep5.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mfcc_fused2 -c 1 -f -o gpurun_out/r2_mfcc2d python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log

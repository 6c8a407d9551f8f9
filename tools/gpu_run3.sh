set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/debug_mfcc2.py > gpurun_out/dbg1.log 2>&1; tail -4 gpurun_out/dbg1.log
bash tools/sweep_mfcc2.sh 13,2,0 13,2,1 12,3,0 12,3,1 13,2,5 13,2,13 13,2,29 > gpurun_out/sweep2.log 2>&1; cat gpurun_out/sweep2.log
timeout 300 python tools/bench_cqt_cwt.py --cwt-batch 8 > gpurun_out/cqt_tc.json 2> gpurun_out/cqt_tc.err; cat gpurun_out/cqt_tc.json; tail -2 gpurun_out/cqt_tc.err
AFB200_CQT_KERNEL=fp32 timeout 300 python tools/bench_cqt_cwt.py --cwt-batch 8 > gpurun_out/cqt_fp32.json 2> gpurun_out/cqt_fp32.err; cat gpurun_out/cqt_fp32.json | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mfcc_fused2 -c 1 -f -o gpurun_out/r2_mfcc2b python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log

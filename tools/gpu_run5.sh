set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/debug_mfcc2.py > gpurun_out/dbg1.log 2>&1; tail -4 gpurun_out/dbg1.log
timeout 120 python tools/debug_cqt.py > gpurun_out/dbgcqt.log 2>&1; tail -20 gpurun_out/dbgcqt.log
bash tools/sweep_mfcc2.sh 13,3,3,0 13,3,3,1 > gpurun_out/sweep4.log 2>&1; cat gpurun_out/sweep4.log
timeout 200 python tools/bench_cqt_cwt.py --cwt-batch 2 > gpurun_out/cqt_umma.json 2> gpurun_out/cqt_umma.err; cut -c1-330 gpurun_out/cqt_umma.json; tail -2 gpurun_out/cqt_umma.err
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
bash tools/sweep_mfcc2.sh 13,5,1,0 13,4,2,0 14,4,1,0 > gpurun_out/sweep6.log 2>&1; cat gpurun_out/sweep6.log
timeout 200 python tools/bench_cqt_cwt.py --cwt-batch 2 > gpurun_out/cqt_umma.json 2> gpurun_out/cqt_umma.err; cut -c1-330 gpurun_out/cqt_umma.json; tail -2 gpurun_out/cqt_umma.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cqt_octave_umma -c 7 -f -o gpurun_out/r2_cqt_umma python tools/bench_cqt_cwt.py --cwt-batch 1 > gpurun_out/ncu_cqt.log 2>&1; tail -2 gpurun_out/ncu_cqt.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log

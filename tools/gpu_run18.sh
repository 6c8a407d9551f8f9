set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2d_pytest_gpu.log
bash tools/sweep_cwt.sh 4
AFB200_CWT_L2PERSIST=0 bash tools/sweep_cwt.sh 4
AFB200_CWT_NOPRUNE=1 bash tools/sweep_cwt.sh 4
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_cwt_fused -c 2 python tools/cwt_prof.py 2>&1 | grep -E "k_cwt|duration|dram__" | head -12
AFB200_CWT_L2PERSIST=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_cwt_fused -c 1 python tools/cwt_prof.py 2>&1 | grep -E "k_cwt|duration|dram__" | head -6
bash tools/sweep_mfcc2.sh 13,4,2,0 13,3,2,0 13,3,3,0 13,2,2,0 12,2,1,0 13,4,2,0 2>&1 | grep cfg=
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; cut -c1-330 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python tools/debug_mfcc2.py > gpurun_out/dbg1.log 2>&1; tail -30 gpurun_out/dbg1.log
DBG_B=5 DBG_L=240000 timeout 300 python tools/debug_mfcc2.py > gpurun_out/dbg2.log 2>&1; tail -12 gpurun_out/dbg2.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; cat gpurun_out/bench_v2.json | cut -c1-600; tail -3 gpurun_out/bench_v2.err
AFB200_MFCC_KERNEL=v1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; cat gpurun_out/bench_v1.json | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log

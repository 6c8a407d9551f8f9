set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/debug_mfcc2.py > gpurun_out/dbg1.log 2>&1; tail -4 gpurun_out/dbg1.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; cat gpurun_out/bench_v2.json | cut -c1-400; tail -3 gpurun_out/bench_v2.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mfcc_fused2 -c 1 -f -o gpurun_out/r2_mfcc2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu.log 2>&1; tail -3 gpurun_out/ncu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; cat gpurun_out/bench_full.json | cut -c1-3000; tail -3 gpurun_out/bench_full.err

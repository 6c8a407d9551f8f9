set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/debug_cqt.py > gpurun_out/dbgcqt.log 2>&1; grep -E "tcgen05 bo=0|fp32 vs" gpurun_out/dbgcqt.log
timeout 120 python tools/debug_mfcc2.py > gpurun_out/dbg1.log 2>&1; tail -3 gpurun_out/dbg1.log
timeout 200 python tools/bench_cqt_cwt.py --cwt-batch 2 > gpurun_out/cqt_umma2.json 2> gpurun_out/cqt_umma2.err; cut -c1-330 gpurun_out/cqt_umma2.json; tail -2 gpurun_out/cqt_umma2.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_v2b.json 2> gpurun_out/bench_v2b.err; cut -c1-300 gpurun_out/bench_v2b.json; tail -2 gpurun_out/bench_v2b.err
AFB200_MFCC_KERNEL=v1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_v1b.json 2> gpurun_out/bench_v1b.err; cut -c1-300 gpurun_out/bench_v1b.json

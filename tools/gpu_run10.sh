set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','roofline','cpu_baseline','clocks')})
print('cqt', {k:v for k,v in d['cqt'].items() if k not in ('per_step_ms','config')})
print('cwt', {k:v for k,v in d['cwt'].items() if k not in ('per_step_ms','config')})
"; tail -3 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-1500 gpurun_out/bench_ref.json

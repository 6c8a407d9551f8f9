set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest_gpu.log 2>&1; tail -25 gpurun_out/r2h_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2h_bench_quick.json 2> gpurun_out/r2h_bench_quick.err; tail -3 gpurun_out/r2h_bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench_quick.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['cqt']['ms_per_step'], d['cqt']['e2e']); print(d['cwt']['ms_per_step'], d['cwt']['e2e'])
PY

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2_final_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; cut -c1-300 gpurun_out/r2_final_bench_n1.json; tail -3 gpurun_out/r2_final_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_final_bench_n1.json'))
print('MFCC', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['frame_read_model_frac'], d['roofline']['traffic'])
print('CQT', d['cqt']['value'], d['cqt']['ms_per_step'], d['cqt']['e2e']['value'])
print('CWT', d['cwt']['value'], d['cwt']['ms_per_step'], d['cwt']['e2e']['value'], d['cwt']['roofline'])
print('CPU', d['cpu_baseline']['value'], d['cqt']['cpu_baseline']['value'], d['cwt']['cpu_baseline']['value'], d['clocks'])
PY

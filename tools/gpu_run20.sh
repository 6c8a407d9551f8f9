set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -x -q -k "multi or istft_16384 or temporal or cwt or pwt or wsst" > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2f_pytest_gpu.log
bash tools/sweep_cwt.sh 4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; cut -c1-400 gpurun_out/r2f_bench_n2.json; tail -3 gpurun_out/r2f_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config'].get('gather_gate_bitexact'), d['e2e'], d['roofline']['kernel_ms'])
PY

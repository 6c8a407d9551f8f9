set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err; cut -c1-300 gpurun_out/r2g_bench_n8.json; tail -3 gpurun_out/r2g_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench_n8.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config'].get('gather_gate_bitexact'), d['e2e'], d['roofline']['kernel_ms'], d['config'].get('host_binding'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2g_bench_n4.json 2> gpurun_out/r2g_bench_n4.err; cut -c1-200 gpurun_out/r2g_bench_n4.json

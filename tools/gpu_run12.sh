set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/debug_cqt.py > gpurun_out/dbgcqt.log 2>&1; grep -E "tcgen05|DONE|Error|error" gpurun_out/dbgcqt.log | head -12
timeout 200 python tools/bench_cqt_cwt.py --cwt-batch 8 > gpurun_out/cqt_umma5.json 2> gpurun_out/cqt_umma5.err; cut -c1-330 gpurun_out/cqt_umma5.json; tail -2 gpurun_out/cqt_umma5.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_cqt_octave|k_decimate" -c 13 --csv --log-file gpurun_out/launches_cqt3.csv python tools/bench_cqt_cwt.py --cqt-batch 1024 --cwt-batch 1 > /dev/null 2>&1; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_cqt3.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]: print(r[ki][:45], r[vi])
PY
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log

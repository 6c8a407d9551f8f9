set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
# 1. full GPU parity suite
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu.log
# 2. the driver's bench line
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cut -c1-600 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_bench_n1.err
# 3. launch list of the same command
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
d=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki][:50]; d.setdefault(k,[]).append(float(r[vi].replace(',','')))
for k,v in d.items(): print(f"{k:52s} n={len(v):4d} total={sum(v)/1e6:10.3f} ms  avg={sum(v)/len(v)/1e3:10.1f} us")
PY
# 4. full captures of the three dominant kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mfcc_fused2 -s 2 -c 1 -f -o gpurun_out/r2_mfcc2 python tools/mfcc_prof.py > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2_mfcc2.ncu-rep > gpurun_out/r2_mfcc_fused2_ncu_summary.txt; cat gpurun_out/r2_mfcc_fused2_ncu_summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_cqt_octave|k_decimate" -s 13 -c 13 -f -o gpurun_out/r2_cqt python tools/cqt_prof.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cwt -s 4 -c 2 -f -o gpurun_out/r2_cwt python tools/cwt_prof.py > /dev/null 2>&1
ncu -i gpurun_out/r2_cqt.ncu-rep --page raw --csv > gpurun_out/r2_cqt_raw.csv
ncu -i gpurun_out/r2_cwt.ncu-rep --page raw --csv > gpurun_out/r2_cwt_raw.csv
ls -la gpurun_out

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/debug_cqt.py > gpurun_out/dbgcqt.log 2>&1; grep -E "tcgen05 bo=0" gpurun_out/dbgcqt.log | head -7
timeout 200 python tools/bench_cqt_cwt.py --cwt-batch 8 > gpurun_out/cqt_umma3.json 2> gpurun_out/cqt_umma3.err; cut -c1-700 gpurun_out/cqt_umma3.json; tail -2 gpurun_out/cqt_umma3.err
AFB200_CWT_FUSED=0 timeout 200 python tools/bench_cqt_cwt.py --cqt-batch 8 --cwt-batch 8 > gpurun_out/cwt_unfused.json 2> gpurun_out/cwt_unfused.err; cut -c330-700 gpurun_out/cwt_unfused.json
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_cqt_octave|k_cwt|k_decimate" -c 40 --csv --log-file gpurun_out/launches_cqt_cwt.csv python tools/bench_cqt_cwt.py --cqt-batch 1024 --cwt-batch 8 > /dev/null 2>&1; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_cqt_cwt.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[ii], r[ki][:40]), {})[r[mi]]=r[vi]
for k,v in list(d.items())[:40]: print(k, v)
PY
timeout 900 python -m pytest tests -m gpu -x -q -k "cwt or pwt or wsst or synsq or squeeze or cqt" > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
bash tools/sweep_mfcc2.sh 12,2,1,0 11,3,1,0 > gpurun_out/sweep7.log 2>&1; cat gpurun_out/sweep7.log

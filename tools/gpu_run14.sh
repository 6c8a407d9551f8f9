set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# --- CQT: FFMA2 decimator + two-issuer tcgen05 kernel ---
timeout 120 python tools/debug_cqt.py > gpurun_out/dbgcqt.log 2>&1; grep -E "tcgen05 bo=0|DONE|Error|error" gpurun_out/dbgcqt.log | head -12
timeout 200 python tools/bench_cqt_cwt.py --cwt-batch 8 > gpurun_out/cqt_umma6.json 2> gpurun_out/cqt_umma6.err; cut -c1-330 gpurun_out/cqt_umma6.json; tail -2 gpurun_out/cqt_umma6.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_cqt_octave|k_decimate" -c 13 --csv --log-file gpurun_out/launches_cqt4.csv python tools/bench_cqt_cwt.py --cqt-batch 1024 --cwt-batch 1 > /dev/null 2>&1; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_cqt4.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]: print(r[ki][:45], r[vi])
PY
timeout 900 python -m pytest tests -m gpu -x -q -k "cqt or chroma or cqcc or pitch or fullbatch" > gpurun_out/pytest_cqt.log 2>&1; tail -4 gpurun_out/pytest_cqt.log
# --- MFCC A/B: new (split bank items) vs the previous kernel source, same box ---
bash tools/sweep_mfcc2.sh 13,4,2,0 2>&1 | grep cfg=
AFB200_MFCC_CAP=14 bash tools/sweep_mfcc2.sh 13,4,2,0 2>&1 | grep cfg=
cp audioflux_b200/csrc/kernels/mfcc_fused2.cu /tmp/new.cu; cp include/afb200_ext.h /tmp/new.h
cp tools/ab/mfcc_fused2_old.cu.txt audioflux_b200/csrc/kernels/mfcc_fused2.cu; cp tools/ab/afb200_ext_old.h.txt include/afb200_ext.h
echo OLD; bash tools/sweep_mfcc2.sh 13,4,2,0 2>&1 | grep cfg=
bash tools/sweep_mfcc2.sh 13,4,2,0 2>&1 | grep cfg=
cp /tmp/new.cu audioflux_b200/csrc/kernels/mfcc_fused2.cu; cp /tmp/new.h include/afb200_ext.h
echo NEW; bash tools/sweep_mfcc2.sh 13,4,2,0 2>&1 | grep cfg=
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_new.json 2>/dev/null; cut -c1-260 gpurun_out/bench_new.json

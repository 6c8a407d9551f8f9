set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
# 1. full GPU parity suite on the round's final code
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_final_pytest_gpu.log
# 2. the driver's bench line
timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; cut -c1-400 gpurun_out/r2_final_bench_n1.json; tail -3 gpurun_out/r2_final_bench_n1.err
# 3. launch list of the same command (no CPU legs under the profiler)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_final_bench_under_ncu.log 2>&1
# 4. full captures: MFCC kernel (summary + per line), CWT fused kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mfcc_fused2 -s 2 -c 1 -f -o gpurun_out/r2_final_mfcc2 python tools/mfcc_prof.py > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2_final_mfcc2.ncu-rep > gpurun_out/r2_final_mfcc_fused2_ncu_summary.txt; head -40 gpurun_out/r2_final_mfcc_fused2_ncu_summary.txt
ncu -i gpurun_out/r2_final_mfcc2.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_final_mfcc2_src.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cwt_fused -s 1 -c 1 -f -o gpurun_out/r2_final_cwt python tools/cwt_prof.py > /dev/null 2>&1
ncu -i gpurun_out/r2_final_cwt.ncu-rep --page raw --csv > gpurun_out/r2_final_cwt_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2_final_cwt_raw.csv > gpurun_out/r2_final_cwt_ncu.txt 2>&1; cat gpurun_out/r2_final_cwt_ncu.txt
ls -la gpurun_out

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "temporal or istft or reassign or cwt or pwt or wsst" > gpurun_out/r2e_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2e_pytest_gpu.log
bash tools/sweep_cwt.sh 4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cwt_fused -s 1 -c 1 -f -o gpurun_out/r2e_cwt python tools/cwt_prof.py > /dev/null 2>&1
ncu -i gpurun_out/r2e_cwt.ncu-rep --page raw --csv > gpurun_out/r2e_cwt_raw.csv 2>/dev/null
ncu -i gpurun_out/r2e_cwt.ncu-rep --page source --csv --print-source sass > gpurun_out/r2e_cwt_src.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2e_cwt_raw.csv

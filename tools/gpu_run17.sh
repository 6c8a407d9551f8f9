set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "mfcc or mel or bft or fused or spectrogram or cwt or wsst" > gpurun_out/r2c_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2c_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; cut -c1-330 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
bash tools/sweep_cwt.sh 2 3 4 6 8
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_cwt_fused -c 2 python tools/cwt_prof.py 2>&1 | grep -E "k_cwt|duration|dram__" | head -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mfcc_fused2 -s 2 -c 1 -f -o gpurun_out/r2c_mfcc2 python tools/mfcc_prof.py > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2c_mfcc2.ncu-rep > gpurun_out/r2c_mfcc_fused2_ncu_summary.txt; cat gpurun_out/r2c_mfcc_fused2_ncu_summary.txt
ncu -i gpurun_out/r2c_mfcc2.ncu-rep --page source --csv --print-source sass > gpurun_out/r2c_mfcc2_src.csv 2>/dev/null

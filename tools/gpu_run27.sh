set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AFB200_FUZZ_DUMP=gpurun_out timeout 900 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -s > gpurun_out/r2j_fuzz.log 2>&1; grep -n "compared\|passed\|failed" gpurun_out/r2j_fuzz.log | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fuzz.py > gpurun_out/r2j_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2j_pytest_gpu.log | cut -c1-300

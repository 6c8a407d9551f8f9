set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# full suite (no -x) with the table-upload fix, then the reassign file three more times (the flaky case)
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2g_pytest_gpu.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_reassign.py -m gpu -q 2>&1 | tail -2; done
timeout 300 python tools/sweep_variants.py 20 > gpurun_out/r2g_sweep_wait.txt 2>&1; cat gpurun_out/r2g_sweep_wait.txt

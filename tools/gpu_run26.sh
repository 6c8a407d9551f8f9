set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_deconv.py -m gpu -q -s > gpurun_out/r2i_fuzz.log 2>&1; tail -c 9000 gpurun_out/r2i_fuzz.log

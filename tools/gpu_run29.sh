set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^\[audioflux_b200\] error\|is error\|^$" > gpurun_out/r2l_pytest_gpu.log; tail -25 gpurun_out/r2l_pytest_gpu.log | cut -c1-300

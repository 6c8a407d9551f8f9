set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AFB200_FUZZ_DUMP=gpurun_out timeout 900 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -s -k "mfcc or istft or xxccstd or cqtpost or stream" > gpurun_out/r2k_fuzz.log 2>&1; grep -n "compared\|passed\|failed\|AssertionError: (" gpurun_out/r2k_fuzz.log | cut -c1-1500

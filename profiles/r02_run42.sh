#!/bin/bash
# compute-sanitizer synccheck + initcheck over the kernel suites (large-batch cases excluded for time)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
SEL="not 40000 and not 8192 and not large_batch and not 700 and not 16000 and not 20000"
timeout 1800 compute-sanitizer --tool synccheck --error-exitcode 1 \
  python -m pytest tests/test_gpu_rbm.py tests/test_gpu_fc_tc.py tests/test_gpu_conv_tc.py -m gpu -q -x -k "$SEL" > gpurun_out/r02X_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/r02X_sanitizer_synccheck.log
tail -4 gpurun_out/r02X_sanitizer_synccheck.log
timeout 1800 compute-sanitizer --tool initcheck --error-exitcode 1 \
  python -m pytest tests/test_gpu_rbm.py tests/test_gpu_fc_tc.py tests/test_gpu_conv_tc.py tests/test_gpu_optimizers.py -m gpu -q -x -k "$SEL" > gpurun_out/r02X_sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?" >> gpurun_out/r02X_sanitizer_initcheck.log
tail -4 gpurun_out/r02X_sanitizer_initcheck.log
grep -c "Uninitialized" gpurun_out/r02X_sanitizer_initcheck.log
echo done

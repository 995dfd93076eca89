#!/bin/bash
# compute-sanitizer memcheck over the whole kernel suite of the final code (large-batch cases excluded for time)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 1 \
  python -m pytest tests/test_gpu_rbm.py tests/test_gpu_fc_tc.py tests/test_gpu_conv_tc.py tests/test_gpu_net.py tests/test_gpu_optimizers.py tests/test_gpu_composites.py -m gpu -q -x \
  -k "not 40000 and not 8192 and not large_batch and not 700 and not 16000 and not 20000" > gpurun_out/r02M_sanitizer_memcheck_all.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02M_sanitizer_memcheck_all.log
tail -5 gpurun_out/r02M_sanitizer_memcheck_all.log
echo done

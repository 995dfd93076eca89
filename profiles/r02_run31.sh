#!/bin/bash
# compute-sanitizer over the kernels added at the end of round 2
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 \
  python -m pytest tests/test_gpu_fc_tc.py tests/test_gpu_optimizers.py tests/test_gpu_rbm.py tests/test_gpu_api.py -m gpu -q -x \
  -k "fc_warp or epoch_end or tensor_core or host_fed or epoch_launch" > gpurun_out/r02K_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02K_sanitizer_memcheck.log
tail -4 gpurun_out/r02K_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 \
  python -m pytest tests/test_gpu_fc_tc.py tests/test_gpu_optimizers.py tests/test_gpu_rbm.py -m gpu -q -x \
  -k "(fc_warp and N20) or epoch_end or (tensor_core and (777 or 500 or 300))" > gpurun_out/r02K_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02K_sanitizer_racecheck.log
tail -4 gpurun_out/r02K_sanitizer_racecheck.log
echo done

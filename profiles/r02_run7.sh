#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fc_tc.py tests/test_gpu_net.py tests/test_gpu_api.py -m gpu -q > gpurun_out/r02i_pytest_fcgrad.log 2>&1; echo "rc=$?" >> gpurun_out/r02i_pytest_fcgrad.log
tail -40 gpurun_out/r02i_pytest_fcgrad.log
for v in 0 1; do
  CGSVMC_FC_TC_GRAD=$v timeout 300 python bench_configs.py --configs c1 --reps 5 >> gpurun_out/r02i_configs_c1_grad${v}.jsonl 2>> gpurun_out/r02i.err
  CGSVMC_FC_TC_GRAD=$v timeout 300 python bench_configs.py --configs c1 --walkers 65536 --reps 5 >> gpurun_out/r02i_configs_c1_grad${v}.jsonl 2>> gpurun_out/r02i.err
done
tail -3 gpurun_out/r02i.err
echo done

#!/bin/bash
# conv_tc_grad.cu: first GPU run -- parity tests, then device times of C3 / C4 with both gradient paths
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q -x -k "grad or accumulate" 2>&1 | tail -30 > gpurun_out/r02l_pytest_convgrad.log
cat gpurun_out/r02l_pytest_convgrad.log
CGSVMC_CONV_TC_GRAD_M64=1 timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q -x -k "grad_sum_vs_oracle" 2>&1 | tail -15 > gpurun_out/r02l_pytest_convgrad_m64.log
cat gpurun_out/r02l_pytest_convgrad_m64.log
timeout 300 python bench_configs.py --configs c3,c4 --reps 5 > gpurun_out/r02l_configs_grad1.jsonl 2>> gpurun_out/r02l.err
CGSVMC_CONV_TC_GRAD=0 timeout 300 python bench_configs.py --configs c3,c4 --reps 5 > gpurun_out/r02l_configs_grad0.jsonl 2>> gpurun_out/r02l.err
CGSVMC_CONV_TC_GRAD_M64=1 timeout 300 python bench_configs.py --configs c3,c4 --reps 5 > gpurun_out/r02l_configs_grad1_m64.jsonl 2>> gpurun_out/r02l.err
tail -5 gpurun_out/r02l.err
cut -c1-420 gpurun_out/r02l_configs_grad1.jsonl gpurun_out/r02l_configs_grad0.jsonl gpurun_out/r02l_configs_grad1_m64.jsonl
echo done

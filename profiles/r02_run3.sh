#!/bin/bash
# Round-2 GPU pass 3: conv_tc interleaved layout A/B, rbm2 reduction variants.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_net.py -m gpu -x -q > gpurun_out/r02c_pytest_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_conv.log
tail -15 gpurun_out/r02c_pytest_conv.log
for il in 0 1; do for c in 1 2; do
  CGSVMC_CONV_TC_IL=$il CGSVMC_CONV_TC_CTAS=$c timeout 300 python bench_configs.py --configs c3,c4,c5conv --reps 3 > gpurun_out/r02c_configs_il${il}_ctas${c}.jsonl 2>> gpurun_out/r02c.err
done; done
for v in 0 1 2; do
  CGSVMC_RBM2_FUSED_REDUCE=$v timeout 300 python bench.py --steps 200 --warmup 20 --configs "" --no-cpu-baseline > gpurun_out/r02c_bench_reduce${v}.json 2>> gpurun_out/r02c.err
done
echo done

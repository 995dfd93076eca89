#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_net.py -m gpu -x -q 2>&1 | tail -3
for il in 0 1; do for c in 1 2; do
  CGSVMC_CONV_TC_IL=$il CGSVMC_CONV_TC_CTAS=$c timeout 300 python bench_configs.py --configs c3,c4,c5conv --reps 3 > gpurun_out/r02g_configs_il${il}_ctas${c}.jsonl 2>> gpurun_out/r02g.err
done; done
tail -3 gpurun_out/r02g.err
echo done

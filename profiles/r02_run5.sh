#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/r02e_*.jsonl
for w in 1024 65536; do
  timeout 300 python profiles/run_fc_tc_phases.py $w >> gpurun_out/r02e_fc_tc_phases.jsonl 2>> gpurun_out/r02e.err
done
timeout 600 python -m pytest tests/test_gpu_fc_tc.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
  CGSVMC_FC_TC=$v timeout 300 python bench_configs.py --configs c1 --reps 5 >> gpurun_out/r02e_configs_c1_fctc${v}.jsonl 2>> gpurun_out/r02e.err
  CGSVMC_FC_TC=$v timeout 300 python bench_configs.py --configs c1 --walkers 65536 --reps 5 >> gpurun_out/r02e_configs_c1_fctc${v}.jsonl 2>> gpurun_out/r02e.err
done
tail -3 gpurun_out/r02e.err
echo done

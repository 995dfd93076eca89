#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for l in 8 16; do
  export CGSVMC_RBM2_LPW=$l
  echo "== CGSVMC_RBM2_LPW=$l" >> gpurun_out/r02U_rbm2_lpw.jsonl
  RBM2_EPOCH_CONFIGS=C2 timeout 300 python profiles/run_rbm2_epoch.py >> gpurun_out/r02U_rbm2_lpw.jsonl 2>> gpurun_out/r02U.err
  timeout 300 python bench_configs.py --configs c2 --reps 3 >> gpurun_out/r02U_rbm2_lpw.jsonl 2>> gpurun_out/r02U.err
done
cut -c1-260 gpurun_out/r02U_rbm2_lpw.jsonl
tail -3 gpurun_out/r02U.err
echo done

#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02w_pytest_gpu.log
grep -E "Error|assert|passed|failed|FAILED" gpurun_out/r02w_pytest_gpu.log | head -30
export RBM2_EPOCH_CONFIGS=C2
for seg in 0 8 16 32 1000; do
  if [ $seg = 0 ]; then export CGSVMC_RBM2_TC_GRAD=0; else export CGSVMC_RBM2_TC_GRAD=1; export CGSVMC_RBM2_TC_SEGMENT=$seg; fi
  echo "== segment $seg" >> gpurun_out/r02w_rbm2_epoch.jsonl
  timeout 300 python profiles/run_rbm2_epoch.py >> gpurun_out/r02w_rbm2_epoch.jsonl 2>> gpurun_out/r02w.err
done
unset CGSVMC_RBM2_TC_GRAD
for seg in 8 16 32 1000; do
  export CGSVMC_RBM2_TC_SEGMENT=$seg
  echo "== segment $seg" >> gpurun_out/r02w_precision.txt
  python profiles/debug_tc_grad_precision.py 2>&1 | grep -E "^---|ALL|e_loc" | grep -v float64 | tail -6 >> gpurun_out/r02w_precision.txt
done
cat gpurun_out/r02w_rbm2_epoch.jsonl | cut -c1-200
cat gpurun_out/r02w_precision.txt
tail -n 5 gpurun_out/r02w.err
echo done

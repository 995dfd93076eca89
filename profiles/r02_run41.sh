#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rbm.py -m gpu -q 2>&1 | tail -5
RBM2_EPOCH_CONFIGS=C2,C5 timeout 300 python profiles/run_rbm2_epoch.py > gpurun_out/r02W_rbm2_epoch.jsonl 2>> gpurun_out/r02W.err
cut -c1-200 gpurun_out/r02W_rbm2_epoch.jsonl
timeout 300 python bench_configs.py --configs c5rbm --reps 3 2>> gpurun_out/r02W.err | cut -c1-400
tail -3 gpurun_out/r02W.err
echo done

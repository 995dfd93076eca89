#!/bin/bash
# multi-iteration walker kernel (cgsvmc_batch_steps): parity tests + device time per step
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rbm.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02m_pytest_rbm.log
cat gpurun_out/r02m_pytest_rbm.log
timeout 600 python profiles/run_rbm2_epoch.py > gpurun_out/r02m_rbm2_epoch.jsonl 2>> gpurun_out/r02m.err
cat gpurun_out/r02m_rbm2_epoch.jsonl
tail -5 gpurun_out/r02m.err
echo done

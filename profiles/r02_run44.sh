#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02Z_pytest_gpu.log
grep -E "Error|assert|passed|failed|FAILED" gpurun_out/r02Z_pytest_gpu.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== walker kernel: 512 threads for every variant (final)" >> gpurun_out/r02Z_rbm2_small_h.jsonl
timeout 600 python profiles/run_rbm2_small_h.py >> gpurun_out/r02Z_rbm2_small_h.jsonl 2>> gpurun_out/r02Z.err
cat gpurun_out/r02Z_rbm2_small_h.jsonl | cut -c1-240
RBM2_EPOCH_CONFIGS=C2 timeout 300 python profiles/run_rbm2_epoch.py 2>> gpurun_out/r02Z.err | cut -c1-200
tail -3 gpurun_out/r02Z.err
echo done

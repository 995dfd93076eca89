#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02v_pytest_gpu.log
grep -E "Error|assert|passed|failed|FAILED" gpurun_out/r02v_pytest_gpu.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python profiles/debug_tc_grad_precision.py 2>&1 | grep -E "^---|ALL|e_loc" | grep -v float64 | tail -12 > gpurun_out/r02v_precision.txt
cat gpurun_out/r02v_precision.txt
timeout 600 python profiles/run_rbm2_epoch.py > gpurun_out/r02v_rbm2_epoch.jsonl 2>> gpurun_out/r02v.err
head -4 gpurun_out/r02v_rbm2_epoch.jsonl
tail -n 5 gpurun_out/r02v.err
echo done

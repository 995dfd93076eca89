#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python profiles/debug_tc_grad_precision.py 2>&1 | grep -E "^---|ALL|e_loc" | grep -v float64 > gpurun_out/r02u_precision.txt
cat gpurun_out/r02u_precision.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02u_pytest_gpu.log
cat gpurun_out/r02u_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python profiles/run_rbm2_epoch.py 2>> gpurun_out/r02u.err | head -4 > gpurun_out/r02u_rbm2_epoch.jsonl
cat gpurun_out/r02u_rbm2_epoch.jsonl
tail -n 5 gpurun_out/r02u.err
echo done

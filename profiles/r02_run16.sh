#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r02q_pytest_gpu.log
cat gpurun_out/r02q_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02q_bench.json 2>> gpurun_out/r02q.err
cut -c1-600 gpurun_out/r02q_bench.json
tail -n 5 gpurun_out/r02q.err
echo done

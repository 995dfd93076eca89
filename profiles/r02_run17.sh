#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02s_pytest_gpu.log
cat gpurun_out/r02s_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python profiles/debug_tc_grad_precision.py 2>&1 | grep -E "^---|ALL" > gpurun_out/r02s_precision.txt
cat gpurun_out/r02s_precision.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02s_bench.json 2>> gpurun_out/r02s.err
cut -c1-400 gpurun_out/r02s_bench.json
tail -n 5 gpurun_out/r02s.err
echo done

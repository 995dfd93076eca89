#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02S_pytest_gpu.log
grep -E "Error|assert|passed|failed|FAILED" gpurun_out/r02S_pytest_gpu.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo done

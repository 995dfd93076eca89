#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02z_pytest_gpu.log
grep -E "Error|assert|passed|failed|FAILED" gpurun_out/r02z_pytest_gpu.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench_steps20.json 2>> gpurun_out/r02z.err
python - <<'PY'
import json
for l in open('gpurun_out/r02z_bench_steps20.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step','step_ms','epoch_end_ms','gpu_launches')}, d['e2e']['ms_per_step'])
PY
tail -3 gpurun_out/r02z.err
echo done

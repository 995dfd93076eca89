#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02h_pytest_gpu.log
grep -E "Error|assert|passed|failed|FAILED" gpurun_out/r02h_pytest_gpu.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02h_bench_steps20.json 2>> gpurun_out/r02h.err
python - <<'PY'
import json
for l in open('gpurun_out/r02h_bench_steps20.json'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']
        print({k:d.get(k) for k in ('value','ms_per_step','step_ms','epoch_end_ms','gpu_launches','vs_baseline')}, 'roof', round(d['roofline']['frac'],3), d['roofline'].get('traffic'))
        print('e2e', e['ms_per_step'], e['host_pack'], e['h2d_bytes_per_step'], e['d2h_bytes_per_step'])
        print(sorted(d.keys()))
PY
tail -2 gpurun_out/r02h.err

#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02C_pytest_gpu.log
grep -E "Error|assert|passed|failed|FAILED" gpurun_out/r02C_pytest_gpu.log | head -30
RBM2_EPOCH_CONFIGS=C2 timeout 300 python profiles/run_rbm2_epoch.py > gpurun_out/r02C_rbm2_epoch.jsonl 2>> gpurun_out/r02C.err
cut -c1-200 gpurun_out/r02C_rbm2_epoch.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02C_bench_steps20.json 2>> gpurun_out/r02C.err
python - <<'PY'
import json
for f in ('gpurun_out/r02C_bench_steps20.json',):
  for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']
        print({k:d.get(k) for k in ('value','ms_per_step','step_ms','epoch_end_ms')})
        print('e2e', e['ms_per_step'], 'host_pack', e['host_pack'], e['host_pack_probe'], 'other', e['other_upload_form']['ms_per_step'], 'packed input', e['packed_host_input']['ms_per_step'])
PY
tail -3 gpurun_out/r02C.err
echo done

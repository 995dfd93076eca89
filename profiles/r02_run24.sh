#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nproc
timeout 600 python -m pytest tests/test_gpu_rbm.py -m gpu -q -k "host_fed" 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 --configs "" --no-cpu-baseline > gpurun_out/r02L_bench_steps20.json 2>> gpurun_out/r02L.err
timeout 900 python bench.py --steps 200 --warmup 5 --configs "" --no-cpu-baseline > gpurun_out/r02L_bench_steps200.json 2>> gpurun_out/r02L.err
python - <<'PY'
import json
for f in ('gpurun_out/r02L_bench_steps20.json','gpurun_out/r02L_bench_steps200.json'):
  for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']
        print({k:d.get(k) for k in ('value','ms_per_step','step_ms','epoch_end_ms')})
        print('e2e', e['ms_per_step'], 'host_pack', e['host_pack'], e['host_pack_probe'], 'other', e['other_upload_form']['ms_per_step'], 'packed input', e['packed_host_input']['ms_per_step'])
PY
tail -3 gpurun_out/r02L.err
echo done

#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --configs "" > gpurun_out/r02Q_bench_n8.json 2> gpurun_out/r02Q_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/r02Q_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']
        print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','step_ms','epoch_end_ms','allreduce_ms','parity_ok')})
        print('e2e', e['ms_per_step'], e['host_pack'], e.get('host_pack_threads'), e['host_pack_probe'], 'other', e['other_upload_form']['ms_per_step'], 'packed', e['packed_host_input']['ms_per_step'])
PY
tail -3 gpurun_out/r02Q_n8.err
echo done

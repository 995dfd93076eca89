#!/bin/bash
# 2-GPU bench: collectives in the timed region, parity self-check, fused epoch end on the all-reduced payload
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02i_bench_n2.json 2> gpurun_out/r02i_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r02i_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','step_ms','epoch_end_ms','allreduce_ms','collectives_in_timed_region','parity_ok')})
        print('e2e', d['e2e']['ms_per_step'], d['e2e']['host_pack'])
        for k,v in d['configs'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('allreduce_ms'), v.get('error'))
PY
tail -5 gpurun_out/r02i_n2.err
echo done

#!/bin/bash
# 8-GPU bench line of the final code
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02J_bench_n$n.json 2> gpurun_out/r02J_n$n.err
done
python - <<'PY'
import json
for n in (8,4):
  for l in open('gpurun_out/r02J_bench_n%d.json'%n):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','step_ms','epoch_end_ms','allreduce_ms','collectives_in_timed_region','parity_ok')})
        print('e2e', d['e2e']['ms_per_step'], d['e2e']['host_pack'])
        for k,v in d['configs'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('allreduce_ms'), v.get('error'))
PY
tail -3 gpurun_out/r02J_n8.err
echo done

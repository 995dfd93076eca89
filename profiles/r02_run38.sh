#!/bin/bash
# config 5: walker-count sweep on 16x16 Heisenberg, RBM H = 256, one GPU
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench_configs.py --configs c5rbm --walker-sweep 4096,16384,65536,131072,262144,524288,1048576 --reps 3 > gpurun_out/r02T_c5_walker_sweep.jsonl 2> gpurun_out/r02T.err
python - <<'PY'
import json
for l in open('gpurun_out/r02T_c5_walker_sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['walkers'], 'sampler_ms', round(d['sampler_ms'],3), 'G wsps', round(d['walker_steps_per_sec']/1e9,3), 'eloc_ms', round(d['local_energy_ms'],3), 'M eloc/s', round(d['eloc_evals_per_sec']/1e6,2), 'acc_ms', round(d['accumulate_ms'],3))
PY
tail -3 gpurun_out/r02T.err
echo done

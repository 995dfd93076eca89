#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02G_pytest_gpu.log
grep -E "Error|assert|passed|failed|FAILED" gpurun_out/r02G_pytest_gpu.log | head -30
for w in 0 auto; do
  if [ $w = 0 ]; then export CGSVMC_FC_WARP=0; else unset CGSVMC_FC_WARP; fi
  echo "== CGSVMC_FC_WARP=$w" >> gpurun_out/r02G_configs_c1.jsonl
  for B in 1024 2048 4096; do
  timeout 300 python bench_configs.py --configs c1 --walkers $B --reps 3 >> gpurun_out/r02G_configs_c1.jsonl 2>> gpurun_out/r02G.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r02G_configs_c1.jsonl'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print(d['walkers'], 'sampler_ms', round(d['sampler_ms'],4), 'eloc', round(d['local_energy_ms'],4), 'acc', round(d['accumulate_ms'],4))
PY
tail -3 gpurun_out/r02G.err
echo done

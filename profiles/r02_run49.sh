#!/bin/bash
# robustness of bench.py against unusual --steps / --warmup
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for args in "--steps 77 --warmup 3" "--steps 20 --warmup 3" "--steps 51 --warmup 3"; do
  echo "== $args"
  timeout 600 python bench.py $args --configs "" --no-cpu-baseline 2>> gpurun_out/r02e.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d.get(k) for k in ('steps', 'warmup', 'value', 'ms_per_step', 'step_ms', 'epoch_len', 'gpu_launches')}, 'e2e', round(d['e2e']['ms_per_step'], 4))
"
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>> gpurun_out/r02e.err | cut -c1-300
tail -3 gpurun_out/r02e.err
echo done

#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for lib in libcgsvmc.so libcgsvmc_cgsvmc_rbm2_reread.so; do
  echo "== $lib" >> gpurun_out/r02a_rbm2_reread.jsonl
  CGSVMC_LIBRARY=cgs_vmc_b200/$lib RBM2_EPOCH_CONFIGS=C2,C5 timeout 300 python profiles/run_rbm2_epoch.py >> gpurun_out/r02a_rbm2_reread.jsonl 2>> gpurun_out/r02a.err
  CGSVMC_LIBRARY=cgs_vmc_b200/$lib timeout 300 python bench_configs.py --configs c5rbm --reps 3 2>> gpurun_out/r02a.err | python -c "import sys,json; [print({k:round(v,3) for k,v in json.loads(l).items() if k in ('sampler_ms','local_energy_ms','accumulate_ms')}) for l in sys.stdin if l.startswith('{')]" >> gpurun_out/r02a_rbm2_reread.jsonl
done
cut -c1-200 gpurun_out/r02a_rbm2_reread.jsonl
CGSVMC_LIBRARY=cgs_vmc_b200/libcgsvmc_cgsvmc_rbm2_reread.so timeout 900 python -m pytest tests/test_gpu_rbm.py -m gpu -q 2>&1 | tail -3
tail -3 gpurun_out/r02a.err
echo done

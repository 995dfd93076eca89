#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for lib in libcgsvmc.so libcgsvmc_cgsvmc_rbm2_all_512.so; do
  echo "== $lib" >> gpurun_out/r02Y_rbm2_small_h.jsonl
  CGSVMC_LIBRARY=cgs_vmc_b200/$lib timeout 600 python profiles/run_rbm2_small_h.py >> gpurun_out/r02Y_rbm2_small_h.jsonl 2>> gpurun_out/r02Y.err
done
cat gpurun_out/r02Y_rbm2_small_h.jsonl
CGSVMC_LIBRARY=cgs_vmc_b200/libcgsvmc_cgsvmc_rbm2_all_512.so timeout 900 python -m pytest tests/test_gpu_rbm.py -m gpu -q 2>&1 | tail -4
tail -3 gpurun_out/r02Y.err
echo done

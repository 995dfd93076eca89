#!/bin/bash
# Round-2 GPU pass 4: fc_tc (tcgen05 fully connected), conv_tc with two activation planes.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fc_tc.py -m gpu -x -q > gpurun_out/r02d_pytest_fc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest_fc.log
tail -30 gpurun_out/r02d_pytest_fc.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest_gpu.log
tail -30 gpurun_out/r02d_pytest_gpu.log
for v in 0 1; do
  CGSVMC_FC_TC=$v timeout 300 python bench_configs.py --configs c1 --reps 5 >> gpurun_out/r02d_configs_c1_fctc${v}.jsonl 2>> gpurun_out/r02d.err
  CGSVMC_FC_TC=$v timeout 300 python bench_configs.py --configs c1 --walkers 65536 --reps 5 >> gpurun_out/r02d_configs_c1_fctc${v}.jsonl 2>> gpurun_out/r02d.err
done
timeout 300 python bench_configs.py --configs c3,c4,c5conv --reps 3 > gpurun_out/r02d_configs_conv.jsonl 2>> gpurun_out/r02d.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2>> gpurun_out/r02d.err
echo done

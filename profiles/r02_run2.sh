#!/bin/bash
# Round-2 GPU pass 2: in-kernel reduction + host-fed step (tests, bench), conv_tc phase counters.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_gpu.log
tail -15 gpurun_out/r02b_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02b_bench.err
CGSVMC_RBM2_FUSED_REDUCE=0 timeout 600 python bench.py --steps 20 --warmup 5 --configs "" --no-cpu-baseline > gpurun_out/r02b_bench_nofuse.json 2>> gpurun_out/r02b_bench.err
timeout 600 python bench.py --steps 200 --warmup 20 --configs "" --no-cpu-baseline > gpurun_out/r02b_bench_steps200.json 2>> gpurun_out/r02b_bench.err
for c in 1 2; do
  timeout 300 python profiles/run_conv_tc_phases.py --ctas $c >> gpurun_out/r02b_conv_tc_phases.jsonl 2>> gpurun_out/r02b_bench.err
done
CGSVMC_LIBRARY=cgs_vmc_b200/libcgsvmc_timing.so timeout 300 python profiles/run_rbm2_phases.py > gpurun_out/r02b_rbm2_phases_fused.json 2>> gpurun_out/r02b_bench.err
echo done

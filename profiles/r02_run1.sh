#!/bin/bash
# Round-2 GPU pass 1: GPU test suite, bench line, ncu launch list of the bench.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
tail -5 gpurun_out/r02a_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02a_bench.err
python bench.py --steps 200 --warmup 20 --configs "" > gpurun_out/r02a_bench_steps200.json 2>> gpurun_out/r02a_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_bench_steps5.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --configs "" > gpurun_out/r02a_bench_under_ncu.json 2> gpurun_out/r02a_ncu.err
CGSVMC_CONV_TC_CTAS=1 python bench_configs.py --configs c3,c4 --reps 3 > gpurun_out/r02a_configs_tc_ctas1.jsonl 2>> gpurun_out/r02a_bench.err
CGSVMC_CONV_TC_CTAS=2 python bench_configs.py --configs c3,c4 --reps 3 > gpurun_out/r02a_configs_tc_ctas2.jsonl 2>> gpurun_out/r02a_bench.err
echo done

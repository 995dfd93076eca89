#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests/test_gpu_rbm.py -m gpu -q 2>&1 | tail -4
timeout 300 python bench_configs.py --configs c5rbm --reps 5 > gpurun_out/r02k_configs_c5.jsonl 2>> gpurun_out/r02k.err
CGSVMC_RBM2_NO_PAIR_TABLE=1 timeout 300 python bench_configs.py --configs c5rbm --reps 5 > gpurun_out/r02k_configs_c5_nopair.jsonl 2>> gpurun_out/r02k.err
NCU="ncu --clock-control none"
$NCU --set full --import-source on --kernel-name-base demangled -k "regex:walker_kernel<1, 8, 5, 1, 1" --launch-skip 5 -c 1 -f -o /tmp/ncu/r02k_rbm2_fused \
  python bench.py --steps 3 --warmup 3 --configs "" --no-cpu-baseline > /dev/null 2>> gpurun_out/r02k_ncu.err
python profiles/summarize_ncu.py /tmp/ncu/r02k_rbm2_fused.ncu-rep > gpurun_out/r02k_rbm2_fused_ncu_full.txt 2>> gpurun_out/r02k_ncu.err
python profiles/source_hotspots.py /tmp/ncu/r02k_rbm2_fused.ncu-rep "walker_kernel" 30 > gpurun_out/r02k_rbm2_fused_hotspots.txt 2>> gpurun_out/r02k_ncu.err
tail -3 gpurun_out/r02k_ncu.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo done

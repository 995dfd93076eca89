#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k "regex:walker_kernel" --launch-skip 4 -c 1 -f -o /tmp/ncu/r02k_rbm2_fused \
  python bench.py --steps 3 --warmup 3 --configs "" --no-cpu-baseline > /dev/null 2>> gpurun_out/r02k_ncu.err
python profiles/summarize_ncu.py /tmp/ncu/r02k_rbm2_fused.ncu-rep > gpurun_out/r02k_rbm2_fused_ncu_full.txt 2>> gpurun_out/r02k_ncu.err
python profiles/source_hotspots.py /tmp/ncu/r02k_rbm2_fused.ncu-rep "walker_kernel" 30 > gpurun_out/r02k_rbm2_fused_hotspots.txt 2>> gpurun_out/r02k_ncu.err
tail -3 gpurun_out/r02k_ncu.err
echo done

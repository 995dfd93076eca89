#!/bin/bash
# phases of the persistent epoch kernel (timing build) + ncu of the conv gradient kernel + bench line
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
CGSVMC_LIBRARY=cgs_vmc_b200/libcgsvmc_timing.so timeout 300 python profiles/run_rbm2_phases.py 8192 --epoch 20 > gpurun_out/r02n_rbm2_phases_epoch20.json 2>> gpurun_out/r02n.err
cat gpurun_out/r02n_rbm2_phases_epoch20.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02n_bench.json 2>> gpurun_out/r02n.err
cut -c1-1500 gpurun_out/r02n_bench.json
NCU="ncu --clock-control none"
$NCU --set full --import-source on --kernel-name-base demangled -k "regex:conv_grad_tc_kernel" --launch-skip 2 -c 1 -f -o /tmp/ncu/r02n_conv_grad_tc \
  python bench_configs.py --configs c3 --reps 2 > /dev/null 2>> gpurun_out/r02n_ncu.err
python profiles/summarize_ncu.py /tmp/ncu/r02n_conv_grad_tc.ncu-rep > gpurun_out/r02n_conv_grad_tc_ncu_full.txt 2>> gpurun_out/r02n_ncu.err
python profiles/source_hotspots.py /tmp/ncu/r02n_conv_grad_tc.ncu-rep "conv_grad_tc_kernel" 30 > gpurun_out/r02n_conv_grad_tc_hotspots.txt 2>> gpurun_out/r02n_ncu.err
head -40 gpurun_out/r02n_conv_grad_tc_ncu_full.txt
tail -3 gpurun_out/r02n_ncu.err gpurun_out/r02n.err
echo done

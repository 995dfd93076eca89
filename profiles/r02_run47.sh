#!/bin/bash
# fresh ncu --set full captures of every hot kernel family on the final code
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 regex=$2 skip=$3 count=$4; shift 4
  $NCU --set full --import-source on -k "regex:$regex" --launch-skip $skip -c $count -f -o /tmp/ncu/$name "$@" > /dev/null 2>> gpurun_out/r02c_ncu.err
  python profiles/summarize_ncu.py /tmp/ncu/$name.ncu-rep > gpurun_out/${name}_ncu_full.txt 2>> gpurun_out/r02c_ncu.err
  python profiles/source_hotspots.py /tmp/ncu/$name.ncu-rep "$regex" 30 > gpurun_out/${name}_hotspots.txt 2>> gpurun_out/r02c_ncu.err
}
cap r02c_conv_c3 'tc_mc_kernel|tc_eloc_kernel|conv_grad_tc' 2 4 python bench_configs.py --configs c3 --reps 1
cap r02c_fc_c1_65536 'fc_mc_kernel|fc_eloc_kernel|fc_grad_kernel' 2 4 python bench_configs.py --configs c1 --walkers 65536 --reps 1
cap r02c_rbm2_c5 'walker_kernel|mc_kernel' 2 3 python bench_configs.py --configs c5rbm --reps 1
tail -5 gpurun_out/r02c_ncu.err
for f in gpurun_out/r02c_*_ncu_full.txt; do echo == $f; grep -E "^kernel|duration|issue_active|tensor_cycles|l1tex__throughput|lts__throughput|dram__bytes_read" $f | head -30; done
echo done

#!/bin/bash
# Round-2 evidence pass: ncu launch list + --set full captures of every hot kernel
# (summarised on the box: the .ncu-rep files stay in /tmp, only text comes back),
# in-kernel phase counters, compute-sanitizer memcheck / racecheck of the GPU tests.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
timeout 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_composites.py -m gpu -x -q 2>&1 | tail -3
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r02h_launches_bench_steps5.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_under_ncu.json 2> gpurun_out/r02h_ncu.err
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 regex=$2 skip=$3 count=$4; shift 4
  $NCU --set full --import-source on -k "regex:$regex" --launch-skip $skip -c $count -f -o /tmp/ncu/$name "$@" > /dev/null 2>> gpurun_out/r02h_ncu.err
  python profiles/summarize_ncu.py /tmp/ncu/$name.ncu-rep > gpurun_out/${name}_ncu_full.txt 2>> gpurun_out/r02h_ncu.err
  python profiles/source_hotspots.py /tmp/ncu/$name.ncu-rep "$regex" 30 > gpurun_out/${name}_hotspots.txt 2>> gpurun_out/r02h_ncu.err
}
cap r02h_rbm2_fused 'walker_kernel' 8 1 python bench.py --steps 3 --warmup 3 --configs "" --no-cpu-baseline
cap r02h_conv_c3 'tc_mc_kernel|tc_eloc_kernel|conv_grad' 2 4 python bench_configs.py --configs c3 --reps 1
cap r02h_fc_c1_65536 'fc_mc_kernel|fc_eloc_kernel|mlp_grad' 2 4 python bench_configs.py --configs c1 --walkers 65536 --reps 1
cap r02h_rbm2_c5 'walker_kernel|mc_kernel' 2 3 python bench_configs.py --configs c5rbm --reps 1
tail -5 gpurun_out/r02h_ncu.err
for c in 2 1; do
  timeout 300 python profiles/run_conv_tc_phases.py --ctas $c >> gpurun_out/r02h_conv_tc_phases.jsonl 2>> gpurun_out/r02h.err
done
CGSVMC_LIBRARY=cgs_vmc_b200/libcgsvmc_timing.so timeout 300 python profiles/run_rbm2_phases.py > gpurun_out/r02h_rbm2_phases_fused.json 2>> gpurun_out/r02h.err
timeout 300 python profiles/run_fc_tc_phases.py 65536 > gpurun_out/r02h_fc_tc_phases.jsonl 2>> gpurun_out/r02h.err
timeout 300 python bench_configs.py --configs c1,c3,c4,c5rbm,c5conv --reps 3 > gpurun_out/r02h_configs.jsonl 2>> gpurun_out/r02h.err
# compute-sanitizer: memcheck over the kernel parity tests, racecheck over a smaller selection
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 \
  python -m pytest tests/test_gpu_rbm.py tests/test_gpu_fc_tc.py tests/test_gpu_conv_tc.py tests/test_gpu_net.py -m gpu -q -x \
  -k "not 40000 and not 8192 and not large_batch and not 700" > gpurun_out/r02h_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02h_sanitizer_memcheck.log
tail -4 gpurun_out/r02h_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 \
  python -m pytest tests/test_gpu_rbm.py tests/test_gpu_fc_tc.py tests/test_gpu_conv_tc.py -m gpu -q -x \
  -k "golden or in_kernel or (log_amp and 127)" > gpurun_out/r02h_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02h_sanitizer_racecheck.log
tail -4 gpurun_out/r02h_sanitizer_racecheck.log
du -sh gpurun_out
echo done

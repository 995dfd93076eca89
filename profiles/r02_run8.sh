#!/bin/bash
# Round-2 evidence pass 2: fused C2 kernel, fc_tc gradient, conv gradient under ncu --set full; final bench lines.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 regex=$2 skip=$3 count=$4; shift 4
  $NCU --set full --import-source on -k "regex:$regex" --launch-skip $skip -c $count -f -o /tmp/ncu/$name "$@" > /dev/null 2>> gpurun_out/r02j_ncu.err
  python profiles/summarize_ncu.py /tmp/ncu/$name.ncu-rep > gpurun_out/${name}_ncu_full.txt 2>> gpurun_out/r02j_ncu.err
  python profiles/source_hotspots.py /tmp/ncu/$name.ncu-rep "$regex" 30 > gpurun_out/${name}_hotspots.txt 2>> gpurun_out/r02j_ncu.err
}
cap r02j_rbm2_fused 'walker_kernel<1, 8, 5, 1, 1' 5 1 python bench.py --steps 3 --warmup 3 --configs "" --no-cpu-baseline
cap r02j_fc_grad 'fc_grad_kernel' 1 1 python bench_configs.py --configs c1 --walkers 65536 --reps 1
cap r02j_conv_grad 'conv_grad_kernel' 1 1 python bench_configs.py --configs c3 --reps 1
tail -3 gpurun_out/r02j_ncu.err
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02j_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest_gpu.log
tail -4 gpurun_out/r02j_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
timeout 600 python bench.py --steps 200 --warmup 20 --configs "" > gpurun_out/r02j_bench_steps200.json 2>> gpurun_out/r02j_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02j_bench_reference.json 2>> gpurun_out/r02j_bench.err
timeout 300 python bench_configs.py --configs c1,c3,c4,c5rbm,c5conv --reps 3 > gpurun_out/r02j_configs.jsonl 2>> gpurun_out/r02j_bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo done

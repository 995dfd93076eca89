#!/bin/bash
# evidence pass on the final rbm2 kernel (tensor-core gradient): tests, sanitizer, ncu of the epoch kernel, bench
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_rbm.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r02y_pytest.log
cat gpurun_out/r02y_pytest.log
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k "regex:walker_kernel" --launch-skip 3 -c 1 -f -o /tmp/ncu/r02y_rbm2_epoch python profiles/run_rbm2_epoch_once.py > /dev/null 2>> gpurun_out/r02y_ncu.err
python profiles/summarize_ncu.py /tmp/ncu/r02y_rbm2_epoch.ncu-rep > gpurun_out/r02y_rbm2_epoch_ncu_full.txt 2>> gpurun_out/r02y_ncu.err
python profiles/source_hotspots.py /tmp/ncu/r02y_rbm2_epoch.ncu-rep walker_kernel 40 > gpurun_out/r02y_rbm2_epoch_hotspots.txt 2>> gpurun_out/r02y_ncu.err
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02y_launches_bench_steps20.csv \
  python bench.py --steps 20 --warmup 3 --configs "" --no-cpu-baseline > gpurun_out/r02y_bench_under_ncu.json 2>> gpurun_out/r02y_ncu.err
tail -3 gpurun_out/r02y_ncu.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 \
  python -m pytest tests/test_gpu_rbm.py -m gpu -q -x -k "tensor_core or batch_steps or graphed or fused or host_fed" > gpurun_out/r02y_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02y_sanitizer_memcheck.log
tail -4 gpurun_out/r02y_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 \
  python -m pytest tests/test_gpu_rbm.py -m gpu -q -x -k "tensor_core and (777 or 130 or 300)" > gpurun_out/r02y_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02y_sanitizer_racecheck.log
tail -4 gpurun_out/r02y_sanitizer_racecheck.log
timeout 900 python bench.py > gpurun_out/r02y_bench.json 2>> gpurun_out/r02y.err
tail -c 600 gpurun_out/r02y_bench.json
tail -3 gpurun_out/r02y.err
echo done

#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r02x_pytest_api.log
cat gpurun_out/r02x_pytest_api.log
export CGSVMC_LIBRARY=cgs_vmc_b200/libcgsvmc_timing.so
timeout 300 python profiles/run_rbm2_phases.py > gpurun_out/r02x_rbm2_phases_step.json 2>> gpurun_out/r02x.err
timeout 300 python profiles/run_rbm2_phases.py 8192 --epoch 20 > gpurun_out/r02x_rbm2_phases_epoch20.json 2>> gpurun_out/r02x.err
unset CGSVMC_LIBRARY
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k "regex:walker_kernel" --launch-skip 4 -c 1 -f -o /tmp/ncu/r02x_rbm2_epoch python bench.py --steps 20 --warmup 3 --configs "" --no-cpu-baseline > /dev/null 2>> gpurun_out/r02x_ncu.err
python profiles/summarize_ncu.py /tmp/ncu/r02x_rbm2_epoch.ncu-rep > gpurun_out/r02x_rbm2_epoch_ncu_full.txt 2>> gpurun_out/r02x_ncu.err
python profiles/source_hotspots.py /tmp/ncu/r02x_rbm2_epoch.ncu-rep walker_kernel 40 > gpurun_out/r02x_rbm2_epoch_hotspots.txt 2>> gpurun_out/r02x_ncu.err
tail -3 gpurun_out/r02x_ncu.err
tail -3 gpurun_out/r02x.err
echo done

#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rbm.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r02D_pytest.log
cat gpurun_out/r02D_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python profiles/debug_tc_grad_precision.py 2>&1 | grep -E "^---|ALL|e_loc" | grep -v float64 | tail -6
RBM2_EPOCH_CONFIGS=C2 timeout 300 python profiles/run_rbm2_epoch.py > gpurun_out/r02D_rbm2_epoch.jsonl 2>> gpurun_out/r02D.err
cut -c1-200 gpurun_out/r02D_rbm2_epoch.jsonl
CGSVMC_LIBRARY=cgs_vmc_b200/libcgsvmc_timing.so timeout 300 python profiles/run_rbm2_phases.py > gpurun_out/r02D_rbm2_phases_step.json 2>> gpurun_out/r02D.err
grep -A3 "tiles done" gpurun_out/r02D_rbm2_phases_step.json | head -8
grep "first_entry" gpurun_out/r02D_rbm2_phases_step.json
tail -3 gpurun_out/r02D.err
echo done

#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rbm.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02p_pytest_rbm.log
cat gpurun_out/r02p_pytest_rbm.log
timeout 600 python profiles/run_rbm2_epoch.py > gpurun_out/r02p_rbm2_epoch.jsonl 2>> gpurun_out/r02p.err
cat gpurun_out/r02p_rbm2_epoch.jsonl
CGSVMC_RBM2_TC_GRAD=0 timeout 600 python profiles/run_rbm2_epoch.py > gpurun_out/r02p_rbm2_epoch_simt.jsonl 2>> gpurun_out/r02p.err
head -4 gpurun_out/r02p_rbm2_epoch_simt.jsonl
python __graft_entry__.py --timing > /dev/null 2>&1
CGSVMC_LIBRARY=cgs_vmc_b200/libcgsvmc_timing.so timeout 300 python profiles/run_rbm2_phases.py 8192 --epoch 20 > gpurun_out/r02p_rbm2_phases_epoch20.json 2>> gpurun_out/r02p.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02p_rbm2_phases_epoch20.json'))
for k,v in d['last_iteration_phases'].items(): print('%-70s %6.2f %6.2f'%(k,v['median_us'],v['max_us']))
print(d['us_per_iteration'])
PY
tail -n 5 gpurun_out/r02p.err
echo done

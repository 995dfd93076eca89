#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rbm.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r02N_pytest.log
cat gpurun_out/r02N_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python profiles/debug_tc_grad_precision.py 2>&1 | grep -E "^---|ALL|e_loc" | grep -v float64 | tail -6
for f in 1 2; do
  export CGSVMC_RBM2_TC_GRAD=$f
  echo "== CGSVMC_RBM2_TC_GRAD=$f (1: free-running warps, 2: CTA barriers + centred weights)" >> gpurun_out/r02N_rbm2_epoch.jsonl
  RBM2_EPOCH_CONFIGS=C2 timeout 300 python profiles/run_rbm2_epoch.py >> gpurun_out/r02N_rbm2_epoch.jsonl 2>> gpurun_out/r02N.err
done
unset CGSVMC_RBM2_TC_GRAD
cut -c1-200 gpurun_out/r02N_rbm2_epoch.jsonl
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 \
  python -m pytest tests/test_gpu_rbm.py -m gpu -q -x -k "tensor_core and (777 or 500 or 300)" > gpurun_out/r02N_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02N_sanitizer_racecheck.log
tail -4 gpurun_out/r02N_sanitizer_racecheck.log
tail -3 gpurun_out/r02N.err
echo done

#!/bin/bash
# flakiness soak: the GPU suite five times back to back (different process each time)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -1 | tee -a gpurun_out/r02d_soak.log
done
for i in 1 2 3; do python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/r02d_soak.log; done
echo done

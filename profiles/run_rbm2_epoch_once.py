"""One GraphedEpoch (cgsvmc_batch_steps: the persistent C2 kernel, 20 batch
iterations per launch) replayed a few times -- the target of the ncu capture
of the headline kernel (profiles/r02_run22.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cgs_vmc_b200 import engine  # noqa: E402

dev = torch.device('cuda', 0)
w = bench.EnergyGradientWorkload('C2', 8192, 0, 1, dev)
g = engine.GraphedEpoch(w.state, w.ansatz, w.ham, w.sums, w.sweep_steps, 20)
for _ in range(4):
  g.replay()
torch.cuda.synchronize()
print('ok', float(w.sums.stats[0] / w.sums.stats[2]))

"""Device time per batch iteration of the C2 workload: one captured graph per
step (cgsvmc_batch_step) against one persistent kernel per epoch
(cgsvmc_batch_steps).  L2 flushed between launches.  One JSON line per variant."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cgs_vmc_b200 import engine  # noqa: E402

dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
only = os.environ.get('RBM2_EPOCH_CONFIGS', 'C2,C5').split(',')
for name, walkers in (('C2', 8192), ('C5', 131072)):
  if name not in only:
    continue
  w = bench.EnergyGradientWorkload(name, walkers, 0, 1, dev)
  ev = lambda: torch.cuda.Event(enable_timing=True)
  for _ in range(5):
    w.graphed.replay(); flush.zero_()
  marks = [(ev(), ev()) for _ in range(20)]
  for a, b in marks:
    a.record(); w.graphed.replay(); b.record(); flush.zero_()
  torch.cuda.synchronize()
  per_step = float(np.mean([a.elapsed_time(b) for a, b in marks]))
  print(json.dumps({'config': name, 'variant': 'one graph per step (cgsvmc_batch_step)', 'ms_per_step': per_step}))
  a0, b0 = ev(), ev()
  a0.record()
  for _ in range(50):
    w.graphed.replay()
  b0.record()
  torch.cuda.synchronize()
  print(json.dumps({'config': name, 'variant': 'one graph per step, 50 replays back to back (no L2 flush)',
                    'ms_per_step': a0.elapsed_time(b0) / 50}))
  for nb in (5, 20, 50):
    if name == 'C5' and nb > 20:
      continue
    g = engine.GraphedEpoch(w.state, w.ansatz, w.ham, w.sums, w.sweep_steps, nb)
    for _ in range(2):
      g.replay(); flush.zero_()
    marks = [(ev(), ev()) for _ in range(5)]
    for a, b in marks:
      a.record(); g.replay(); b.record(); flush.zero_()
    torch.cuda.synchronize()
    t = float(np.mean([a.elapsed_time(b) for a, b in marks]))
    print(json.dumps({'config': name, 'variant': 'one persistent kernel per %d steps (cgsvmc_batch_steps)' % nb,
                      'ms_per_launch': t, 'ms_per_step': t / nb}))
  w.sums.reset()

#!/usr/bin/env python
"""Driver for ncu captures of the convolutional gradient kernel (C3 shape)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgs_vmc_b200 import _native, engine   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
a = _native.Ansatz('conv_2d', 100, num_layers=5, num_filters=16, kernel_size=5, size_x=10, size_y=10)
a.set_params(torch.randn(a.num_params, generator=torch.Generator().manual_seed(1)) / math.sqrt(400))
state = engine.WalkerState(B, 100, seed=3)
w = torch.ones(2, B, device='cuda')
for _ in range(2):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  out = a.weighted_grad_sum(state.packed, w)
  e1.record()
  torch.cuda.synchronize()
  print('grad ms', e0.elapsed_time(e1))

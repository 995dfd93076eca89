#!/usr/bin/env python
"""Small driver for ncu captures of the tensor-core convolution kernels (C3
shape): one log_amp, one local_energy and a few sampler steps."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_configs import square_bonds   # noqa: E402
from cgs_vmc_b200 import _native, engine   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
a = _native.Ansatz('conv_2d', 100, num_layers=5, num_filters=16, kernel_size=5, size_x=10, size_y=10)
a.set_params(torch.randn(a.num_params, generator=torch.Generator().manual_seed(1)) / math.sqrt(400))
ij, jx, jz = square_bonds(10, True)
ham = _native.Hamiltonian(ij, jx, jz, 100)
state = engine.WalkerState(B, 100, seed=3)
for _ in range(2):
  z = a.log_amp(state.packed)
  e, _ = a.local_energy(ham, state.packed[:256].contiguous())
  state.mc_steps(a, 4)
torch.cuda.synchronize()
print('ok', float(z.mean()), float(e.mean()))

#!/usr/bin/env python
"""Where the C2 step goes: device time of cgsvmc_mc_steps against n_steps
(slope = per-step cost, intercept = launch + table load + state build) and of
the walker-kernel phases (local energy only / gradient only / accumulate),
warm and with the L2 flushed before each launch.

  python profiles/run_rbm2_scaling.py [walkers]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgs_vmc_b200 import _native, engine, lattices, wavefunctions   # noqa: E402


def timed(fn, flush, reps=30):
  ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
  for _ in range(5):
    fn()
  for a, b in ev:
    if flush is not None:
      flush.zero_()
    a.record()
    fn()
    b.record()
  torch.cuda.synchronize()
  t = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3
  return float(np.median(t))


def main():
  B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
  N, H = 36, 144
  _native.require_cuda()
  ansatz = _native.Ansatz('rbm', N, num_layers=0, layer_size=H)
  gen = torch.Generator().manual_seed(1234)
  shapes = [(N, 1), (1,), (N, H), (H,)]
  ansatz.set_params(torch.cat([t.reshape(-1) for t in wavefunctions._sonnet_init(shapes, gen)]))
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(6, 6), -1.0, 1.0)
  ham = _native.Hamiltonian(ij, jx, jz, N)
  state = engine.WalkerState(B, N)
  sums = engine.EnergyGradientSums(ansatz, B)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  w = torch.ones(2, B, device='cuda')
  out = {'walkers': B}
  for mode, fl in (('warm', None), ('l2_flushed', flush)):
    r = {}
    for n in (0, 1, 36, 72, 144, 360):
      r['mc_steps_%d_us' % n] = timed(lambda: state.mc_steps(ansatz, n), fl)
    r['mc_us_per_step'] = (r['mc_steps_360_us'] - r['mc_steps_36_us']) / 324.0
    r['local_energy_us'] = timed(lambda: ansatz.local_energy(ham, state.packed), fl)
    r['weighted_grad_sum_k2_us'] = timed(lambda: ansatz.weighted_grad_sum(state.packed, w), fl)
    r['accumulate_us'] = timed(lambda: sums.accumulate(ham, state.packed), fl)
    r['log_amp_us'] = timed(lambda: ansatz.log_amp(state.packed), fl)
    out[mode] = r
  print(json.dumps(out))


if __name__ == '__main__':
  main()

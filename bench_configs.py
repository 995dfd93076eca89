#!/usr/bin/env python
"""Device-time measurements of the other BASELINE.json configurations (C1, C3,
C4 shape, C5) on one GPU; `bench.py` is the contract benchmark (C2).  Prints
one JSON line per configuration:

  python bench_configs.py [--configs c1,c3,c5rbm,c5conv] [--reps 5]

Each line reports the sampler (one sweep = N Metropolis steps per walker), the
local energy and the fused accumulate (E_loc + both gradient sums) with CUDA
events on the launching stream, inputs resident in HBM, after warm-up.
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)


def square_bonds(size, nnn=False):
  nn, d = [], []
  for x in range(size):
    for y in range(size):
      s = x * size + y
      nn.append((s, ((x + 1) % size) * size + y))
      nn.append((s, x * size + (y + 1) % size))
      d.append((s, ((x + 1) % size) * size + (y + 1) % size))
      d.append((s, ((x + 1) % size) * size + (y - 1) % size))
  ij = nn + (d if nnn else [])
  jx = [-1.0] * len(nn) + ([0.5] * len(d) if nnn else [])
  jz = [1.0] * len(nn) + ([0.5] * len(d) if nnn else [])
  return np.asarray(ij, np.int32), np.asarray(jx, np.float32), np.asarray(jz, np.float32)


CONFIGS = {
    # name: (ansatz kwargs, lattice, walkers, flop per forward)
    'c1': dict(kind='fully_connected', n=20, kw=dict(num_layers=3, layer_size=80), chain=True,
               walkers=1024, f_fwd=28960, desc='C1: chain-20 Heisenberg, fully_connected 20-80-80-80-1, 1024 walkers'),
    'c3': dict(kind='conv_2d', n=100, kw=dict(num_layers=5, num_filters=16, kernel_size=5, size_x=10, size_y=10),
               size=10, nnn=True, walkers=8192, f_fwd=5.2e6,
               desc='C3: 10x10 J1-J2 (J2=0.5), conv_2d 5x16x5x5, 8192 walkers (one GPU share of 65536)'),
    'c4': dict(kind='conv_2d', n=36, kw=dict(num_layers=5, num_filters=16, kernel_size=5, size_x=6, size_y=6),
               size=6, nnn=True, walkers=8192, f_fwd=1.872e6,
               desc='C4 shape: 6x6 J1-J2, conv_2d trainee, 8192 walkers'),
    'c5rbm': dict(kind='rbm', n=256, kw=dict(num_layers=0, layer_size=256), size=16, nnn=False,
                  walkers=131072, f_fwd=131584,
                  desc='C5: 16x16 Heisenberg, rbm H=256, 131072 walkers (one GPU share of 1M)'),
    'c5conv': dict(kind='conv_2d', n=256, kw=dict(num_layers=5, num_filters=16, kernel_size=5, size_x=16, size_y=16),
                   size=16, nnn=False, walkers=4096, f_fwd=13.312e6,
                   desc='C5: 16x16 Heisenberg, conv_2d 5x16x5x5, 4096 walkers'),
}


def time_call(fn, reps):
  fn()
  torch.cuda.synchronize()
  times = []
  for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) * 1e-3)
  return float(np.median(times))


def run(name, reps, walkers=None, sweep_fraction=1.0):
  from cgs_vmc_b200 import _native, engine
  c = CONFIGS[name]
  n, B = c['n'], walkers or c['walkers']
  a = _native.Ansatz(c['kind'], n, **c['kw'])
  gen = torch.Generator().manual_seed(1234)
  flat = torch.randn(a.num_params, generator=gen) * (1.0 / math.sqrt(n if c['kind'] != 'conv_2d' else 25 * 16))
  a.set_params(flat)
  if c.get('chain'):
    ij = np.asarray([(i, (i + 1) % n) for i in range(n)], np.int32)
    jx, jz = np.full(n, -1.0, np.float32), np.ones(n, np.float32)
  else:
    ij, jx, jz = square_bonds(c['size'], c.get('nnn', False))
  ham = _native.Hamiltonian(ij, jx, jz, n)
  state = engine.WalkerState(B, n, seed=0xC65)
  sums = engine.EnergyGradientSums(a, B)
  steps = max(1, int(round(n * sweep_fraction)))
  state.mc_steps(a, steps)                                  # warm-up / equilibrate a little
  t_mc = time_call(lambda: state.mc_steps(a, steps), reps)
  t_eloc = time_call(lambda: a.local_energy(ham, state.packed), reps)
  t_acc = time_call(lambda: sums.accumulate(ham, state.packed), reps)
  mask, _ = ham.flip_enum(state.packed, want_flipped=False)
  n_act = float(sum(bin(int(v) & 0xffffffff).count('1') for v in mask[:256].cpu().numpy().reshape(-1))) / min(B, 256)
  line = {
      'config': c['desc'], 'walkers': B, 'n_sites': n, 'n_bonds': int(len(ij)), 'n_params': a.num_params,
      'mc_steps_per_launch': steps,
      'sampler_ms': t_mc * 1e3, 'walker_steps_per_sec': B * steps / t_mc,
      'local_energy_ms': t_eloc * 1e3, 'eloc_evals_per_sec': B / t_eloc,
      'accumulate_ms': t_acc * 1e3, 'accumulate_evals_per_sec': B / t_acc,
      'n_active_bonds_mean': n_act,
      'sampler_tflops_algorithmic': B * steps * c['f_fwd'] / t_mc / 1e12 if c['kind'] != 'rbm' else None,
      'eloc_tflops_algorithmic': B * (1 + n_act) * c['f_fwd'] / t_eloc / 1e12 if c['kind'] != 'rbm' else None,
  }
  print(json.dumps(line), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--configs', default='c1,c3,c4,c5rbm,c5conv')
  ap.add_argument('--reps', type=int, default=5)
  ap.add_argument('--walkers', type=int, default=None)
  ap.add_argument('--sweep-fraction', type=float, default=1.0,
                  help='fraction of a sweep (N steps) per sampler launch')
  args = ap.parse_args()
  for name in args.configs.split(','):
    run(name.strip(), args.reps, args.walkers, args.sweep_fraction)


if __name__ == '__main__':
  main()

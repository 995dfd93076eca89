#!/usr/bin/env python
"""Per-phase device time of the pure-RBM kernels at C2, from time stamps taken
inside the kernels (development build with -DCGSVMC_RBM2_TIMING):

  python __graft_entry__.py --timing          # builds cgs_vmc_b200/libcgsvmc_timing.so
  CGSVMC_LIBRARY=cgs_vmc_b200/libcgsvmc_timing.so python profiles/run_rbm2_phases.py

Replays the captured batch-step graph with the L2 flushed, then reads the
marks of the last replay: median / max over CTAs of each phase (clock64
cycles of thread 0's SM) and kernel begin / end on the global timer.
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cgs_vmc_b200 import _native, engine, lattices, wavefunctions   # noqa: E402

WALKER_MARKS = ['entry', 'tables+bonds loaded', 'state built (warp 0)', 'E_loc done (warp 0)',
                'gradient inputs staged (warp 0)', 'sweep done (warp 0)', 'barrier passed',
                'gradient tiles done (thread 0)', 'exit']
MC_MARKS = ['entry', 'tables + select table', 'state built (warp 0)', 'sweep done (warp 0)', 'exit']


def main():
  B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
  fused = '--split' not in sys.argv
  epoch = int(sys.argv[sys.argv.index('--epoch') + 1]) if '--epoch' in sys.argv else 0
  N, H = 36, 144
  lib = _native.load()
  ansatz = _native.Ansatz('rbm', N, num_layers=0, layer_size=H)
  gen = torch.Generator().manual_seed(1234)
  shapes = [(N, 1), (1,), (N, H), (H,)]
  ansatz.set_params(torch.cat([t.reshape(-1) for t in wavefunctions._sonnet_init(shapes, gen)]))
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(6, 6), -1.0, 1.0)
  ham = _native.Hamiltonian(ij, jx, jz, N)
  state = engine.WalkerState(B, N)
  sums = engine.EnergyGradientSums(ansatz, B)
  state.mc_steps(ansatz, 20 * N)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  if epoch:
    g = engine.GraphedEpoch(state, ansatz, ham, sums, N, epoch)     # cgsvmc_batch_steps: one persistent kernel
    run = g.replay
  elif fused:
    g = engine.GraphedBatchStep(state, ansatz, ham, sums, N)
    run = g.replay
  else:
    def run():
      sums.accumulate(ham, state.packed)
      state.mc_steps(ansatz, N)
  for _ in range(10):
    run()
    flush.zero_()
  torch.cuda.synchronize()
  marks = np.zeros((2, 160, 12, 2), dtype=np.uint64)
  rc = lib.cgsvmc_debug_rbm2_marks(ctypes.c_void_p(marks.ctypes.data))
  assert rc == 0
  sm_hz = 1.965e9
  out = {'walkers': B, 'mode': 'fused batch step (graph)' if fused else 'accumulate + mc_steps'}
  n_cta = min(148, (B + 55) // 56)
  if epoch:
    # marks of the LAST iteration of the launch: 10 = iteration start, then 2 .. 7; 8 / 9 after the loop
    m = marks[1, :n_cta].astype(np.int64)
    order = [(10, 'iteration start'), (2, 'state built (warp 0)'), (3, 'E_loc done (warp 0)'),
             (4, 'gradient inputs staged (warp 0)'), (5, 'sweep done (warp 0)'), (6, 'barrier passed'),
             (7, 'gradient tiles done (thread 0)'), (8, 'loop left'), (9, 'exit (after the grid reduction)')]
    phases = {}
    for (k0, n0), (k1, n1) in zip(order[:-1], order[1:]):
      d = (m[:, k1, 1] - m[:, k0, 1]) / sm_hz * 1e6
      phases['%s -> %s' % (n0, n1)] = {'median_us': float(np.median(d)), 'max_us': float(d.max())}
    span = float((m[:, 9, 0].max() - m[:, 0, 0].min()) * 1e-3)
    out = {'walkers': B, 'mode': 'cgsvmc_batch_steps, %d iterations in one launch (graph)' % epoch,
           'last_iteration_phases': phases, 'first_entry_to_last_exit_us (globaltimer)': span,
           'us_per_iteration': span / epoch}
    print(json.dumps(out, indent=1))
    return
  for kern, names in ((1, WALKER_MARKS), (0, MC_MARKS)):
    m = marks[kern, :n_cta].astype(np.int64)
    if m[:, 0, 0].max() == 0 or (fused and kern == 0):
      continue
    clk = m[:, :len(names), 1]
    gt = m[:, :len(names), 0]
    phases = {}
    for k in range(1, len(names)):
      d = (clk[:, k] - clk[:, k - 1]) / sm_hz * 1e6
      phases['%s -> %s' % (names[k - 1], names[k])] = {'median_us': float(np.median(d)),
                                                       'max_us': float(d.max())}
    out['walker_kernel' if kern else 'mc_kernel'] = {
        'phases': phases,
        'first_entry_to_last_exit_us (globaltimer)': float((gt[:, len(names) - 1].max() - gt[:, 0].min()) * 1e-3),
        'entry_spread_us': float((gt[:, 0].max() - gt[:, 0].min()) * 1e-3),
    }
  if not fused and marks[0, 0, 0, 0] and marks[1, 0, 0, 0]:
    out['walker_exit_to_mc_entry_us (globaltimer, includes reduce_kernel)'] = float(
        (marks[0, :n_cta, 0, 0].astype(np.int64).min() - marks[1, :n_cta, 8, 0].astype(np.int64).max()) * 1e-3)
  print(json.dumps(out, indent=1))


if __name__ == '__main__':
  main()

#!/usr/bin/env python
"""Where an fc_tc forward spends its cycles (timing build, see run_conv_tc_phases.py).

  CGSVMC_LIBRARY=cgs_vmc_b200/libcgsvmc_timing.so python profiles/run_fc_tc_phases.py [walkers]
"""
import ctypes
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
  walkers = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
  os.environ.setdefault('CGSVMC_LIBRARY', os.path.join(REPO, 'cgs_vmc_b200', 'libcgsvmc_timing.so'))
  import numpy as np
  import torch
  from cgs_vmc_b200 import _native, engine, lattices
  lib = _native.load()
  a = _native.Ansatz('fully_connected', 20, num_layers=3, layer_size=80)
  gen = torch.Generator().manual_seed(1234)
  a.set_params(torch.randn(a.num_params, generator=gen) * 0.1)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.chain_bonds(20), -1.0, 1.0)
  ham = _native.Hamiltonian(ij, jx, jz, 20)
  state = engine.WalkerState(walkers, 20, seed=0xC65)
  out = (ctypes.c_ulonglong * 6)()

  def measure(name, fn):
    fn()
    lib.cgsvmc_debug_fc_tc_phases(out)          # clear
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    lib.cgsvmc_debug_fc_tc_phases(out)
    wait, epi, n_epi, fwd, n_fwd, iss = [int(v) for v in out]
    print(json.dumps({'kernel': name, 'walkers': walkers, 'ms': e0.elapsed_time(e1), 'forwards': n_fwd,
                      'epilogues': n_epi, 'cycles_per_forward': fwd / max(1, n_fwd),
                      'wait_for_mma_cycles_per_epilogue': wait / max(1, n_epi),
                      'epilogue_cycles': epi / max(1, n_epi),
                      'issue_cycles_per_epilogue': iss / max(1, n_epi)}))

  measure('fc_mc_kernel (20 steps)', lambda: state.mc_steps(a, 20))
  measure('fc_eloc_kernel', lambda: a.local_energy(ham, state.packed))
  measure('fc_log_amp_kernel', lambda: a.log_amp(state.packed))


if __name__ == '__main__':
  main()

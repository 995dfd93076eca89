"""Device time per batch iteration of pure-RBM shapes with FEW hidden units
(6x6 Heisenberg, H = 32 / 64 / 80: the walker-kernel variants with at most 10
hidden units per lane), one graph per step and one persistent kernel per 20
steps.  Used to compare thread geometries of those variants
(CGSVMC_LIBRARY=cgs_vmc_b200/libcgsvmc_cgsvmc_rbm2_all_512.so: 512 threads, no
register cap at 64)."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
from cgs_vmc_b200 import _native, engine   # noqa: E402
from oracle import ansatz as oansatz       # noqa: E402
from oracle import lattices                # noqa: E402
from gpu_util import make_native           # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ev = lambda: torch.cuda.Event(enable_timing=True)
ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(6))
for hidden in (32, 64, 80):
  for walkers in (8192, 65536):
    spec = oansatz.AnsatzSpec('rbm', 36, num_layers=0, layer_size=hidden, size_x=6, size_y=6)
    params = oansatz.init_params(spec, seed=3, bias_scale=0.1, dtype=torch.float64)
    a = make_native(spec, oansatz.flatten(params).numpy())
    ham = _native.Hamiltonian(ij, jx, jz, 36)
    st = engine.WalkerState(walkers, 36, seed=5)
    sums = engine.EnergyGradientSums(a, walkers)
    st.mc_steps(a, 20 * 36)
    g1 = engine.GraphedBatchStep(st, a, ham, sums, 36)
    for _ in range(5):
      g1.replay(); flush.zero_()
    marks = [(ev(), ev()) for _ in range(20)]
    for x, y in marks:
      x.record(); g1.replay(); y.record(); flush.zero_()
    torch.cuda.synchronize()
    per_step = float(np.mean([x.elapsed_time(y) for x, y in marks]))
    g20 = engine.GraphedEpoch(st, a, ham, sums, 36, 20)
    for _ in range(2):
      g20.replay(); flush.zero_()
    marks = [(ev(), ev()) for _ in range(5)]
    for x, y in marks:
      x.record(); g20.replay(); y.record(); flush.zero_()
    torch.cuda.synchronize()
    per_epoch = float(np.mean([x.elapsed_time(y) for x, y in marks])) / 20
    m0, m1 = ev(), ev()
    m0.record()
    for _ in range(10):
      st.mc_steps(a, 36)
    m1.record()
    torch.cuda.synchronize()
    print(json.dumps({'hidden': hidden, 'walkers': walkers, 'ms_per_step_one_graph': per_step,
                      'ms_per_step_epoch_kernel': per_epoch, 'ms_sampler_sweep': m0.elapsed_time(m1) / 10,
                      'mean_energy': float(sums.stats[0] / sums.stats[2])}))

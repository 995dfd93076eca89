"""smoke(): one small pass of the hot path on cuda:0 checked against the
oracle (the only place in the package tree that imports oracle/, and only as
the checker)."""
import numpy as np
import torch


def run():
  from cgs_vmc_b200 import _native
  from oracle import ansatz as oansatz
  from oracle import bits, hamiltonian, lattices

  _native.require_cuda()
  torch.cuda.set_device(0)
  spec = oansatz.AnsatzSpec('rbm', 36, num_layers=0, layer_size=144, size_x=6, size_y=6)
  params = oansatz.init_params(spec, seed=1234, bias_scale=0.1, dtype=torch.float64)
  a = _native.Ansatz('rbm', 36, num_layers=0, layer_size=144)
  a.set_params(oansatz.flatten(params).float())
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(6))
  ham = _native.Hamiltonian(ij, jx, jz, 36)

  packed = _native.random_configs(256, 36, seed=0xC65)
  count = torch.zeros(1, dtype=torch.int64, device='cuda')
  a.mc_steps(packed, 36, 0xC65, accept_count=count)              # one sweep
  e, z = a.local_energy(ham, packed)
  w = torch.stack([torch.ones_like(e), e])
  s = a.weighted_grad_sum(packed, w)
  stats = _native.energy_stats(e)
  # the fused batch step (estimators + sweep in one kernel) on a copy of the walkers
  from cgs_vmc_b200 import engine
  state = engine.WalkerState(256, 36, packed=packed.clone())
  sums = engine.EnergyGradientSums(a, 256)
  e_fused = sums.batch_step(ham, state, 36).clone()
  torch.cuda.synchronize()
  assert torch.equal(e_fused, e), 'fused batch step: local energies differ from cgsvmc_local_energy'
  assert torch.allclose(sums.sums, s, rtol=1e-5, atol=1e-5), 'fused batch step: gradient sums differ'

  cfg = bits.unpack(packed.cpu().numpy().view(np.uint64), 36)
  assert np.all(cfg.sum(axis=1) == 0), 'Sz not conserved'
  cfg64 = torch.from_numpy(cfg).to(torch.float64)
  fn = lambda c: oansatz.log_amp(spec, params, c)
  zo = fn(cfg64).numpy()
  eo = hamiltonian.local_energy(cfg64, ij, jx, jz, fn).numpy()
  assert np.allclose(z.cpu().numpy(), zo, atol=2e-4, rtol=1e-5), 'log-amplitude mismatch'
  assert np.allclose(e.cpu().numpy(), eo, atol=2e-3, rtol=1e-4), 'local energy mismatch'
  assert abs(stats[0].item() - eo.sum()) < 1e-2 * (1 + abs(eo.sum()))
  assert torch.isfinite(s).all()
  print('smoke ok: accept=%d/%d  <E>/N=%.5f  |G1|=%.3f' % (
      int(count.item()), 256 * 36, eo.mean() / 36, float(s[0].norm())))

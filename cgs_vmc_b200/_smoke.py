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
  # (the fused kernel forms the sums on the tensor cores, cgsvmc_weighted_grad_sum in FP32 register tiles)
  assert float((sums.sums - s).abs().max()) <= 2e-6 * float(s.abs().max()), 'fused batch step: gradient sums differ'

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
  # the tensor-core ansaetze: fully connected (fc_tc.cu, gradient fc_tc_grad.cu)
  # and periodic convolution (conv_tc.cu), amplitudes against the oracle
  for tc_spec in (oansatz.AnsatzSpec('fully_connected', 20, num_layers=3, layer_size=80),
                  oansatz.AnsatzSpec('conv_2d', 36, num_layers=3, num_filters=16, kernel_size=3,
                                     size_x=6, size_y=6)):
    tc_params = oansatz.init_params(tc_spec, seed=7, bias_scale=0.1, dtype=torch.float64)
    t = _native.Ansatz(tc_spec.kind, tc_spec.n_sites, num_layers=tc_spec.num_layers,
                       layer_size=tc_spec.layer_size, num_filters=tc_spec.num_filters,
                       kernel_size=tc_spec.kernel_size, size_x=tc_spec.size_x, size_y=tc_spec.size_y)
    t.set_params(oansatz.flatten(tc_params).float())
    tc_cfg = bits.random_sz0_configs(tc_spec.n_sites, 64, np.random.default_rng(5))
    tc_packed = torch.from_numpy(bits.pack(tc_cfg).view(np.int64)).cuda()
    z_tc = t.log_amp(tc_packed).cpu().numpy()
    z_ref = oansatz.log_amp(tc_spec, tc_params, torch.from_numpy(tc_cfg).to(torch.float64)).numpy()
    assert np.allclose(z_tc, z_ref, atol=3e-4, rtol=3e-5), '%s: log-amplitude mismatch' % tc_spec.kind
    g_tc = t.weighted_grad_sum(tc_packed, torch.ones(1, 64, device='cuda'))
    assert torch.isfinite(g_tc).all(), '%s: gradient not finite' % tc_spec.kind
  print('smoke ok: accept=%d/%d  <E>/N=%.5f  |G1|=%.3f' % (
      int(count.item()), 256 * 36, eo.mean() / 36, float(s[0].norm())))

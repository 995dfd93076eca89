"""The oracle against the golden vectors recorded from the reference's own
Python (tests/golden/make_golden.py) -- CPU only."""
import numpy as np
import torch

from oracle import ansatz, bits, estimators, hamiltonian, sampler

F64 = torch.float64


def _params(spec, flat, dtype=F64):
  return ansatz.unflatten(spec, torch.from_numpy(flat).to(dtype))


def test_golden_files_present():
  from conftest import golden_names
  assert len(golden_names()) >= 11


def test_amplitudes(golden):
  """wavefunctions.py _build: psi = exp(z - shift).  The reference ran in
  float32; the float64 oracle must agree to float32 rounding."""
  _, spec, g = golden
  params = _params(spec, g['params_flat'])
  cfg = torch.from_numpy(g['configs']).to(F64)
  psi = ansatz.psi(spec, params, cfg, shift=float(g['shift']))
  np.testing.assert_allclose(psi.numpy(), g['psi'], rtol=2e-5)
  psi0 = ansatz.psi(spec, params, cfg, shift=-10.0)   # wavefunctions.py:209
  np.testing.assert_allclose(psi0.numpy(), g['psi_default_shift'], rtol=2e-5)
  # and the literal log(cosh) form equals the stable one here
  psi_lit = ansatz.psi(spec, params, cfg, shift=float(g['shift']),
                       literal_log_cosh=True)
  np.testing.assert_allclose(psi_lit.numpy(), psi.numpy(), rtol=1e-12)


def test_param_count(golden):
  _, spec, g = golden
  assert ansatz.num_params(spec) == g['params_flat'].size


def test_metropolis_replay(golden):
  """graph_builders.py:54-89 with the recorded uniforms: identical post-step
  configurations and acceptance count."""
  _, spec, g = golden
  params = _params(spec, g['params_flat'])
  fn = lambda c: ansatz.log_amp(spec, params, c)
  for s in range(g['mc_before'].shape[0]):
    before = torch.from_numpy(g['mc_before'][s]).to(F64)
    new, accept, _, down, up = sampler.mc_step(
        before, torch.from_numpy(g['mc_u_sites'][s]).to(F64),
        torch.from_numpy(g['mc_u_acc'][s]).to(F64), fn)
    assert np.array_equal(new.numpy().astype(np.float32), g['mc_after'][s])
    assert float(accept.sum()) == float(g['mc_accept_count'][s])
    rows = np.arange(before.shape[0])
    assert np.all(g['mc_before'][s][rows, down.numpy()] == -1)
    assert np.all(g['mc_before'][s][rows, up.numpy()] == 1)
    # the amplitude (op-order) form used by the CPU baseline agrees too
    psi_fn = lambda c: ansatz.psi(spec, params, c, shift=float(g['shift']))
    new2, count = sampler.mc_step_reference_form(
        before, torch.from_numpy(g['mc_u_sites'][s]).to(F64),
        torch.from_numpy(g['mc_u_acc'][s]).to(F64), psi_fn)
    assert np.array_equal(new2.numpy().astype(np.float32), g['mc_after'][s])
    assert float(count) == float(g['mc_accept_count'][s])


def test_flipped_configs_bit_exact(golden):
  """operators.py:154-167: swapped configurations and antiparallel mask."""
  _, spec, g = golden
  ij = g['bonds_ij']
  active, updated = bits.flip_enum_dense(g['configs'], ij)
  assert np.array_equal(updated, g['flipped_configs'])
  packed = bits.pack(g['configs'])
  mask, flipped = bits.flip_enum(packed, ij, spec.n_sites)
  nb = len(ij)
  for k in range(nb):
    assert np.array_equal(bits.unpack(flipped[:, k], spec.n_sites),
                          g['flipped_configs'][:, k])
    got = (mask[:, k >> 5] >> np.uint32(k & 31)) & np.uint32(1)
    assert np.array_equal(got.astype(bool), active[:, k])
    # active <=> the reference's off-diagonal prefactor is non-zero
    assert np.array_equal(active[:, k], g['bond_offdiag'][:, k] != 0)
  assert np.array_equal(bits.unpack(packed, spec.n_sites), g['configs'])


def test_local_energy(golden):
  """operators.py:227-271: diag, offdiag, local_value, apply_in_place."""
  _, spec, g = golden
  params = _params(spec, g['params_flat'])
  cfg = torch.from_numpy(g['configs']).to(F64)
  shift = float(g['shift'])
  psi_fn = lambda c: ansatz.psi(spec, params, c, shift=shift)
  diag, off = hamiltonian.build(cfg, g['bonds_ij'], g['bonds_jx'],
                                g['bonds_jz'], psi_fn)
  np.testing.assert_allclose(diag.numpy(), g['ham_diag'], rtol=0, atol=1e-6)
  scale = np.abs(g['bond_offdiag']).sum(axis=1)
  assert np.all(np.abs(off.numpy() - g['ham_offdiag']) <= 3e-5 * scale + 1e-30)
  e = hamiltonian.local_energy(cfg, g['bonds_ij'], g['bonds_jx'], g['bonds_jz'],
                               lambda c: ansatz.log_amp(spec, params, c))
  e_scale = scale / g['psi'] + np.abs(g['ham_diag'])
  assert np.all(np.abs(e.numpy() - g['local_energy']) <= 5e-5 * e_scale)
  aip = hamiltonian.apply_in_place(cfg, g['bonds_ij'], g['bonds_jx'],
                                   g['bonds_jz'], psi_fn)
  assert np.all(np.abs(aip.numpy() - g['apply_in_place'])
                <= 5e-5 * (scale + np.abs(g['ham_diag']) * g['psi']))


def test_energy_gradient(golden):
  """training.py:539-564 for one batch: g = G2 - mean(E) * G1."""
  _, spec, g = golden
  if 'eg_gradient' not in g:
    return
  params = _params(spec, g['params_flat'])
  cfg = torch.from_numpy(g['eg_configs']).to(F64)
  e = hamiltonian.local_energy(cfg, g['bonds_ij'], g['bonds_jx'], g['bonds_jz'],
                               lambda c: ansatz.log_amp(spec, params, c))
  acc = estimators.EnergyGradientAccumulator(ansatz.num_params(spec))
  acc.accumulate(spec, params, cfg, e)
  grad = acc.gradient().numpy()
  assert abs(acc.mean_energy - float(g['eg_mean_energy'])) < 2e-5 * (1 + abs(acc.mean_energy))
  ref = g['eg_gradient']
  # the reference gradient is a float32 difference of two O(B * |E|) sums
  s = estimators.weighted_grad_sum(spec, params, cfg,
                                   torch.stack([torch.ones_like(e), e.abs()]))
  tol = 2e-5 * (s[1].abs().numpy() + abs(acc.mean_energy) * s[0].abs().numpy()) + 1e-6
  tol = np.maximum(tol, 2e-5 * np.abs(grad).max())
  assert np.all(np.abs(grad - ref) <= tol), np.abs(grad - ref).max()
  assert np.linalg.norm(grad - ref) <= 1e-4 * np.linalg.norm(ref) + 1e-6


def test_swo_loss_and_gradient(golden):
  """training.py:166-175."""
  _, spec, g = golden
  if 'swo_gradient' not in g:
    return
  params = _params(spec, g['params_flat'])
  tparams = _params(spec, g['swo_target_params_flat'])
  cfg = torch.from_numpy(g['swo_configs']).to(F64)
  psi_t = ansatz.psi(spec, tparams, cfg, shift=float(g['swo_target_shift']))
  loss, grad = estimators.swo_loss_and_grad(spec, params, cfg, psi_t,
                                            shift=float(g['shift']))
  assert abs(float(loss) - float(g['swo_loss'])) <= 1e-4 * abs(float(loss)) + 1e-6
  ref = g['swo_gradient']
  assert np.linalg.norm(grad.numpy() - ref) <= 2e-4 * np.linalg.norm(ref) + 1e-6

"""Known answers that pin sign / coupling conventions of the oracle
(SURVEY.md appendix C): exact-diagonalisation energies, the zero-variance
identity E_loc(s) == E0 for the exact ground vector, and uniform sampling for
a constant amplitude -- CPU only."""
import numpy as np
import pytest
import torch

from oracle import bits, ed, hamiltonian, lattices, sampler

E0 = {
    ('chain', 4): -2.0,
    ('chain', 8): -3.6510934089,
    ('chain', 12): -5.3873909174,
    ('chain', 16): -7.1422963606,
    ('square', 4): -11.2284832084,
    ('j1j2', 4): -8.4579233514,
}


def _system(kind, size, jx_sign=-1.0):
  if kind == 'chain':
    ij, jx, jz = lattices.heisenberg_couplings(lattices.chain_bonds(size), jx_sign, 1.0)
    return size, ij, jx, jz
  if kind == 'square':
    ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(size), jx_sign, 1.0)
    return size * size, ij, jx, jz
  ij, jx, jz = lattices.j1j2_couplings(size, 0.5, marshall=jx_sign < 0)
  return size * size, ij, jx, jz


@pytest.mark.parametrize('kind,size', [k for k in E0 if k != ('chain', 16)])
def test_ed_energies(kind, size):
  for sign in (-1.0, 1.0):     # invariant under the Marshall rotation
    n, ij, jx, jz = _system(kind, size, sign)
    e0, _, _ = ed.ground_state(n, ij, jx, jz)
    assert abs(e0 - E0[(kind, size)]) < 1e-8


def test_zero_variance_chain8():
  n, ij, jx, jz = _system('chain', 8)
  e0, basis, vec = ed.ground_state(n, ij, jx, jz)
  psi_fn = ed.lookup_amplitude(basis, vec)
  cfg = torch.from_numpy(bits.unpack(
      basis.astype(np.uint64)[:, None], n, dtype=np.float64))
  diag, off = hamiltonian.build(cfg, ij, jx, jz, psi_fn)
  e_loc = diag + off / psi_fn(cfg)
  assert cfg.shape[0] == 70
  np.testing.assert_allclose(e_loc.numpy(), e0, atol=1e-10)
  # and the log form on |psi| (Marshall-rotated ground state is sign-free)
  assert np.all(vec > 0) or np.all(vec < 0)
  log_fn = lambda c: torch.log(psi_fn(c).abs())
  e2 = hamiltonian.local_energy(cfg, ij, jx, jz, log_fn)
  np.testing.assert_allclose(e2.numpy(), e0, atol=1e-10)


def test_constant_amplitude_samples_uniformly():
  """Detailed balance of the exchange move (graph_builders.py:59-88): with a
  constant psi every move is accepted with probability P(1 > sqrt(u)) = 1 and
  the chain is uniform on the 70 Sz = 0 states of N = 8."""
  n, batch, steps = 8, 512, 60
  rng = np.random.default_rng(7)
  cfg = torch.from_numpy(bits.random_sz0_configs(n, batch, rng)).to(torch.float64)
  fn = lambda c: torch.zeros(c.shape[0], dtype=torch.float64)
  counts = {}
  for s in range(steps):
    u_sites = torch.from_numpy(rng.random((batch, n)))
    u_acc = torch.from_numpy(rng.random(batch))
    cfg, accept, _, _, _ = sampler.mc_step(cfg, u_sites, u_acc, fn)
    assert torch.all(cfg.sum(dim=1) == 0)
    if s >= 20:
      for key in bits.pack(cfg.numpy())[:, 0]:
        counts[int(key)] = counts.get(int(key), 0) + 1
  assert len(counts) == 70
  obs = np.array(list(counts.values()), dtype=np.float64)
  exp = obs.sum() / 70.0
  chi2 = ((obs - exp) ** 2 / exp).sum()
  assert chi2 < 130.0          # 69 dof; P(chi2 > 130) ~ 1e-5


def test_sampled_energy_matches_ed_chain8():
  """Sampling |psi_0|^2 with the oracle sampler reproduces E0 exactly
  (zero variance) and conserves Sz."""
  n, ij, jx, jz = _system('chain', 8)
  e0, basis, vec = ed.ground_state(n, ij, jx, jz)
  psi_fn = ed.lookup_amplitude(basis, vec)
  log_fn = lambda c: torch.log(psi_fn(c).abs())
  rng = np.random.default_rng(3)
  cfg = torch.from_numpy(bits.random_sz0_configs(n, 64, rng)).to(torch.float64)
  for _ in range(16):
    cfg, _, _, _, _ = sampler.mc_step(
        cfg, torch.from_numpy(rng.random((64, n))), torch.from_numpy(rng.random(64)), log_fn)
  e = hamiltonian.local_energy(cfg, ij, jx, jz, log_fn)
  np.testing.assert_allclose(e.numpy(), e0, atol=1e-9)


def test_exact_expectation_and_acceptance_known_answers():
  """oracle/ed.py helpers behind the evaluator parity test: the exact
  expectation of the ground vector is E0, and a constant amplitude is accepted
  with probability one."""
  import numpy as np
  from oracle import ed, lattices
  n = 8
  bonds = lattices.chain_bonds(n)
  ij, jx, jz = lattices.heisenberg_couplings(bonds)
  e0, basis, vec = ed.ground_state(n, ij, jx, jz)
  psi_fn = ed.lookup_amplitude(basis, vec)
  assert abs(ed.exact_expectation(n, ij, jx, jz, lambda c: psi_fn(c).numpy()) - e0) < 1e-10
  assert abs(ed.exact_acceptance_rate(n, lambda c: np.ones(len(c))) - 1.0) < 1e-12
  rate = ed.exact_acceptance_rate(n, lambda c: psi_fn(c).numpy())
  assert 0.0 < rate < 1.0


def test_periodic_padding_against_reference_layers():
  """oracle.ansatz._pad_periodic_1d / _2d + _conv_valid against the outputs of
  the reference's own layers.Conv1dPeriodic / Conv2dPeriodic modules
  (tests/golden/make_golden_layers.py; layers.py:24-160), odd and even kernels."""
  import os
  import numpy as np
  import torch
  from oracle import ansatz as oansatz
  g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'layers_periodic.npz'))
  n = 0
  while 'case%d_meta' % n in g:
    rank, c_in, c_out, k = (int(v) for v in g['case%d_meta' % n])
    x, w, b = (torch.from_numpy(g['case%d_%s' % (n, key)]).double() for key in 'xwb')
    pad = oansatz._pad_periodic_1d if rank == 1 else oansatz._pad_periodic_2d
    y = oansatz._conv_valid(pad(x, k), w, b).numpy()
    np.testing.assert_allclose(y, g['case%d_y' % n], rtol=2e-5, atol=2e-5)
    n += 1
  assert n == 8

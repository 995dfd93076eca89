"""Signed output activations and sum / diff / prod composites on the CUDA path
(amplitude-agnostic route: cgsvmc_log_amp of the parts + cgsvmc_flip_enum +
cgsvmc_local_energy_from_amps + cgsvmc_propose_exchange / cgsvmc_accept_exchange)
against vectors recorded from the reference (tests/golden/cmp_*.npz) and the
float64 oracle."""
import numpy as np
import pytest
import torch

from conftest import cmp_golden_names, load_cmp_golden
from oracle import bits, composite, ed, lattices

pytestmark = pytest.mark.gpu
F64 = torch.float64


def _leaves(wf):
  return wf.leaves()


def _build(name):
  """The wavefunction of a golden case through the public API, with the
  recorded parameters and shifts."""
  from cgs_vmc_b200 import utils, wavefunctions
  kind, oleaves, hp_over, g = load_cmp_golden(name)
  hp_over = {k: (tuple(v) if isinstance(v, list) else v) for k, v in hp_over.items()}
  hp = utils.create_hparams(num_sites=8, batch_size=int(g['configs'].shape[0]), **hp_over)
  wf = wavefunctions.build_wavefunction(hp)
  wf.connect(8)
  off = 0
  for k, leaf in enumerate(_leaves(wf)):
    n = int(g['leaf_sizes'][k])
    assert leaf.native().num_params == n
    leaf.native().set_params(torch.from_numpy(g['leaf_params_flat'][off:off + n]))
    off += n
    if leaf.fast_path:
      leaf._exp_norm_shift = float(g['leaf_shifts'][k])
  return wf, hp, kind, oleaves, g


@pytest.mark.parametrize('name', cmp_golden_names())
def test_amplitudes_and_local_energy_golden(name):
  from cgs_vmc_b200 import operators
  wf, hp, kind, oleaves, g = _build(name)
  cfg = torch.from_numpy(g['configs']).cuda()
  psi = wf(cfg).cpu().numpy()
  np.testing.assert_allclose(psi, g['psi'], rtol=1e-4, atol=1e-5)
  assert (np.sign(psi) == np.sign(g['psi'])).all()
  ham = operators.HeisenbergHamiltonian([tuple(b) for b in g['bonds_ij']], -1.0, 1.0)
  e = ham.local_value(wf, cfg).cpu().numpy()
  scale = 1.0 + np.abs(g['local_energy'])
  assert np.all(np.abs(e - g['local_energy']) <= 5e-4 * scale * (1.0 + 1e-2 / np.abs(g['psi'])))
  hpsi = ham.apply_in_place(wf, cfg).cpu().numpy()
  np.testing.assert_allclose(hpsi, g['apply_in_place'], rtol=5e-4, atol=5e-5)
  diag, off = ham.build(wf, cfg)
  np.testing.assert_allclose((diag.cpu().numpy() * psi + off.cpu().numpy()), g['apply_in_place'],
                             rtol=5e-4, atol=5e-5)


@pytest.mark.parametrize('name', cmp_golden_names())
def test_energy_gradient_golden(name):
  from cgs_vmc_b200 import graph_builders, operators, training
  from cgs_vmc_b200.session import Session
  wf, hp, kind, oleaves, g = _build(name)
  ham = operators.HeisenbergHamiltonian([tuple(b) for b in g['bonds_ij']], -1.0, 1.0)
  opt = training.GROUND_STATE_OPTIMIZERS['EnergyGradient']()
  shared = {}
  ops = opt.build_opt_ops(wavefunction=wf, hamiltonian=ham, hparams=hp, shared_resources=shared)
  shared[graph_builders.ResourceName.CONFIGS].assign(torch.from_numpy(g['eg_configs']))
  s = Session()
  s.run(ops.reset_gradients)
  s.run(ops.accumulate_gradients)
  grad = opt.sums.gradient().cpu().numpy()
  ref = g['eg_gradient']
  assert grad.shape == ref.shape
  np.testing.assert_allclose(grad, ref, rtol=0, atol=2e-3 * np.abs(ref).max())
  assert abs(s.run(ops.metrics) - float(g['eg_mean_energy'])) <= 5e-4 * (1 + abs(float(g['eg_mean_energy'])))
  before = [leaf.native().params.clone() for leaf in wf.leaves()]
  s.run(ops.apply_gradients)                       # every leaf gets its slice of the update
  assert all(not torch.equal(b, leaf.native().params) for b, leaf in zip(before, wf.leaves()))


def test_generic_sampler_reproduces_fused_trajectories():
  """(rbm * 1.0) takes the amplitude-agnostic sampler; same Philox convention
  as the fused kernel, so the walkers follow the same trajectories up to
  float32 differences in the acceptance ratio."""
  from cgs_vmc_b200 import graph_builders, utils, wavefunctions
  from cgs_vmc_b200.session import Session
  hp = utils.create_hparams(wavefunction_type='rbm', num_sites=16, num_fc_layers=0, fc_layer_size=12)
  fused = wavefunctions.build_wavefunction(hp).seed(3)
  generic = wavefunctions.build_wavefunction(hp).seed(3) * 1.0
  assert fused.fast_path and not generic.fast_path
  s = Session()
  out = []
  for wf in (fused, generic):
    shared = {}
    configs = graph_builders.get_configs(shared, 512, 16, seed=7)
    mc_step, acc = graph_builders.get_monte_carlo_sampling(shared, configs, wf)
    s.run(mc_step, n_steps=24)
    out.append((configs.value().cpu().numpy(), s.run(acc)))
  same = (out[0][0] == out[1][0]).all(axis=1).mean()
  assert same > 0.98, same
  assert abs(out[0][1] - out[1][1]) <= 0.02 * out[0][1] + 5
  assert np.all(out[1][0].sum(axis=1) == 0)


def test_generic_sampler_distribution_signed_wavefunction():
  """|psi|^2 sampling of a sign-changing composite on 8 sites: chi^2 against
  the exact distribution over the 70 Sz = 0 states."""
  from cgs_vmc_b200 import graph_builders
  from cgs_vmc_b200.session import Session
  wf, hp, kind, oleaves, g = _build('cmp_sum_rbm_fc_tanh')
  basis = ed.sz0_basis(8)
  all_cfg = torch.from_numpy(bits.unpack(basis.astype(np.uint64)[:, None], 8, np.float64))
  p = composite.psi(kind, oleaves, all_cfg).numpy() ** 2
  p /= p.sum()
  shared = {}
  B = 8192
  configs = graph_builders.get_configs(shared, B, 8, seed=11)
  mc_step, _ = graph_builders.get_monte_carlo_sampling(shared, configs, wf)
  s = Session()
  s.run(mc_step, n_steps=8 * 25)
  counts = np.zeros(len(basis))
  index = {int(b): k for k, b in enumerate(basis)}
  for _ in range(6):
    s.run(mc_step, n_steps=8 * 3)
    packed = configs.packed.cpu().numpy().view(np.uint64)[:, 0]
    for v, c in zip(*np.unique(packed, return_counts=True)):
      counts[index[int(v)]] += c
  expected = p * counts.sum()
  keep = expected > 5
  chi2 = float((((counts - expected) ** 2) / np.maximum(expected, 1e-9))[keep].sum())
  # correlated samples: allow a generous factor over dof
  assert chi2 < 6.0 * keep.sum(), (chi2, keep.sum())


def test_composite_energy_training_lowers_energy():
  """EnergyGradientOptimizer end to end on sum(rbm, fc tanh), chain of 8:
  the energy drops towards exact diagonalisation (-3.6511)."""
  from cgs_vmc_b200 import graph_builders, operators, training, utils, wavefunctions
  from cgs_vmc_b200.session import Session
  graph_builders.reset_num_epochs()
  hp = utils.create_hparams(wavefunction_type='sum', composite_wavefunction_types=('rbm', 'fully_connected'),
                            composite_output_activations=('exp', 'tanh'), num_sites=8, num_fc_layers=1,
                            fc_layer_size=8, batch_size=1024, num_batches_per_epoch=2,
                            num_equilibration_sweeps=2, learning_rates=[0.02, 0.01, 0.005, 0.002],
                            learning_rate_stops=[20, 40, 60])
  wf = wavefunctions.build_wavefunction(hp)
  wf.connect(8)
  for k, leaf in enumerate(wf.leaves()):
    gen = torch.Generator().manual_seed(40 + k)
    leaf.native().set_params(0.3 * torch.randn(leaf.native().num_params, generator=gen))
  wf.leaves()[0]._exp_norm_shift = 0.0
  ham = operators.HeisenbergHamiltonian(lattices.chain_bonds(8), -1.0, 1.0)
  opt = training.GROUND_STATE_OPTIMIZERS['EnergyGradient']()
  ops = opt.build_opt_ops(wavefunction=wf, hamiltonian=ham, hparams=hp, shared_resources={})
  s = Session()
  energies = [opt.run_optimization_epoch(ops, s, hp) for _ in range(40)]
  e0, _, _ = ed.ground_state(8, *lattices.heisenberg_couplings(lattices.chain_bonds(8)))
  assert np.all(np.isfinite(energies))
  assert np.mean(energies[-5:]) < np.mean(energies[:5]) - 0.3, (energies[:5], energies[-5:])
  assert np.mean(energies[-5:]) > e0 - 0.2


def test_checkpoint_roundtrip_of_a_composite(tmp_path):
  """checkpoint.Saver over the leaves of a composite: variables and the
  per-leaf exp_norm_shift survive, so psi is reproduced exactly."""
  from cgs_vmc_b200 import checkpoint
  wf, hp, kind, oleaves, g = _build('cmp_sum_rbm_fc_tanh')
  cfg = torch.from_numpy(g['configs']).cuda()
  psi = wf(cfg).clone()
  saver = checkpoint.Saver(wf)
  path = saver.save(None, str(tmp_path / 'model_prior_0_epochs'))
  wf2, _, _, _, _ = _build('cmp_sum_rbm_fc_tanh')
  for leaf in wf2.leaves():
    with torch.no_grad():
      leaf.native().params.mul_(0.5)
    if leaf.fast_path:
      leaf._exp_norm_shift = -10.0
  assert not torch.allclose(wf2(cfg), psi)
  checkpoint.Saver(wf2).restore(None, checkpoint.latest_checkpoint(str(tmp_path)))
  assert path.endswith('.pt') and torch.equal(wf2(cfg), psi)


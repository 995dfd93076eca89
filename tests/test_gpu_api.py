"""-m gpu tests of the reference-shaped Python API (wavefunctions, operators,
graph_builders, training, evaluation) and of the three drivers."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import ansatz as oansatz
from oracle import bits, ed, estimators, hamiltonian, lattices

pytestmark = pytest.mark.gpu
F64 = torch.float64
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _fresh_epoch_counter():
  from cgs_vmc_b200 import graph_builders
  graph_builders.reset_num_epochs()
  yield


def _oracle_view(wf, kind, hp):
  """(spec, float64 params) of a built API wavefunction."""
  if kind in ('fully_connected', 'rbm'):
    spec = oansatz.AnsatzSpec(kind, hp.num_sites, num_layers=hp.num_fc_layers,
                              layer_size=hp.fc_layer_size, nonlinearity=hp.nonlinearity)
  else:
    spec = oansatz.AnsatzSpec(kind, hp.num_sites, num_layers=hp.num_conv_layers,
                              num_filters=hp.num_conv_filters, kernel_size=hp.kernel_size,
                              size_x=hp.size_x, size_y=hp.size_y, nonlinearity=hp.nonlinearity)
  params = [v.detach().cpu().to(F64) for v in wf.get_trainable_variables()]
  assert [tuple(p.shape) for p in params] == [tuple(s) for _, s in oansatz.param_shapes(spec)]
  return spec, params


@pytest.mark.parametrize('kind,overrides', [
    ('fully_connected', dict(num_sites=20)),
    ('rbm', dict(num_sites=36, num_fc_layers=0, fc_layer_size=144)),
    ('conv_1d', dict(num_sites=12, num_conv_layers=2, num_conv_filters=4, kernel_size=3)),
    ('conv_2d', dict(num_sites=36, size_x=6, size_y=6, num_conv_layers=3, num_conv_filters=8, kernel_size=3)),
])
def test_wavefunction_and_operator_api(kind, overrides):
  from cgs_vmc_b200 import operators, utils, wavefunctions
  hp = utils.create_hparams(wavefunction_type=kind, **overrides)
  wf = wavefunctions.build_wavefunction(hp).seed(3)
  n = hp.num_sites
  cfg = bits.random_sz0_configs(n, 40, np.random.default_rng(1))
  inputs = torch.from_numpy(cfg).cuda()
  psi = wf(inputs)                                           # Wavefunction.__call__
  spec, params = _oracle_view(wf, kind, hp)
  ref = oansatz.psi(spec, params, torch.from_numpy(cfg).to(F64), shift=-10.0)
  np.testing.assert_allclose(psi.cpu().numpy(), ref.numpy(), rtol=1e-4)
  with pytest.raises(ValueError):                            # wrong number of sites
    wf(torch.ones(3, n + 1).cuda())
  bonds = lattices.chain_bonds(n)
  ham = operators.HeisenbergHamiltonian(bonds, -1.0, 1.0)
  e = ham.local_value(wf, inputs)
  eo = hamiltonian.local_energy(torch.from_numpy(cfg).to(F64), bonds, [-1.0] * n, [1.0] * n,
                                lambda c: oansatz.log_amp(spec, params, c))
  np.testing.assert_allclose(e.cpu().numpy(), eo.numpy(), rtol=2e-4, atol=2e-4)
  diag, off = ham.build(wf, inputs)
  aip = ham.apply_in_place(wf, inputs)
  torch.testing.assert_close(diag * psi + off, aip, rtol=1e-4, atol=1e-30)
  torch.testing.assert_close(diag + off / psi, e, rtol=1e-4, atol=1e-4)
  # a Hamiltonian is the sum of its bonds (operators.py:241-247)
  total = sum(operators.HeisenbergBond(b, -1.0, 1.0).local_value(wf, inputs) for b in bonds)
  torch.testing.assert_close(total, e, rtol=1e-4, atol=1e-4)
  with pytest.raises(NotImplementedError):
    ham.apply(wf)


def test_graph_builders_api():
  from cgs_vmc_b200 import graph_builders, utils, wavefunctions
  from cgs_vmc_b200.session import Session
  shared = {}
  configs = graph_builders.get_configs(shared, 64, 12)
  assert graph_builders.get_configs(shared, 64, 12) is configs
  with pytest.raises(ValueError, match='does not match'):    # graph_builders.py:117-118
    graph_builders.get_configs(shared, 32, 12)
  assert configs.value().shape == (64, 12)
  assert torch.all(configs.value().sum(dim=1) == 0)
  wf = wavefunctions.build_wavefunction(utils.create_hparams(
      wavefunction_type='rbm', num_sites=12, num_fc_layers=0, fc_layer_size=8)).seed(1)
  mc_step, acc = graph_builders.get_monte_carlo_sampling(shared, configs, wf)
  assert graph_builders.get_monte_carlo_sampling(shared, configs, wf)[0] is mc_step
  s = Session()
  before = configs.value().clone()
  s.run(mc_step)                                             # one step, like the reference
  changed = (configs.value() != before).any(dim=1).sum().item()
  assert changed == s.run(acc)                               # accepted moves of that step
  assert 0 < changed <= 64
  s.run(mc_step, n_steps=24)
  assert torch.all(configs.value().sum(dim=1) == 0)
  r = utils.random_configurations(12, 5, seed=3)
  assert r.shape == (5, 12) and torch.all(r.sum(dim=1) == 0)


def test_energy_gradient_accumulators_match_oracle():
  """Two accumulate batches of EnergyGradientOptimizer against the oracle's
  restatement of training.py:550-564 on the same configurations."""
  from cgs_vmc_b200 import operators, training, utils, wavefunctions
  from cgs_vmc_b200.session import Session
  hp = utils.create_hparams(wavefunction_type='rbm', num_sites=16, size_x=4, size_y=4,
                            num_fc_layers=0, fc_layer_size=24, batch_size=96)
  wf = wavefunctions.build_wavefunction(hp).seed(11)
  ij, jx, jz = lattices.j1j2_couplings(4, 0.5)
  ham = operators.HeisenbergHamiltonian(ij.tolist(), jx, jz)       # per-bond couplings
  opt = training.EnergyGradientOptimizer()
  shared = {}
  ops = opt.build_opt_ops(wavefunction=wf, hamiltonian=ham, hparams=hp, shared_resources=shared)
  s = Session()
  spec, params = _oracle_view(wf, 'rbm', hp)
  acc = estimators.EnergyGradientAccumulator(oansatz.num_params(spec))
  fn = lambda c: oansatz.log_amp(spec, params, c)
  s.run(ops.reset_gradients)
  from cgs_vmc_b200 import graph_builders
  configs = shared[graph_builders.ResourceName.CONFIGS]
  for _ in range(2):
    cfg64 = configs.value().cpu().to(F64)
    acc.accumulate(spec, params, cfg64, hamiltonian.local_energy(cfg64, ij, jx, jz, fn))
    s.run(ops.accumulate_gradients)
    s.run(ops.mc_step, n_steps=16)
  assert abs(s.run(ops.metrics) - acc.mean_energy) < 1e-4
  grad = opt.sums.gradient().cpu().numpy()
  ref = acc.gradient().numpy()
  assert np.linalg.norm(grad - ref) <= 5e-4 * np.linalg.norm(ref) + 1e-4


def test_energy_gradient_training_reaches_ed_energy():
  """run_optimization_epoch end to end on the 8-site chain: the variational
  energy converges to within 2% of exact diagonalisation (-3.6511)."""
  from cgs_vmc_b200 import operators, training, utils, wavefunctions
  from cgs_vmc_b200.session import Session
  hp = utils.create_hparams(wavefunction_type='rbm', num_sites=8, num_fc_layers=0, fc_layer_size=16,
                            batch_size=1024, num_batches_per_epoch=4, num_equilibration_sweeps=5,
                            learning_rates=[0.02, 0.005, 0.002, 0.001], learning_rate_stops=[60, 100, 140])
  wf = wavefunctions.build_wavefunction(hp).seed(2)
  ham = operators.HeisenbergHamiltonian(lattices.chain_bonds(8), -1.0, 1.0)
  opt = training.GROUND_STATE_OPTIMIZERS['EnergyGradient']()
  ops = opt.build_opt_ops(wavefunction=wf, hamiltonian=ham, hparams=hp, shared_resources={})
  s = Session()
  energies = [opt.run_optimization_epoch(ops, s, hp) for _ in range(120)]
  e0, _, _ = ed.ground_state(8, *lattices.heisenberg_couplings(lattices.chain_bonds(8)))
  assert energies[0] > np.mean(energies[-10:])
  assert abs(np.mean(energies[-10:]) - e0) < 0.02 * abs(e0), (energies[0], energies[-10:], e0)
  assert np.mean(energies[-10:]) > e0 - 0.05          # variational within MC noise


@pytest.mark.parametrize('kind,overrides', [
    ('rbm', dict(num_fc_layers=0, fc_layer_size=24)),                       # one persistent kernel per epoch
    ('fully_connected', dict(num_fc_layers=2, fc_layer_size=20)),          # the launches of every iteration
])
def test_energy_gradient_epoch_launch_equals_per_batch_ops(kind, overrides):
  """EnergyGradientOptimizer.run_optimization_epoch (training.py:589-623) with
  the inner loop as one cgsvmc_batch_steps call, as one captured
  cgsvmc_batch_step per batch, and as the reference's separate
  accumulate_gradients / mc_step ops: same walkers, same energies, same
  parameters (float32 summation order of the gradient sums aside)."""
  from cgs_vmc_b200 import graph_builders, operators, training, utils, wavefunctions
  from cgs_vmc_b200.session import Session
  results = []
  for mode in ('epoch', 'batch', 'ops'):
    graph_builders.reset_num_epochs()
    hp = utils.create_hparams(wavefunction_type=kind, num_sites=16, size_x=4, size_y=4, batch_size=640,
                              num_batches_per_epoch=5, num_equilibration_sweeps=1,
                              learning_rates=[0.01] * 4, **overrides)
    wf = wavefunctions.build_wavefunction(hp).seed(5)
    ij, jx, jz = lattices.j1j2_couplings(4, 0.5)
    ham = operators.HeisenbergHamiltonian(ij.tolist(), jx, jz)
    opt = training.EnergyGradientOptimizer()
    opt.use_cuda_graph = mode != 'ops'
    shared = {}
    ops = opt.build_opt_ops(wavefunction=wf, hamiltonian=ham, hparams=hp, shared_resources=shared)
    if mode == 'batch':
      opt._epoch_steps = None
    assert (opt._epoch_steps is not None) == (mode == 'epoch')
    s = Session()
    energies = [opt.run_optimization_epoch(ops, s, hp, e) for e in range(3)]
    configs = shared[graph_builders.ResourceName.CONFIGS]
    results.append((wf.flat_parameters.clone(), configs.packed.clone(), energies, configs.state.step))
  for other in results[1:]:
    assert other[3] == results[0][3]
    assert torch.equal(other[1], results[0][1])
    np.testing.assert_allclose(other[2], results[0][2], rtol=1e-5)
    # the constant a0 (rbm) only rescales psi: its energy gradient <E> - <E> is
    # pure rounding noise that Adam normalises to steps of +-lr -- left out
    keep = torch.ones_like(other[0], dtype=torch.bool)
    if kind == 'rbm':
      keep[16] = False
    torch.testing.assert_close(other[0][keep], results[0][0][keep], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize('rank,shape,c_in,c_out,k', [
    (1, (7,), 1, 5, 3), (1, (10,), 3, 16, 4), (1, (12,), 2, 3, 5), (1, (9,), 4, 6, 6), (1, (5,), 1, 2, 1),
    (2, (4, 6), 1, 5, 2), (2, (10, 10), 1, 16, 5), (2, (6, 6), 16, 16, 3), (2, (5, 7), 3, 4, 4), (2, (3, 3), 2, 2, 5),
])
def test_periodic_conv_layers_standalone(rank, shape, c_in, c_out, k):
  """layers.Conv1dPeriodic / Conv2dPeriodic called by themselves
  (cgsvmc_conv_periodic) against the oracle's restatement of layers.py:51-80 /
  117-160 -- wrap padding (with the 1-D / 2-D asymmetry of even kernels, also
  for kernels wider than the lattice) + VALID cross-correlation + bias -- which
  the network goldens recorded from the reference pin."""
  from cgs_vmc_b200 import layers
  g = torch.Generator().manual_seed(100 * rank + k)
  x = torch.randn((11,) + shape + (c_in,), generator=g)
  cls = layers.Conv1dPeriodic if rank == 1 else layers.Conv2dPeriodic
  layer = cls(c_out, k).initialize(c_in, generator=g)
  layer.b = torch.randn(c_out, generator=g).cuda()
  got = layer(x.cuda()).cpu().double()
  pad = oansatz._pad_periodic_1d if rank == 1 else oansatz._pad_periodic_2d
  ref = oansatz._conv_valid(pad(x.double(), k), layer.w.cpu().double(), layer.b.cpu().double())
  assert got.shape == ref.shape == (11,) + shape + (c_out,)
  torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)
  assert layer.pad_sizes() == ((k - 1) // 2, (k - 1) // 2) if k % 2 else True
  with pytest.raises(ValueError):
    layer(x.cuda()[0])                       # wrong rank
  lazy = cls(c_out, k)                       # Sonnet-style lazy initialisation on first call
  out = lazy(x.cuda())
  assert out.shape == got.shape and lazy.w.shape == layer.w.shape and float(lazy.b.abs().max()) == 0.0


def test_periodic_conv_layers_reference_golden():
  """cgsvmc_conv_periodic through layers.Conv1dPeriodic / Conv2dPeriodic against
  the outputs recorded from the reference's own modules on the same inputs and
  variables (tests/golden/make_golden_layers.py)."""
  from cgs_vmc_b200 import layers
  g = np.load(os.path.join(REPO, 'tests', 'golden', 'layers_periodic.npz'))
  n = 0
  while 'case%d_meta' % n in g:
    rank, c_in, c_out, k = (int(v) for v in g['case%d_meta' % n])
    layer = (layers.Conv1dPeriodic if rank == 1 else layers.Conv2dPeriodic)(c_out, k)
    layer.w = torch.from_numpy(g['case%d_w' % n]).cuda()
    layer.b = torch.from_numpy(g['case%d_b' % n]).cuda()
    y = layer(torch.from_numpy(g['case%d_x' % n]).cuda()).cpu().numpy()
    np.testing.assert_allclose(y, g['case%d_y' % n], rtol=2e-5, atol=2e-5)
    n += 1
  assert n == 8


def test_supervised_training_reduces_loss():
  from cgs_vmc_b200 import training, utils, wavefunctions
  from cgs_vmc_b200.session import Session
  hp = utils.create_hparams(wavefunction_type='rbm', num_sites=16, num_fc_layers=0, fc_layer_size=12,
                            batch_size=2048, num_batches_per_epoch=10, learning_rates=[0.01, 0.01, 0.01, 0.01])
  target = wavefunctions.build_wavefunction(hp).seed(7)
  trainee = wavefunctions.build_wavefunction(hp).seed(8)
  target.native(16)
  # scale so that psi_target * sqrt(2^N) is comparable to psi (SWO is not scale free)
  target._exp_norm_shift += 0.5 * 16 * np.log(2.0)
  opt = training.SUPERVISED_OPTIMIZERS['SWO']()
  ops = opt.build_opt_ops(wavefunction=trainee, target_wavefunction=target, hparams=hp,
                          shared_resources={})
  s = Session()
  first = s.run(ops.metrics)
  for epoch in range(30):
    opt.run_optimization_epoch(ops, s, hp, epoch)
  last = s.run(ops.metrics)
  assert last < 0.2 * first, (first, last)


def test_monte_carlo_operator_evaluator():
  """MonteCarloOperatorEvaluator (evaluation.py:95-152) on the 12-site chain:
  the mean of the sampled values agrees with the exact <psi|H|psi>/<psi|psi>
  of the same parameters within Monte-Carlo error bars, and the acceptance
  rate the reference computes and discards (evaluation.py:145-151) agrees with
  the exact stationary acceptance probability of the exchange sampler."""
  from cgs_vmc_b200 import evaluation, operators, utils, wavefunctions
  from cgs_vmc_b200.session import Session
  n = 12
  hp = utils.create_hparams(wavefunction_type='fully_connected', num_sites=n, num_fc_layers=1,
                            fc_layer_size=8, batch_size=4096, num_equilibration_sweeps=20,
                            num_monte_carlo_sweeps=2, num_evaluation_samples=40)
  wf = wavefunctions.build_wavefunction(hp).seed(4)
  bonds = lattices.chain_bonds(n)
  ham = operators.HeisenbergHamiltonian(bonds, -1.0, 1.0)
  ev = evaluation.MonteCarloOperatorEvaluator()
  ops = ev.build_eval_ops(wavefunction=wf, operator=ham, hparams=hp, shared_resources={})
  session = Session()
  values = ev.run_evaluation(ops, session, hp, epoch_num=0)
  assert len(values) == hp.num_evaluation_samples
  spec, params = _oracle_view(wf, 'fully_connected', hp)
  psi_fn = lambda c: oansatz.psi(spec, params, torch.from_numpy(np.asarray(c)).to(F64), shift=-10.0).numpy()
  exact = ed.exact_expectation(n, bonds, [-1.0] * n, [1.0] * n, psi_fn)
  # each value is the mean over 4096 walkers; successive samples are 2 sweeps
  # apart, so their spread gives the error bar of the grand mean (x2 for the
  # residual autocorrelation)
  values = np.asarray(values)
  err = 2.0 * values.std(ddof=1) / np.sqrt(len(values))
  assert abs(values.mean() - exact) < 4.0 * err + 1e-4, (values.mean(), exact, err)
  assert err < 0.01 * abs(exact)
  # acceptance: one more sweep group, counted by the op the evaluator exposes
  steps = 5 * n
  session.run(ops.mc_step, n_steps=steps)
  rate = session.run(ops.acceptance_rate) / (steps * hp.batch_size)
  exact_rate = ed.exact_acceptance_rate(n, psi_fn)
  assert abs(rate - exact_rate) < 0.01, (rate, exact_rate)


def test_update_norm_on_device_amplitudes():
  """wavefunctions.py:261-288 through the API: the shift moves by log(max psi)
  - log(max_value) iff the batch maximum exceeds max_value, and psi of the
  same configurations drops accordingly."""
  from cgs_vmc_b200 import graph_builders, utils, wavefunctions
  hp = utils.create_hparams(wavefunction_type='rbm', num_sites=16, num_fc_layers=0, fc_layer_size=12)
  wf = wavefunctions.build_wavefunction(hp).seed(5)
  configs = graph_builders.get_configs({}, 128, 16)
  psi0 = wf(configs).clone()
  assert wf._exp_norm_shift == -10.0
  top = float(psi0.max())
  assert wf.update_norm(lambda: wf(configs), max_value=2.0 * top)() == -10.0        # below: unchanged
  got = wf.update_norm(lambda: wf(configs), max_value=0.25 * top)()
  assert abs(got - (-10.0 + np.log(4.0))) < 1e-4
  torch.testing.assert_close(wf(configs), psi0 / 4.0, rtol=1e-4, atol=0.0)
  # log-domain form used by the optimizers: same rule, no exp overflow
  got2 = wf.update_norm(None, max_value=1.0, log_amplitudes=lambda: wf.log_amplitude(configs))()
  assert abs(got2 - (got + np.log(top / 4.0))) < 1e-4
  assert abs(float(wf(configs).max()) - 1.0) < 1e-4
  # normalize_batch (wavefunctions.py:234-257) always rescales
  got3 = wf.normalize_batch(lambda: wf(configs), max_value=8.0)()
  assert abs(got3 - (got2 - np.log(8.0))) < 1e-4


def test_unseeded_wavefunctions_differ():
  """Two unseeded modules start from different parameters (the reference draws
  from TF's unseeded stream): `diff` of two same-type leaves is not psi = 0."""
  from cgs_vmc_b200 import utils, wavefunctions
  hp = utils.create_hparams(wavefunction_type='diff', num_sites=12, num_fc_layers=1, fc_layer_size=6,
                            composite_wavefunction_types=('fully_connected', 'fully_connected'),
                            composite_output_activations=('exp', 'exp'))
  wf = wavefunctions.build_wavefunction(hp)
  cfg = torch.from_numpy(bits.random_sz0_configs(12, 32, np.random.default_rng(2))).cuda()
  psi = wf(cfg)
  assert torch.isfinite(psi).all() and (psi != 0).all()
  a, b = wf.leaves()
  assert not torch.equal(a.native().params, b.native().params)


def test_drivers_end_to_end(tmp_path):
  """run_training -> run_energy_evaluation -> run_supervised_training."""
  ckpt = str(tmp_path / 'gs')
  env = dict(os.environ, PYTHONPATH=REPO)
  common = 'batch_size=256,num_batches_per_epoch=3,num_equilibration_sweeps=2,fc_layer_size=8,num_fc_layers=0'
  subprocess.run([sys.executable, os.path.join(REPO, 'run_training.py'), '--checkpoint_dir', ckpt,
                  '--num_sites', '8', '--num_epochs', '3', '--wavefunction_type', 'rbm',
                  '--heisenberg_jx', '-1.0', '--optimizer', 'EnergyGradient', '--hparams', common],
                 check=True, env=env, timeout=300)
  metrics = open(os.path.join(ckpt, 'metrics.txt')).read().split()
  assert len(metrics) == 3 and all(np.isfinite(float(m)) for m in metrics)
  assert os.path.exists(os.path.join(ckpt, 'model_prior_2_epochs.pt'))
  assert os.path.exists(os.path.join(ckpt, 'hparams.pbtxt'))
  out = subprocess.run([sys.executable, os.path.join(REPO, 'run_energy_evaluation.py'),
                        '--checkpoint_dir', ckpt, '--heisenberg_jx', '-1.0',
                        '--hparams', 'num_evaluation_samples=4'],
                       check=True, env=env, timeout=300, capture_output=True, text=True)
  assert 'Energy:' in out.stdout
  swo = str(tmp_path / 'swo')
  subprocess.run([sys.executable, os.path.join(REPO, 'run_supervised_training.py'),
                  '--checkpoint_dir', swo, '--supervisor_dir', ckpt, '--num_epochs', '2',
                  '--wavefunction_type', 'fully_connected',
                  '--hparams', 'batch_size=128,num_batches_per_epoch=2,num_fc_layers=1,fc_layer_size=8'],
                 check=True, env=env, timeout=300)
  assert os.path.exists(os.path.join(swo, 'model_after_1_epochs.pt'))
  assert len(open(os.path.join(swo, 'metrics.txt')).read().split()) == 2


def test_empty_batches_through_the_late_entry_points():
  """B = 0 / n = 0 are no-ops with status OK (the reference's ops accept empty
  batches): host packing, upload, standalone periodic convolution, epoch end."""
  from cgs_vmc_b200 import _native
  out = torch.zeros(0, 1, dtype=torch.int64)
  _native.pack_configs_host(torch.zeros(0, 36), out)
  y = _native.conv_periodic(torch.zeros(0, 6, 6, 1, device='cuda'), torch.zeros(3, 3, 1, 4, device='cuda'))
  assert tuple(y.shape) == (0, 6, 6, 4)
  z = torch.zeros(0, device='cuda')
  _native.epoch_end(z, z.clone(), z.clone(), torch.zeros(2, 0, device='cuda'),
                    torch.zeros(4, dtype=torch.float64, device='cuda'),
                    torch.zeros(1, dtype=torch.int32, device='cuda'), t=1)
  with pytest.raises(ValueError):
    _native.conv_periodic(torch.zeros(2, 6, 6, 2, device='cuda'), torch.zeros(3, 3, 1, 4, device='cuda'))

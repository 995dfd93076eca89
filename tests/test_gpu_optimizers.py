"""LogOverlapSWO, DualSamplingSWO and LogOverlapImaginaryTimeSWO on the CUDA
path (cgs_vmc_b200.training) against gradients recorded from the reference's
own build_opt_ops (tests/golden/opt_*.npz, made by make_golden_optimizers.py),
plus end-to-end convergence checks against exact diagonalisation."""
import numpy as np
import pytest
import torch

from conftest import load_golden, opt_golden_names
from oracle import ansatz as oansatz
from oracle import ed, lattices

pytestmark = pytest.mark.gpu

GRAD_RTOL = 3e-4     # of the largest entry (float32 estimator sums)


def _hparams(spec, **extra):
  from cgs_vmc_b200 import utils
  if spec.kind in ('fully_connected', 'rbm'):
    kw = dict(num_fc_layers=spec.num_layers, fc_layer_size=spec.layer_size)
  else:
    kw = dict(num_conv_layers=spec.num_layers, num_conv_filters=spec.num_filters,
              kernel_size=spec.kernel_size, size_x=spec.size_x, size_y=spec.size_y)
  kw.update(extra)
  return utils.create_hparams(wavefunction_type=spec.kind, num_sites=spec.n_sites,
                              nonlinearity=spec.nonlinearity, **kw)


def _wavefunction(hp, flat, shift):
  from cgs_vmc_b200 import wavefunctions
  wf = wavefunctions.build_wavefunction(hp)
  wf.native(hp.num_sites).set_params(torch.from_numpy(flat))
  wf._exp_norm_shift = float(shift)
  return wf


def _close(got, ref, rtol=GRAD_RTOL):
  scale = float(np.abs(ref).max())
  np.testing.assert_allclose(got.detach().cpu().numpy(), ref, rtol=0, atol=rtol * scale)


@pytest.mark.parametrize('name', opt_golden_names())
def test_log_overlap_swo_gradient_golden(name):
  from cgs_vmc_b200 import graph_builders, training
  from cgs_vmc_b200.session import Session
  spec, g = load_golden(name)
  hp = _hparams(spec, batch_size=g['lo_configs'].shape[0])
  wf = _wavefunction(hp, g['params_flat'], g['shift'])
  target = _wavefunction(hp, g['target_params_flat'], g['target_shift'])
  opt = training.SUPERVISED_OPTIMIZERS['LogOverlapSWO']()
  shared = {}
  ops = opt.build_opt_ops(wavefunction=wf, target_wavefunction=target, hparams=hp,
                          shared_resources=shared)
  shared[graph_builders.ResourceName.CONFIGS].assign(torch.from_numpy(g['lo_configs']))
  s = Session()
  s.run(ops.reset_gradients)
  s.run(ops.accumulate_gradients)
  _close(opt.sums.gradient(), g['lo_gradient'])
  # tf.metrics.mean_tensor semantics: two identical accumulate calls average back
  s.run(ops.accumulate_gradients)
  _close(opt.sums.gradient(), g['lo_gradient'])


@pytest.mark.parametrize('name', opt_golden_names())
def test_dual_sampling_swo_golden(name):
  from cgs_vmc_b200 import graph_builders, training
  spec, g = load_golden(name)
  half = g['ds_psi_configs'].shape[0]
  hp = _hparams(spec, batch_size=2 * half)
  wf = _wavefunction(hp, g['params_flat'], g['shift'])
  # ds_target_shift already holds the -N/2 ln 2 that the optimizer's sqrt(2^N) undoes
  target = _wavefunction(hp, g['target_params_flat'], g['ds_target_shift'])
  opt = training.SUPERVISED_OPTIMIZERS['DualSamplingSWO']()
  shared = {}
  ops = opt.build_opt_ops(wavefunction=wf, target_wavefunction=target, hparams=hp,
                          shared_resources=shared)
  assert shared[graph_builders.ResourceName.CONFIGS].shape == (half, spec.n_sites)
  assert shared[graph_builders.ResourceName.TARGET_CONFIGS].shape == (half, spec.n_sites)
  shared[graph_builders.ResourceName.CONFIGS].assign(torch.from_numpy(g['ds_psi_configs']))
  shared[graph_builders.ResourceName.TARGET_CONFIGS].assign(torch.from_numpy(g['ds_target_configs']))
  packed, loss, weights = opt.loss_and_weights()
  assert abs(float(loss) - float(g['ds_loss'])) <= 2e-4 * abs(float(g['ds_loss']))
  grad = wf.native().weighted_grad_sum(packed, weights)
  _close(grad[0], g['ds_gradient'])
  assert ops.accumulate_gradients is None and ops.reset_gradients is None


@pytest.mark.parametrize('name', opt_golden_names())
def test_log_overlap_itswo_golden(name):
  from cgs_vmc_b200 import graph_builders, operators, training
  from cgs_vmc_b200.session import Session
  spec, g = load_golden(name)
  hp = _hparams(spec, batch_size=g['it_configs'].shape[0], time_evolution_beta=float(g['it_beta']))
  wf = _wavefunction(hp, g['params_flat'], g['shift'])
  ham = operators.HeisenbergHamiltonian([tuple(b) for b in g['bonds_ij']],
                                        float(g['bonds_jx'][0]), float(g['bonds_jz'][0]))
  opt = training.GROUND_STATE_OPTIMIZERS['LogOverlapITSWO']()
  shared = {}
  ops = opt.build_opt_ops(wavefunction=wf, hamiltonian=ham, hparams=hp, shared_resources=shared)
  assert opt.supervisor is not wf and opt.supervisor._unique_name.startswith('dc_')
  opt.supervisor.native().set_params(torch.from_numpy(g['it_omega_params_flat']))
  opt.supervisor._exp_norm_shift = float(g['it_omega_shift'])
  shared[graph_builders.ResourceName.CONFIGS].assign(torch.from_numpy(g['it_configs']))
  s = Session()
  s.run(ops.reset_gradients)
  s.run(ops.accumulate_gradients)
  _close(opt.sums.gradient(), g['it_gradient'])
  e = s.run(ops.energy)
  assert abs(e - float(g['it_energy'])) <= 5e-5 * (1 + abs(float(g['it_energy'])))
  # update_supervisor copies the trainee's variables (wavefunctions.py:300-325)
  s.run(ops.update_supervisor)
  assert torch.equal(opt.supervisor.native().params, wf.native().params)
  assert opt.supervisor.native().params.data_ptr() != wf.native().params.data_ptr()


def _all_sz0(n):
  idx = [i for i in range(2 ** n) if bin(i).count('1') == n // 2]
  c = np.array([[1.0 if (i >> k) & 1 else -1.0 for k in range(n)] for i in idx], dtype=np.float32)
  return torch.from_numpy(c)


def _overlap(wf, target, cfg):
  a = wf.log_amplitude(cfg).double()
  b = target.log_amplitude(cfg).double()
  pa, pb = torch.exp(a - a.max()), torch.exp(b - b.max())
  return float((pa * pb).sum() ** 2 / ((pa * pa).sum() * (pb * pb).sum()))


def test_log_overlap_itswo_reaches_ed_energy():
  """run_optimization_epoch end to end on the 8-site chain: imaginary-time
  SWO drives the supervisor energy to exact diagonalisation (-3.6511)."""
  from cgs_vmc_b200 import graph_builders, operators, training, utils, wavefunctions
  from cgs_vmc_b200.session import Session
  graph_builders.reset_num_epochs()
  hp = utils.create_hparams(wavefunction_type='rbm', num_sites=8, num_fc_layers=0, fc_layer_size=16,
                            batch_size=2048, num_batches_per_epoch=10, num_equilibration_sweeps=3,
                            time_evolution_beta=0.1,
                            learning_rates=[0.01, 0.005, 0.002, 0.001], learning_rate_stops=[60, 100, 140])
  wf = wavefunctions.build_wavefunction(hp).seed(2)
  ham = operators.HeisenbergHamiltonian(lattices.chain_bonds(8), -1.0, 1.0)
  opt = training.GROUND_STATE_OPTIMIZERS['LogOverlapITSWO']()
  ops = opt.build_opt_ops(wavefunction=wf, hamiltonian=ham, hparams=hp, shared_resources={})
  s = Session()
  energies = [opt.run_optimization_epoch(ops, s, hp) for _ in range(120)]
  e0, _, _ = ed.ground_state(8, *lattices.heisenberg_couplings(lattices.chain_bonds(8)))
  assert energies[0] > np.mean(energies[-10:])
  assert abs(np.mean(energies[-10:]) - e0) < 0.03 * abs(e0), (energies[0], energies[-10:], e0)


@pytest.mark.parametrize('optimizer', ['LogOverlapSWO', 'DualSamplingSWO'])
def test_supervised_optimizers_increase_overlap(optimizer):
  """Exact overlap |<psi|phi>|^2 / (<psi|psi><phi|phi>) over all 70 Sz=0 states
  of 8 sites grows under training."""
  from cgs_vmc_b200 import graph_builders, training, utils, wavefunctions
  from cgs_vmc_b200.session import Session
  graph_builders.reset_num_epochs()
  hp = utils.create_hparams(wavefunction_type='rbm', num_sites=8, num_fc_layers=0, fc_layer_size=8,
                            batch_size=2048, num_batches_per_epoch=10,
                            learning_rates=[0.01, 0.01, 0.01, 0.01])
  target = wavefunctions.build_wavefunction(hp).seed(7)
  trainee = wavefunctions.build_wavefunction(hp).seed(8)
  target.native(8)
  with torch.no_grad():
    target.native().params.mul_(3.0)          # a target with structure
  trainee.native(8)
  cfg = _all_sz0(8).cuda()
  if optimizer == 'DualSamplingSWO':          # the plain L2 loss is not scale free
    za, zb = trainee.log_amplitude(cfg), target.log_amplitude(cfg)
    target._exp_norm_shift += float((zb - za).mean()) + 0.5 * 8 * np.log(2.0)
    trainee._exp_norm_shift += float(za.max())
    target._exp_norm_shift += float(za.max())
  before = _overlap(trainee, target, cfg)
  opt = training.SUPERVISED_OPTIMIZERS[optimizer]()
  ops = opt.build_opt_ops(wavefunction=trainee, target_wavefunction=target, hparams=hp,
                          shared_resources={})
  s = Session()
  for epoch in range(40):
    opt.run_optimization_epoch(ops, s, hp, epoch)
  after = _overlap(trainee, target, cfg)
  assert after > before and 1.0 - after < 0.5 * (1.0 - before), (before, after)


# ---------------------------------------------------------------------------
# SWO loss / weights and the optimizer update on the device
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['rbm_6x6', 'fc_chain20', 'conv2d_6x6_k3'])
def test_swo_weights_kernel_golden(name):
  """cgsvmc_swo_weights + cgsvmc_weighted_grad_sum against the loss and the
  gradient the reference's SupervisedWavefunctionOptimizer graph produced
  (training.py:166-175)."""
  from cgs_vmc_b200 import _native
  from gpu_util import make_native, packed_cuda
  spec, g = load_golden(name)
  a = make_native(spec, g['params_flat'])
  target = make_native(spec, g['swo_target_params_flat'])
  packed = packed_cuda(g['swo_configs'])
  b = packed.shape[0]
  z, zt = a.log_amp(packed), target.log_amp(packed)
  log_norm = 0.5 * spec.n_sites * np.log(2.0) - float(g['swo_target_shift']) + float(g['shift'])
  acc = torch.zeros(2, dtype=torch.float64, device='cuda')
  w = _native.swo_weights(z, zt, log_norm, b, loss_acc=acc)
  assert acc[1].item() == b
  loss = (acc[0] / acc[1]).item()
  assert abs(loss - float(g['swo_loss'])) <= 5e-4 * abs(loss) + 1e-6
  grad = a.weighted_grad_sum(packed, w)[0].cpu().numpy()
  ref = g['swo_gradient']
  assert np.linalg.norm(grad - ref) <= 1e-3 * np.linalg.norm(ref) + 1e-6
  # signs: r -> sign * sign_t * r
  sg = torch.where(torch.arange(b, device='cuda') % 2 == 0, 1.0, -1.0).float()
  w2 = _native.swo_weights(z, zt, log_norm, b, sign=sg)
  r = 1.0 - w[0] * b / 2.0
  torch.testing.assert_close(w2[0], 2.0 * (1.0 - sg * r) / b, rtol=1e-5, atol=1e-6)


def test_adam_step_kernel_matches_tf_rule():
  """cgsvmc_adam_step against tf.train.AdamOptimizer's update rule in float64
  (training.py:76-91), with the gradient given and with the energy gradient
  formed from the estimator sums (training.py:562-564); device-side t / lr."""
  from cgs_vmc_b200 import _native
  rng = np.random.default_rng(0)
  n, b1, b2, eps, lr = 1000, 0.9, 0.99, 1e-8, 0.01
  p0 = rng.normal(size=n)
  params = torch.from_numpy(p0).float().cuda()
  m, v = torch.zeros_like(params), torch.zeros_like(params)
  pr, mr, vr = p0.astype(np.float32).astype(np.float64), np.zeros(n), np.zeros(n)
  t_dev = torch.zeros(1, dtype=torch.int64, device='cuda')
  lr_dev = torch.full((1,), lr, dtype=torch.float32, device='cuda')
  for t in range(1, 6):
    sums = rng.normal(size=(2, n)).astype(np.float32)
    stats = np.array([-3.5 * 64, 900.0, 64.0, 0.0])
    nb = 4.0
    if t % 2:
      g = rng.normal(size=n).astype(np.float32).astype(np.float64)
      kw = dict(grad=torch.from_numpy(g).float().cuda())
    else:
      g = sums[1].astype(np.float64) / nb - (stats[0] / stats[2]) * sums[0].astype(np.float64) / nb
      kw = dict(sums=torch.from_numpy(sums).cuda(), stats=torch.from_numpy(stats).cuda(), num_batches=nb)
    version = params._version
    if t <= 3:
      _native.adam_step(params, m, v, lr=lr, beta1=b1, beta2=b2, eps=eps, t=t, **kw)
    else:      # device scalars (CUDA-graph form): t_dev holds the steps taken so far
      t_dev.fill_(t - 1)
      _native.adam_step(params, m, v, lr_dev=lr_dev, beta1=b1, beta2=b2, eps=eps, t_dev=t_dev, **kw)
      assert t_dev.item() == t
    assert params._version > version          # ansatz handles see the change
    mr = b1 * mr + (1 - b1) * g
    vr = b2 * vr + (1 - b2) * g * g
    pr = pr - lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t) * mr / (np.sqrt(vr) + eps)
    np.testing.assert_allclose(params.cpu().numpy(), pr, rtol=2e-5, atol=2e-6)
  with pytest.raises(ValueError):
    _native.adam_step(params, m, v, lr=lr, t=1)           # neither grad nor sums


@pytest.mark.parametrize('use_payload', [False, True])
def test_epoch_end_kernel_equals_separate_ops(use_payload):
  """cgsvmc_epoch_end (apply_gradients + metrics + reset_gradients of
  training.py:618-622 in one kernel) against cgsvmc_adam_step on the same sums
  followed by the read-back and the reset: identical parameters and moments,
  statistics in pinned host memory, accumulators zeroed, ticket re-armed."""
  from cgs_vmc_b200 import _native
  _native.require_cuda()
  g = torch.Generator().manual_seed(5)
  n = 70001
  params0 = torch.randn(n, generator=g).cuda()
  ticket = torch.zeros(1, dtype=torch.int32, device='cuda')
  host = torch.zeros(4, dtype=torch.float64).pin_memory()
  pa, ma, va = params0.clone(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
  pb, mb, vb = params0.clone(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
  for t in range(1, 4):
    sums = torch.randn(2, n, generator=g).cuda() * 50
    stats = torch.tensor([-1234.5 * t, 2.0e6, 4096.0 * t, 0.0], dtype=torch.float64, device='cuda')
    local_sums, local_stats = sums.clone(), stats.clone()
    kw = {}
    if use_payload:      # walker-sharded run: the all-reduced float64 payload of two equal shards
      sums, stats = 2 * sums, 2 * stats
      kw['total_payload'] = torch.cat([sums.reshape(-1).double(), stats])
    _native.adam_step(pa, ma, va, sums=sums, stats=stats, num_batches=3, lr=0.01, beta2=0.99, t=t)
    version = pb._version
    _native.epoch_end(pb, mb, vb, local_sums, local_stats, ticket, num_batches=3, lr=0.01, beta2=0.99, t=t,
                      stats_out=host, **kw)
    torch.cuda.synchronize()
    assert pb._version > version
    assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb)
    assert torch.equal(host, stats.cpu())
    assert float(local_sums.abs().max()) == 0.0 and float(local_stats.abs().max()) == 0.0
    assert int(ticket.item()) == 0


def test_supervised_captured_batch_equals_eager_ops():
  """SupervisedWavefunctionOptimizer: the replayed graph of one batch (sweep +
  train step) leaves the same parameters and walkers as session.run(mc_step) x
  N followed by session.run(apply_gradients) (training.py:208-212)."""
  from cgs_vmc_b200 import graph_builders, training, utils, wavefunctions
  from cgs_vmc_b200.session import Session
  results = []
  for use_graph in (True, False):
    graph_builders.reset_num_epochs()
    hp = utils.create_hparams(wavefunction_type='rbm', num_sites=16, num_fc_layers=0, fc_layer_size=12,
                              batch_size=512, num_batches_per_epoch=3, learning_rates=[0.01] * 4)
    target = wavefunctions.build_wavefunction(hp).seed(7)
    trainee = wavefunctions.build_wavefunction(hp).seed(8)
    target.native(16)
    target._exp_norm_shift += 0.5 * 16 * np.log(2.0)
    opt = training.SUPERVISED_OPTIMIZERS['SWO']()
    opt.use_cuda_graph = use_graph
    shared = {}
    ops = opt.build_opt_ops(wavefunction=trainee, target_wavefunction=target, hparams=hp,
                            shared_resources=shared)
    assert (opt._batch_step is not None) == use_graph
    s = Session()
    for epoch in range(2):
      opt.run_optimization_epoch(ops, s, hp, epoch)
    loss = s.run(ops.metrics)
    configs = shared[graph_builders.ResourceName.CONFIGS]
    results.append((trainee.flat_parameters.clone(), configs.packed.clone(), loss, configs.state.step))
  (p1, c1, l1, s1), (p2, c2, l2, s2) = results
  assert s1 == s2 == 2 * 3 * 16
  assert torch.equal(c1, c2)
  torch.testing.assert_close(p1, p2, rtol=1e-5, atol=1e-6)
  assert abs(l1 - l2) <= 1e-5 * abs(l2) + 1e-7

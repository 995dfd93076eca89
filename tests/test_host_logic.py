"""CPU tests of the host-side mirror: hyper-parameters, session shim,
registries / error behaviour, sharding arithmetic and the packed all-reduce
(world_size 2, gloo)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cgs_vmc_b200 import distributed, drivers, graph_builders, session, training, utils, wavefunctions


def test_hparams_defaults_match_reference():
  hp = utils.create_hparams()
  # utils.py:87-148
  assert (hp.num_sites, hp.num_fc_layers, hp.fc_layer_size) == (40, 3, 80)
  assert (hp.num_conv_layers, hp.kernel_size, hp.num_conv_filters) == (5, 5, 16)
  assert (hp.num_equilibration_sweeps, hp.num_monte_carlo_sweeps) == (100, 1)
  assert (hp.batch_size, hp.num_batches_per_epoch, hp.num_epochs) == (200, 50, 500)
  assert hp.learning_rates == [1e-3, 1e-4, 2e-5, 1e-5] and hp.learning_rate_stops == [300, 600, 1000]
  assert (hp.optimizer, hp.beta2, hp.nonlinearity, hp.output_activation) == ('adam', 0.99, 'relu', 'exp')


def test_hparams_parse_override_roundtrip(tmp_path):
  hp = utils.create_hparams(num_sites=36, wavefunction_type='rbm')
  hp.parse('batch_size=8192,fc_layer_size=144,num_fc_layers=0,learning_rates=[0.01,0.001],nonlinearity=tanh')
  assert hp.batch_size == 8192 and hp.fc_layer_size == 144 and hp.num_fc_layers == 0
  assert hp.learning_rates == [0.01, 0.001] and hp.nonlinearity == 'tanh'
  with pytest.raises(KeyError):
    hp.set_hparam('no_such_param', 1)
  with pytest.raises(KeyError):
    utils.create_hparams(bogus=3)
  path = tmp_path / 'hparams.pbtxt'
  utils.save_hparams(hp, str(path))
  back = utils.load_hparams(str(path))
  assert back.values() == hp.values()


def test_piecewise_constant_boundaries():
  # tf.train.piecewise_constant: boundaries are inclusive on the left value
  f = lambda x: training.piecewise_constant(x, [300, 600, 1000], [1e-3, 1e-4, 2e-5, 1e-5])
  assert (f(0), f(300), f(301), f(600), f(601), f(1000), f(1001)) == \
      (1e-3, 1e-3, 1e-4, 1e-4, 2e-5, 2e-5, 1e-5)


def test_session_runs_ops_and_lists():
  s = session.Session()
  calls = []
  op = session.Op(lambda n_steps=1: calls.append(n_steps) or n_steps * 2, 'op')
  assert s.run(op) == 2 and s.run(op, n_steps=5) == 10
  assert s.run([op, None, torch.tensor(3.5)]) == [2, None, 3.5]
  assert calls == [1, 5, 1]


def test_registries_and_error_types():
  assert set(wavefunctions.WAVEFUNCTION_TYPES) == {'fully_connected', 'rbm', 'conv_1d', 'conv_2d',
                                                   'res_net_1d', 'res_net_2d'}
  assert set(training.GROUND_STATE_OPTIMIZERS) == {'EnergyGradient', 'LogOverlapITSWO', 'ITSWO'}
  assert set(training.SUPERVISED_OPTIMIZERS) == {'SWO', 'LogOverlapSWO', 'DualSamplingSWO', 'BasisIterSWO'}
  with pytest.raises(ValueError, match='not registered'):          # wavefunctions.py:1196
    wavefunctions.build_wavefunction(utils.create_hparams(wavefunction_type='nope'))
  with pytest.raises(NotImplementedError):
    wavefunctions.build_wavefunction(utils.create_hparams(wavefunction_type='mps'))
  with pytest.raises(NotImplementedError):
    training.GROUND_STATE_OPTIMIZERS['ITSWO']()
  # signed output activations and composites are built (amplitude-agnostic route)
  assert not wavefunctions.FullyConnectedNetwork(2, 8, output_activation='cos').fast_path
  assert wavefunctions.FullyConnectedNetwork(2, 8).fast_path
  with pytest.raises(ValueError):
    wavefunctions.FullyConnectedNetwork(2, 8, output_activation='softplus')
  a, b = wavefunctions.RestrictedBoltzmannNetwork(0, 8), wavefunctions.FullyConnectedNetwork(1, 4)
  assert (a + b)._unique_name == 'restricted_boltzmann_network_plus_fully_connected_network'
  assert (a * b)._unique_name == 'fully_connected_network_times_restricted_boltzmann_network'
  assert (a * -1.)._unique_name == 'neg_1.0_times_restricted_boltzmann_network'      # wavefunctions.py:130-133
  assert [type(w).__name__ for w in (a - b)._sub_wavefunctions] == [
      'RestrictedBoltzmannNetwork', 'ProductOfWavefunctions']
  with pytest.raises(ValueError, match='not supported'):
    type(a + b).from_hparams(utils.create_hparams())
  comp = wavefunctions.build_wavefunction(utils.create_hparams(
      wavefunction_type='diff', composite_wavefunction_types=('rbm', 'fully_connected'),
      composite_output_activations=('exp', 'tanh'), num_sites=8))
  assert not comp.fast_path and len(comp._sub_wavefunctions) == 2
  wf = wavefunctions.build_wavefunction(utils.create_hparams(
      wavefunction_type='conv_2d', num_sites=36, size_x=6, size_y=6))
  assert wf._n_sites == 36 and wf._param_shapes(36)[0] == (5, 5, 1, 16)
  assert graph_builders.ResourceName.CONFIGS.value == 'CONFIGS'


def test_param_shapes_match_oracle_layout():
  from oracle import ansatz as oansatz
  cases = [
      (dict(wavefunction_type='fully_connected', num_sites=20), oansatz.AnsatzSpec('fully_connected', 20)),
      (dict(wavefunction_type='rbm', num_sites=36, num_fc_layers=0, fc_layer_size=144),
       oansatz.AnsatzSpec('rbm', 36, num_layers=0, layer_size=144)),
      (dict(wavefunction_type='rbm', num_sites=12, num_fc_layers=2, fc_layer_size=10),
       oansatz.AnsatzSpec('rbm', 12, num_layers=2, layer_size=10)),
      (dict(wavefunction_type='conv_1d', num_sites=12, num_conv_layers=3, kernel_size=4, num_conv_filters=3),
       oansatz.AnsatzSpec('conv_1d', 12, num_layers=3, num_filters=3, kernel_size=4)),
      (dict(wavefunction_type='conv_2d', num_sites=100, size_x=10, size_y=10),
       oansatz.AnsatzSpec('conv_2d', 100, num_layers=5, num_filters=16, kernel_size=5, size_x=10, size_y=10)),
  ]
  for overrides, spec in cases:
    wf = wavefunctions.build_wavefunction(utils.create_hparams(**overrides))
    assert [tuple(s) for s in wf._param_shapes(spec.n_sites)] == \
        [tuple(s) for _, s in oansatz.param_shapes(spec)]


def test_load_bonds(tmp_path):
  bonds, jx, jz = drivers.load_bonds(str(tmp_path), 6, -1.0)
  assert bonds == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 0)] and (jx, jz) == (-1.0, 1.0)
  (tmp_path / 'J.txt').write_text('0 1\n1 2\n2 0\n')
  bonds, jx, jz = drivers.load_bonds(str(tmp_path), 3, 1.0)
  assert bonds == [(0, 1), (1, 2), (2, 0)] and (jx, jz) == (1.0, 1.0)
  (tmp_path / 'J.txt').write_text('0 1 -1.0 1.0\n0 2 0.5 0.5\n')
  bonds, jx, jz = drivers.load_bonds(str(tmp_path), 3, 1.0)
  assert bonds == [(0, 1), (0, 2)]
  np.testing.assert_allclose(jx, [-1.0, 0.5]); np.testing.assert_allclose(jz, [1.0, 0.5])


def test_pack_unpack_sums():
  sums = torch.arange(10, dtype=torch.float32).reshape(2, 5)
  stats = torch.tensor([1.5, 2.5, 7.0, 0.0], dtype=torch.float64)
  payload = distributed.pack_sums(sums, stats)
  assert payload.dtype == torch.float64 and payload.numel() == 14
  s2, t2 = torch.zeros_like(sums), torch.zeros_like(stats)
  distributed.unpack_sums(payload, s2, t2)
  assert torch.equal(s2, sums) and torch.equal(t2, stats)
  assert distributed.world_size() == 1 and distributed.shard(64) == (64, 0)


def _worker(rank, world, port, out_dir):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    local, w0 = distributed.shard(64)
    assert (local, w0) == (32, rank * 32)
    with pytest.raises(ValueError):
      distributed.shard(63)
    # each rank holds the partial sums of its walkers
    g = torch.Generator().manual_seed(5)
    per_walker = torch.randn(64, 2, 7, generator=g)          # same on both ranks
    e = torch.randn(64, generator=g).double()
    mine = slice(w0, w0 + local)
    sums = per_walker[mine].sum(0)
    stats = torch.tensor([e[mine].sum(), (e[mine] ** 2).sum(), float(local), 0.0], dtype=torch.float64)
    distributed.allreduce_sums(sums, stats)
    torch.testing.assert_close(sums, per_walker.sum(0), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(stats[:3], torch.tensor([e.sum(), (e ** 2).sum(), 64.0], dtype=torch.float64),
                               rtol=1e-6, atol=1e-6)
    m = distributed.allreduce_(torch.tensor([float(rank)]), op='max')
    assert m.item() == world - 1
    # the payload cgsvmc_epoch_end reads directly: all-reduced in float64, local accumulators untouched
    part_s = per_walker[mine].sum(0)
    part_t = torch.tensor([e[mine].sum(), (e[mine] ** 2).sum(), float(local), 0.0], dtype=torch.float64)
    before = (part_s.clone(), part_t.clone())
    payload = distributed.allreduce_payload(part_s, part_t)
    assert payload.dtype == torch.float64 and payload.numel() == part_s.numel() + 4
    assert torch.equal(part_s, before[0]) and torch.equal(part_t, before[1])
    torch.testing.assert_close(payload[:part_s.numel()].view_as(part_s).float(), sums, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(payload[part_s.numel():], stats, rtol=1e-12, atol=1e-12)
    # totals into separate buffers: the local accumulators stay local, so a
    # second reduce after further accumulation does not double count
    # (metrics in the middle of an epoch, then accumulate again)
    local_sums = per_walker[mine].sum(0)
    local_stats = torch.tensor([e[mine].sum(), (e[mine] ** 2).sum(), float(local), 0.0], dtype=torch.float64)
    keep = (local_sums.clone(), local_stats.clone())
    tot_s, tot_t = torch.zeros_like(local_sums), torch.zeros_like(local_stats)
    for _ in range(2):
      distributed.allreduce_sums(local_sums, local_stats, tot_s, tot_t)
      assert torch.equal(local_sums, keep[0]) and torch.equal(local_stats, keep[1])
      torch.testing.assert_close(tot_s, per_walker.sum(0), rtol=1e-5, atol=1e-5)
      assert tot_t[2].item() == 64.0
    # walker counts beyond 2^24 and sum E^2 stay exact (float64 payload)
    big = torch.tensor([1.0, 3.0, float(2 ** 24 + 1 + rank), 0.0], dtype=torch.float64)
    _, big_tot = distributed.allreduce_sums(torch.zeros(1, 3), big, torch.zeros(1, 3), torch.zeros(4, dtype=torch.float64))
    assert big_tot[2].item() == float(2 * (2 ** 24 + 1) + 1)
    # update_norm (wavefunctions.py:261-288): tf.reduce_max runs over the WHOLE
    # batch, so the shards' maxima are all-reduced and every rank applies the
    # same shift; only rank 1 holds the amplitude above max_value here
    wf = wavefunctions.RestrictedBoltzmannNetwork(0, 4)
    wf._exp_norm_shift = -10.0
    amps = torch.tensor([1.0, 2.0]) if rank == 0 else torch.tensor([3.0, 1e12])
    shift = wf.update_norm(amps)()
    assert abs(shift - (-10.0 + np.log(1e12) - np.log(1e10))) < 1e-5
    shift2 = wf.update_norm(None, log_amplitudes=torch.log(amps.double()))()    # log-domain form
    assert abs(shift2 - (shift + np.log(1e12) - np.log(1e10))) < 1e-5
    assert wf.update_norm(torch.tensor([5.0, 7.0]))() == shift2              # below max_value: unchanged
    # parameter replicas start from rank 0's values
    prm = torch.full((3,), float(rank + 1))
    distributed.broadcast_(prm)
    assert torch.equal(prm, torch.ones(3))
    open(os.path.join(out_dir, 'ok%d' % rank), 'w').close()
  finally:
    dist.destroy_process_group()


def test_sharded_allreduce_gloo_world2(tmp_path):
  """The N > 1 host path on CPU: sharding + packed all-reduce give the
  single-rank sums."""
  port = 29500 + os.getpid() % 2000
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  assert sorted(os.listdir(tmp_path)) == ['ok0', 'ok1']


def test_update_norm_reference_rule():
  """wavefunctions.py:261-288 / 234-257 on one rank: shift += log(max psi) -
  log(max_value) iff max psi > max_value; normalize_batch always."""
  wf = wavefunctions.FullyConnectedNetwork(1, 4)
  wf._exp_norm_shift = -10.0
  assert wf.update_norm(torch.tensor([1.0, 9.9e9]))() == -10.0
  got = wf.update_norm(torch.tensor([1.0, 4e10]))()
  assert abs(got - (-10.0 + np.log(4e10 / 1e10))) < 1e-5
  got2 = wf.normalize_batch(torch.tensor([0.5, 2.0]), max_value=4.0)()
  assert abs(got2 - (got + np.log(2.0 / 4.0))) < 1e-6
  # callable inputs are evaluated when the op runs (graph semantics)
  box = {'v': torch.tensor([1.0])}
  op = wf.update_norm(lambda: box['v'], max_value=10.0)
  assert op() == got2
  box['v'] = torch.tensor([1000.0])
  assert abs(op() - (got2 + np.log(100.0))) < 1e-5
  # signed outputs carry no exp_norm_shift: no op (wavefunctions.py:280-281)
  assert wavefunctions.FullyConnectedNetwork(1, 4, output_activation='tanh').update_norm(torch.ones(2)) is None


def test_unseeded_wavefunctions_get_distinct_init_streams():
  a, b = wavefunctions._next_init_seed(), wavefunctions._next_init_seed()
  assert a != b and 0 <= a < 2 ** 63 and 0 <= b < 2 ** 63


def test_hparams_parse_unquoted_lists():
  """tf.contrib.training.HParams.parse accepts name=[a,b] with bare strings."""
  h = utils.create_hparams()
  h.parse('composite_wavefunction_types=[rbm,fully_connected],composite_output_activations=[exp,tanh],'
          'batch_size=8,learning_rates=[0.1,0.2],nonlinearity=relu')
  assert h.composite_wavefunction_types == ('rbm', 'fully_connected')
  assert h.composite_output_activations == ('exp', 'tanh')
  assert h.batch_size == 8 and h.learning_rates == [0.1, 0.2] and h.nonlinearity == 'relu'
  with pytest.raises(ValueError):
    h.parse('composite_wavefunction_types=[rbm]')
  with pytest.raises(ValueError):
    h.parse('learning_rates=0.1')

"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference's Python (wavefunctions.py, layers.py, operators.py,
graph_builders.py, training.py under /root/reference/cgs_vmc) is imported
as-is on top of the eager TensorFlow/Sonnet stand-in in tests/golden/tf_shim
(TF 1.x cannot be installed here).  For every case we record the inputs the
reference consumed (parameters, configurations, uniform draws) and the outputs
it produced (psi, post-step configurations, flipped configurations, diag /
offdiag / local energies, energy gradients, SWO loss + gradient).  The oracle
(oracle/) and the CUDA library are both tested against these vectors.

Parameters are drawn by oracle.ansatz.init_params (Sonnet default
initialisers, plus non-zero biases so that the bias paths are exercised) and
written into the reference's variables in their creation order, which doubles
as a check that the flat parameter layout of include/cgsvmc.h matches the
reference's variable shapes.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
sys.path.insert(0, '/root/reference/cgs_vmc')
sys.path.insert(0, REPO)

import tensorflow as tf              # noqa: E402  (the shim)
import graph_builders                # noqa: E402  (reference)
import operators                     # noqa: E402  (reference)
import training                      # noqa: E402  (reference)
import utils                         # noqa: E402  (reference)
import wavefunctions                 # noqa: E402  (reference)

from oracle import ansatz as oansatz   # noqa: E402
from oracle import lattices            # noqa: E402

CASES = {
    # name: (hparams overrides, lattice)
    'fc_chain20': (dict(wavefunction_type='fully_connected', num_sites=20,
                        num_fc_layers=3, fc_layer_size=80), 'chain'),
    'fc_chain8_small': (dict(wavefunction_type='fully_connected', num_sites=8,
                             num_fc_layers=2, fc_layer_size=12), 'chain'),
    'rbm_6x6': (dict(wavefunction_type='rbm', num_sites=36, size_x=6, size_y=6,
                     num_fc_layers=0, fc_layer_size=144), 'square'),
    'rbm_chain12_hidden': (dict(wavefunction_type='rbm', num_sites=12,
                                num_fc_layers=1, fc_layer_size=10), 'chain'),
    'rbm_4x4_j1j2': (dict(wavefunction_type='rbm', num_sites=16, size_x=4,
                          size_y=4, num_fc_layers=0, fc_layer_size=24), 'j1j2'),
    'conv1d_chain12_k3': (dict(wavefunction_type='conv_1d', num_sites=12,
                               num_conv_layers=2, num_conv_filters=4,
                               kernel_size=3), 'chain'),
    'conv1d_chain12_k4': (dict(wavefunction_type='conv_1d', num_sites=12,
                               num_conv_layers=3, num_conv_filters=3,
                               kernel_size=4), 'chain'),
    'conv2d_6x6_k3': (dict(wavefunction_type='conv_2d', num_sites=36, size_x=6,
                           size_y=6, num_conv_layers=3, num_conv_filters=4,
                           kernel_size=3), 'square'),
    'conv2d_4x4_k2': (dict(wavefunction_type='conv_2d', num_sites=16, size_x=4,
                           size_y=4, num_conv_layers=2, num_conv_filters=3,
                           kernel_size=2), 'square'),
    'conv2d_4x6_k3': (dict(wavefunction_type='conv_2d', num_sites=24, size_x=4,
                           size_y=6, num_conv_layers=2, num_conv_filters=5,
                           kernel_size=3), 'rect'),
    'conv2d_10x10': (dict(wavefunction_type='conv_2d', num_sites=100,
                          size_x=10, size_y=10, num_conv_layers=5,
                          num_conv_filters=16, kernel_size=5), 'j1j2'),
}
BATCH = {'conv2d_10x10': 4, 'fc_chain20': 16}
DEFAULT_BATCH = 12


def spec_from_hparams(hp):
  kind = hp.wavefunction_type
  if kind in ('fully_connected', 'rbm'):
    return oansatz.AnsatzSpec(kind, hp.num_sites, num_layers=hp.num_fc_layers,
                              layer_size=hp.fc_layer_size)
  if kind in ('res_net_1d', 'res_net_2d'):
    return oansatz.AnsatzSpec(kind, hp.num_sites, num_layers=hp.num_resnet_blocks,
                              num_filters=hp.num_conv_filters,
                              kernel_size=hp.kernel_size, size_x=hp.size_x,
                              size_y=hp.size_y, nonlinearity='selu')
  return oansatz.AnsatzSpec(kind, hp.num_sites, num_layers=hp.num_conv_layers,
                            num_filters=hp.num_conv_filters,
                            kernel_size=hp.kernel_size, size_x=hp.size_x,
                            size_y=hp.size_y)


def bonds_for(lattice, hp):
  if lattice == 'chain':
    ij, jx, jz = lattices.heisenberg_couplings(
        lattices.chain_bonds(hp.num_sites), -1.0, 1.0)
  elif lattice in ('square', 'rect'):
    ij, jx, jz = lattices.heisenberg_couplings(
        lattices.square_nn_bonds(hp.size_x, hp.size_y), -1.0, 1.0)
  elif lattice == 'j1j2':
    ij, jx, jz = lattices.j1j2_couplings(hp.size_x, 0.5)
  else:
    raise ValueError(lattice)
  return ij, jx, jz


def build_reference_wavefunction(hp, seed, dummy):
  """Builds the reference ansatz and loads oracle-drawn parameters into it."""
  wf = wavefunctions.build_wavefunction(hp)
  wf(dummy)                                   # creates the variables
  spec = spec_from_hparams(hp)
  params = oansatz.init_params(spec, seed=seed, bias_scale=0.1)
  variables = wf.get_trainable_variables()
  assert len(variables) == len(params), (len(variables), len(params))
  for var, p, (name, shape) in zip(variables, params, oansatz.param_shapes(spec)):
    assert tuple(var.shape) == tuple(shape), (var.name, tuple(var.shape), shape)
    tf.assign(var, p)
  return wf, spec, params


class Recorder:
  """Callable proxy that records every configuration batch the reference
  hands to the wavefunction (so flipped configurations can be pinned)."""

  def __init__(self, wf):
    self.wf = wf
    self.inputs = []

  def __call__(self, inputs):
    self.inputs.append(inputs.detach().clone().numpy())
    return self.wf(inputs)


def make_case(name, overrides, lattice, seed):
  tf._reset_shim_state()
  torch.manual_seed(seed)
  batch = BATCH.get(name, DEFAULT_BATCH)
  hp = utils.create_hparams(batch_size=batch, **overrides)
  n = hp.num_sites
  rng = np.random.RandomState(seed)
  out = {}

  configs_np = utils.random_configurations(n, batch)       # reference's own init
  dummy = torch.from_numpy(configs_np)
  wf, spec, params = build_reference_wavefunction(hp, seed, dummy)
  out['params_flat'] = oansatz.flatten(params).numpy()
  out['configs'] = configs_np

  # (1) amplitudes, wavefunctions.py _build
  with torch.no_grad():
    out['psi_default_shift'] = wf(torch.from_numpy(configs_np)).numpy()
    # Recentre exp_norm_shift (normalize_batch, wavefunctions.py:234-259) so
    # that amplitudes stay below update_norm's 1e10 threshold: in graph mode
    # update_wf_norm is a separate session.run, run eagerly inside
    # build_opt_ops it would otherwise change the shift between psi and E_loc.
    wf.normalize_batch(wf(torch.from_numpy(configs_np)), max_value=1e2)
    out['psi'] = wf(torch.from_numpy(configs_np)).numpy()
  out['shift'] = np.float32(wf._exp_norm_shift.detach())

  # (2) two Metropolis steps, graph_builders.py:38-89, with recorded uniforms
  state = tf.get_variable('mc_state', initializer=configs_np.copy(), trainable=False)
  steps = []
  for _ in range(2):
    u_sites = torch.from_numpy(rng.random_sample((batch, n)).astype(np.float32))
    u_acc = torch.from_numpy(rng.random_sample((batch,)).astype(np.float32))
    tf._UNIFORM_QUEUE.extend([u_sites, u_acc])
    before = state.detach().clone().numpy()
    with torch.no_grad():
      mc_step, acc = graph_builders.build_monte_carlo_sampling(state, wf)
    steps.append((before, u_sites.numpy(), u_acc.numpy(),
                  state.detach().clone().numpy(), float(acc)))
  out['mc_before'] = np.stack([s[0] for s in steps])
  out['mc_u_sites'] = np.stack([s[1] for s in steps])
  out['mc_u_acc'] = np.stack([s[2] for s in steps])
  out['mc_after'] = np.stack([s[3] for s in steps])
  out['mc_accept_count'] = np.array([s[4] for s in steps], dtype=np.float32)

  # (3) Hamiltonian, operators.py:137-169 and 227-271
  ij, jx, jz = bonds_for(lattice, hp)
  out['bonds_ij'], out['bonds_jx'], out['bonds_jz'] = ij, jx, jz
  inputs = torch.from_numpy(configs_np)
  rec = Recorder(wf)
  with torch.no_grad():
    diag_terms, off_terms = [], []
    for k in range(len(ij)):
      bond = operators.HeisenbergBond((int(ij[k, 0]), int(ij[k, 1])),
                                      np.float32(jx[k]), np.float32(jz[k]))
      d, o = bond.build(rec, inputs)
      diag_terms.append(d.numpy())
      off_terms.append(o.numpy())
    psi = wf(inputs)
    out['flipped_configs'] = np.stack(rec.inputs, axis=1)      # [B, n_bonds, N]
    out['bond_diag'] = np.stack(diag_terms, axis=1)            # [B, n_bonds]
    out['bond_offdiag'] = np.stack(off_terms, axis=1)
    uniform = bool(np.all(jx == jx[0]) and np.all(jz == jz[0]))
    if uniform:
      ham = operators.HeisenbergHamiltonian(
          [(int(a), int(b)) for a, b in ij], np.float32(jx[0]), np.float32(jz[0]))
      diag, off = ham.build(wf, inputs)
      out['ham_diag'], out['ham_offdiag'] = diag.numpy(), off.numpy()
      out['local_energy'] = ham.local_value(wf, inputs).numpy()
      out['apply_in_place'] = ham.apply_in_place(wf, inputs).numpy()
    else:   # per-bond couplings: the reference expresses this as a sum of bonds
      diag = torch.from_numpy(np.sum(out['bond_diag'], axis=1))
      off = torch.from_numpy(np.sum(out['bond_offdiag'], axis=1))
      out['ham_diag'], out['ham_offdiag'] = diag.numpy(), off.numpy()
      out['local_energy'] = (diag + off / psi).numpy()
      out['apply_in_place'] = (diag * psi + off).numpy()

  # (4) energy gradient, training.py:531-586 (uniform couplings only: the
  # reference Hamiltonian class has one (j_x, j_z), operators.py:215-225)
  if uniform and name != 'conv2d_10x10':
    ham = operators.HeisenbergHamiltonian(
        [(int(a), int(b)) for a, b in ij], np.float32(jx[0]), np.float32(jz[0]))
    shared = {}
    ops = training.EnergyGradientOptimizer().build_opt_ops(
        wavefunction=wf, hamiltonian=ham, hparams=hp, shared_resources=shared)
    out['eg_configs'] = shared[graph_builders.ResourceName.CONFIGS].detach().numpy().copy()
    out['eg_gradient'] = torch.cat(
        [g.detach().reshape(-1) for g in ops.apply_gradients]).numpy()
    out['eg_mean_energy'] = np.float32(ops.metrics.detach())

  # (5) SWO loss + gradient, training.py:141-189.  Not for N >= 64: the
  # reference's own np.sqrt(2**n_sites) (training.py:170) raises TypeError once
  # 2**N no longer fits a 64-bit integer, so SWO cannot run there at all.
  if name != 'conv2d_10x10' and n < 64:
    target = wavefunctions.build_wavefunction(hp)
    target(dummy)
    tparams = oansatz.init_params(spec, seed=seed + 1000, bias_scale=0.1)
    for var, p in zip(target.get_trainable_variables(), tparams):
      tf.assign(var, p)
    # bring the target to the trainee's scale so the loss is O(1)
    with torch.no_grad():
      scale = float(torch.mean(torch.log(wf(dummy)) - torch.log(target(dummy))))
      tf.assign_add(target._exp_norm_shift, -scale + 0.5 * n * np.log(2.0))
    out['swo_target_params_flat'] = oansatz.flatten(tparams).numpy()
    out['swo_target_shift'] = np.float32(target._exp_norm_shift.detach())
    shared = {}
    ops = training.SupervisedWavefunctionOptimizer().build_opt_ops(
        wavefunction=wf, target_wavefunction=target, hparams=hp,
        shared_resources=shared)
    out['swo_configs'] = shared[graph_builders.ResourceName.CONFIGS].detach().numpy().copy()
    out['swo_loss'] = np.float32(ops.metrics.detach())
    out['swo_gradient'] = torch.cat(
        [torch.zeros_like(v).reshape(-1) if g is None else g.detach().reshape(-1)
         for g, v in zip(ops.apply_gradients, wf.get_trainable_variables())]).numpy()

  assert out['shift'] == np.float32(wf._exp_norm_shift.detach()), 'shift moved'
  out['spec_json'] = np.array(json.dumps(dict(
      kind=spec.kind, n_sites=spec.n_sites, num_layers=spec.num_layers,
      layer_size=spec.layer_size, num_filters=spec.num_filters,
      kernel_size=spec.kernel_size, size_x=spec.size_x, size_y=spec.size_y,
      nonlinearity=spec.nonlinearity)))
  return out


# Added after the first set was frozen (the reference's random_configurations
# is unseeded, so regenerating would change the committed files): generated by
# `python tests/golden/make_golden.py resnet` only.
RESNET_CASES = {
    'resnet1d_chain12_k3': (dict(wavefunction_type='res_net_1d', num_sites=12, num_resnet_blocks=2,
                                 num_conv_filters=4, kernel_size=3), 'chain'),
    'resnet1d_chain12_k4': (dict(wavefunction_type='res_net_1d', num_sites=12, num_resnet_blocks=1,
                                 num_conv_filters=3, kernel_size=4), 'chain'),
    'resnet2d_4x4_k3': (dict(wavefunction_type='res_net_2d', num_sites=16, size_x=4, size_y=4,
                             num_resnet_blocks=2, num_conv_filters=4, kernel_size=3), 'square'),
    'resnet2d_4x6_k2': (dict(wavefunction_type='res_net_2d', num_sites=24, size_x=4, size_y=6,
                             num_resnet_blocks=1, num_conv_filters=3, kernel_size=2), 'rect'),
}


# Round 2: the C3 network (10x10, 5 layers x 16 filters x 5x5) with the energy
# gradient and the SWO gradient recorded from the reference (the frozen
# conv2d_10x10 case skips both).  Uniform NN couplings, because the reference's
# HeisenbergHamiltonian -- which EnergyGradientOptimizer needs -- has a single
# (j_x, j_z) (operators.py:215-225).  `python tests/golden/make_golden.py c3grad`.
C3GRAD_CASES = {
    'conv2d_10x10_grad': (dict(wavefunction_type='conv_2d', num_sites=100, size_x=10, size_y=10,
                               num_conv_layers=5, num_conv_filters=16, kernel_size=5), 'square'),
}
BATCH['conv2d_10x10_grad'] = 2


def main():
  if len(sys.argv) > 1 and sys.argv[1] == 'c3grad':
    for k, (name, (overrides, lattice)) in enumerate(sorted(C3GRAD_CASES.items())):
      out = make_case(name, overrides, lattice, seed=900 + k)
      path = os.path.join(HERE, name + '.npz')
      np.savez_compressed(path, **out)
      print('%-22s %7.1f KB  keys=%d' % (name, os.path.getsize(path) / 1024.0, len(out)))
    return
  if len(sys.argv) > 1 and sys.argv[1] == 'resnet':
    for k, (name, (overrides, lattice)) in enumerate(sorted(RESNET_CASES.items())):
      out = make_case(name, overrides, lattice, seed=700 + k)
      path = os.path.join(HERE, name + '.npz')
      np.savez_compressed(path, **out)
      print('%-22s %7.1f KB  keys=%d' % (name, os.path.getsize(path) / 1024.0, len(out)))
    return
  for k, (name, (overrides, lattice)) in enumerate(sorted(CASES.items())):
    out = make_case(name, overrides, lattice, seed=100 + k)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-22s %7.1f KB  keys=%d' % (name, os.path.getsize(path) / 1024.0, len(out)))


if __name__ == '__main__':
  main()

"""Generates tests/golden/opt_*.npz: gradients of the remaining working
optimizers of the reference's training.py, recorded from the UNMODIFIED
reference modules on the eager TF/Sonnet stand-in (see make_golden.py).

    python tests/golden/make_golden_optimizers.py

Recorded per case:
  * LogOverlapSWO.build_opt_ops (training.py:298-379): overlap gradient
    sum_b O_b - sum_b r_b O_b / mean(r), r = psi_target / psi;
  * DualSamplingSWO.build_opt_ops (training.py:407-480): the two walker sets,
    loss = mean (psi - t)^2 and its gradient;
  * LogOverlapImaginaryTimeSWO.build_opt_ops (training.py:626-727): the
    supervisor copy's parameters (loaded into the reference's own deepcopy),
    beta, the overlap gradient with r = (psi_O - beta H psi_O) / psi and the
    supervisor energy mean(H psi_O / psi_O).
"""
import copy as _copy
import json
import os
import sys
import types

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
import make_golden as base             # noqa: E402

CASES = {
    'opt_rbm_4x4': (dict(wavefunction_type='rbm', num_sites=16, size_x=4, size_y=4,
                         num_fc_layers=0, fc_layer_size=24), 'square'),
    'opt_rbm_chain12_hidden': (dict(wavefunction_type='rbm', num_sites=12,
                                    num_fc_layers=1, fc_layer_size=10), 'chain'),
    'opt_fc_chain8': (dict(wavefunction_type='fully_connected', num_sites=8,
                           num_fc_layers=2, fc_layer_size=12), 'chain'),
    'opt_conv1d_chain12_k3': (dict(wavefunction_type='conv_1d', num_sites=12,
                                   num_conv_layers=2, num_conv_filters=4,
                                   kernel_size=3), 'chain'),
    'opt_conv2d_4x4_k2': (dict(wavefunction_type='conv_2d', num_sites=16, size_x=4,
                               size_y=4, num_conv_layers=2, num_conv_filters=3,
                               kernel_size=2), 'square'),
}
BATCH = 12
BETA = 0.05


def flat_grads(grads, variables):
  return torch.cat([torch.zeros_like(v).reshape(-1) if g is None else g.detach().reshape(-1)
                    for g, v in zip(grads, variables)]).numpy()


def fresh_walker_variables():
  """Each optimizer is built in its own graph in the reference's drivers; the
  eager stand-in keeps one variable store, so drop the walker variables
  between builds (tf.get_variable would otherwise hand back the old one)."""
  for key in list(tf._VARIABLES):
    if key.startswith('ResourceName.'):
      del tf._VARIABLES[key]
  return {}


def load_params(wf, spec, seed, dummy):
  wf(dummy)
  params = oansatz.init_params(spec, seed=seed, bias_scale=0.1)
  for var, p in zip(wf.get_trainable_variables(), params):
    tf.assign(var, p)
  return params


def make_case(name, overrides, lattice, seed):
  tf._reset_shim_state()
  torch.manual_seed(seed)
  hp = utils.create_hparams(batch_size=BATCH, time_evolution_beta=BETA, **overrides)
  n = hp.num_sites
  out = {}
  dummy = torch.from_numpy(utils.random_configurations(n, BATCH))
  wf, spec, params = base.build_reference_wavefunction(hp, seed, dummy)
  with torch.no_grad():
    wf.normalize_batch(wf(dummy), max_value=1e2)
  out['params_flat'] = oansatz.flatten(params).numpy()
  out['shift'] = np.float32(wf._exp_norm_shift.detach())

  target = wavefunctions.build_wavefunction(hp)
  tparams = load_params(target, spec, seed + 1000, dummy)
  with torch.no_grad():
    scale = float(torch.mean(torch.log(wf(dummy)) - torch.log(target(dummy))))
    tf.assign_add(target._exp_norm_shift, -scale)
  out['target_params_flat'] = oansatz.flatten(tparams).numpy()
  out['target_shift'] = np.float32(target._exp_norm_shift.detach())

  # ---- LogOverlapSWO, training.py:298-379 ----
  shared = fresh_walker_variables()
  ops = training.LogOverlapSWO().build_opt_ops(
      wavefunction=wf, target_wavefunction=target, hparams=hp, shared_resources=shared)
  out['lo_configs'] = shared[graph_builders.ResourceName.CONFIGS].detach().numpy().copy()
  out['lo_gradient'] = flat_grads(ops.apply_gradients, wf.get_trainable_variables())

  # ---- DualSamplingSWO, training.py:407-480 (target rescaled by sqrt(2^N)) ----
  with torch.no_grad():
    tf.assign_add(target._exp_norm_shift, 0.5 * n * np.log(2.0))
  out['ds_target_shift'] = np.float32(target._exp_norm_shift.detach())
  shared = fresh_walker_variables()
  u = [torch.rand(BATCH // 2, n), torch.rand(BATCH // 2), torch.rand(BATCH // 2, n),
       torch.rand(BATCH // 2)]
  tf._UNIFORM_QUEUE.extend(u)          # the two samplers run one eager step each
  ops = training.DualSamplingSWO().build_opt_ops(
      wavefunction=wf, target_wavefunction=target, hparams=hp, shared_resources=shared)
  assert not tf._UNIFORM_QUEUE
  out['ds_psi_configs'] = shared[graph_builders.ResourceName.CONFIGS].detach().numpy().copy()
  out['ds_target_configs'] = shared[
      graph_builders.ResourceName.TARGET_CONFIGS].detach().numpy().copy()
  out['ds_loss'] = np.float32(ops.metrics.detach())
  out['ds_gradient'] = flat_grads(ops.apply_gradients, wf.get_trainable_variables())

  # ---- LogOverlapImaginaryTimeSWO, training.py:626-727 ----
  ij, jx, jz = base.bonds_for(lattice, hp)
  ham = operators.HeisenbergHamiltonian(
      [(int(a), int(b)) for a, b in ij], np.float32(jx[0]), np.float32(jz[0]))
  out['bonds_ij'], out['bonds_jx'], out['bonds_jz'] = ij, jx, jz
  omega_params = {}

  def deepcopy_with_known_params(module):
    """The reference's own Wavefunction.__deepcopy__, then known parameters
    (the eager stand-in would otherwise evaluate the supervisor with its
    fresh random initialisation before update_supervisor runs)."""
    new = _copy.deepcopy(module)
    omega_params['params'] = load_params(new, spec, seed + 2000, dummy)
    omega_params['module'] = new
    return new

  real_copy = training.copy
  training.copy = types.SimpleNamespace(deepcopy=deepcopy_with_known_params)
  try:
    shared = fresh_walker_variables()
    ops = training.LogOverlapImaginaryTimeSWO().build_opt_ops(
        wavefunction=wf, hamiltonian=ham, hparams=hp, shared_resources=shared)
  finally:
    training.copy = real_copy
  out['it_configs'] = shared[graph_builders.ResourceName.CONFIGS].detach().numpy().copy()
  out['it_beta'] = np.float32(BETA)
  out['it_omega_params_flat'] = oansatz.flatten(omega_params['params']).numpy()
  out['it_omega_shift'] = np.float32(omega_params['module']._exp_norm_shift.detach())
  out['it_gradient'] = flat_grads(ops.apply_gradients, wf.get_trainable_variables())
  out['it_energy'] = np.float32(ops.energy.detach())
  # update_supervisor ran eagerly at the end of build_opt_ops: the copy now
  # holds the trainee's parameters (module_transfer_ops, wavefunctions.py:300-325)
  after = oansatz.flatten([v.detach() for v in
                           omega_params['module'].get_trainable_variables()]).numpy()
  assert np.array_equal(after, out['params_flat'])

  assert out['shift'] == np.float32(wf._exp_norm_shift.detach()), 'shift moved'
  out['spec_json'] = np.array(json.dumps(dict(
      kind=spec.kind, n_sites=spec.n_sites, num_layers=spec.num_layers,
      layer_size=spec.layer_size, num_filters=spec.num_filters,
      kernel_size=spec.kernel_size, size_x=spec.size_x, size_y=spec.size_y,
      nonlinearity=spec.nonlinearity)))
  return out


def main():
  for k, (name, (overrides, lattice)) in enumerate(sorted(CASES.items())):
    out = make_case(name, overrides, lattice, seed=300 + k)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-26s %7.1f KB  keys=%d' % (name, os.path.getsize(path) / 1024.0, len(out)))


if __name__ == '__main__':
  main()

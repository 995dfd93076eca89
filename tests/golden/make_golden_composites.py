"""Generates tests/golden/cmp_*.npz: signed output activations and the sum /
difference / product composites (wavefunctions.py:61-165, 350-353, 1178-1194),
recorded from the UNMODIFIED reference modules on the eager TF/Sonnet stand-in
(see make_golden.py).

    python tests/golden/make_golden_composites.py

Recorded per case: the leaves' parameters (get_trainable_variables order) and
exp_norm_shift values, configurations, psi, HeisenbergHamiltonian.local_value /
apply_in_place, and the EnergyGradientOptimizer gradient + mean energy
(training.py:531-586) over all trainable variables.
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

N = 8
BATCH = 12
FC = dict(num_fc_layers=2, fc_layer_size=6)
CONV = dict(num_conv_layers=2, num_conv_filters=3, kernel_size=3)
CASES = {
    # name: hparams overrides
    'cmp_fc_tanh': dict(wavefunction_type='fully_connected', output_activation='tanh', **FC),
    'cmp_conv1d_identity': dict(wavefunction_type='conv_1d', output_activation='identity', **CONV),
    'cmp_sum_rbm_fc_tanh': dict(wavefunction_type='sum', composite_wavefunction_types=('rbm', 'fully_connected'),
                                composite_output_activations=('exp', 'tanh'), **FC),
    'cmp_diff_fc_conv1d_tanh': dict(wavefunction_type='diff',
                                    composite_wavefunction_types=('fully_connected', 'conv_1d'),
                                    composite_output_activations=('exp', 'tanh'), **FC, **CONV),
    'cmp_prod_rbm_fc_cos': dict(wavefunction_type='prod', composite_wavefunction_types=('rbm', 'fully_connected'),
                                composite_output_activations=('exp', 'cos'), **FC),
}


def leaves_of(wf):
  subs = getattr(wf, '_sub_wavefunctions', [])
  if not subs:
    return [wf]
  out = []
  for sub in subs:
    out += leaves_of(sub)
  return out


def leaf_spec(leaf, hp):
  kind = {'FullyConnectedNetwork': 'fully_connected', 'RestrictedBoltzmannNetwork': 'rbm',
          'Conv1DNetwork': 'conv_1d'}[type(leaf).__name__]
  if kind in ('fully_connected', 'rbm'):
    return oansatz.AnsatzSpec(kind, N, num_layers=hp.num_fc_layers, layer_size=hp.fc_layer_size)
  return oansatz.AnsatzSpec(kind, N, num_layers=hp.num_conv_layers, num_filters=hp.num_conv_filters,
                            kernel_size=hp.kernel_size)


def make_case(name, overrides, seed):
  tf._reset_shim_state()
  torch.manual_seed(seed)
  hp = utils.create_hparams(batch_size=BATCH, num_sites=N, **overrides)
  configs_np = utils.random_configurations(N, BATCH)
  dummy = torch.from_numpy(configs_np)
  wf = wavefunctions.build_wavefunction(hp)
  wf(dummy)                                           # creates the variables
  leaves = leaves_of(wf)
  out, flats, specs, shifts = {}, [], [], []
  for k, leaf in enumerate(leaves):
    spec = leaf_spec(leaf, hp)
    params = oansatz.init_params(spec, seed=seed + 10 * k, bias_scale=0.2)
    variables = leaf.get_trainable_variables()
    assert len(variables) == len(params)
    for var, p in zip(variables, params):
      assert tuple(var.shape) == tuple(p.shape), (var.name, var.shape, p.shape)
      tf.assign(var, p)
    flats.append(oansatz.flatten(params).numpy())
    specs.append(dict(kind=spec.kind, n_sites=spec.n_sites, num_layers=spec.num_layers,
                      layer_size=spec.layer_size, num_filters=spec.num_filters,
                      kernel_size=spec.kernel_size, size_x=spec.size_x, size_y=spec.size_y,
                      nonlinearity=spec.nonlinearity))
    shift = getattr(leaf, '_exp_norm_shift', None)
    if shift is not None:                             # exp leaf: bring psi to O(1)
      with torch.no_grad():
        leaf.normalize_batch(leaf(dummy), max_value=2.0)
      shifts.append(float(leaf._exp_norm_shift.detach()))
    else:
      shifts.append(float('nan'))
  assert [v is w for v, w in zip(wf.get_trainable_variables(),
                                 sum([l.get_trainable_variables() for l in leaves], []))]
  out['leaf_params_flat'] = np.concatenate(flats)
  out['leaf_sizes'] = np.array([f.size for f in flats], dtype=np.int64)
  out['leaf_shifts'] = np.array(shifts, dtype=np.float64)
  out['configs'] = configs_np
  with torch.no_grad():
    out['psi'] = wf(dummy).numpy()
  ij, jx, jz = lattices.heisenberg_couplings(lattices.chain_bonds(N), -1.0, 1.0)
  out['bonds_ij'], out['bonds_jx'], out['bonds_jz'] = ij, jx, jz
  ham = operators.HeisenbergHamiltonian([(int(a), int(b)) for a, b in ij], np.float32(-1.0), np.float32(1.0))
  with torch.no_grad():
    out['local_energy'] = ham.local_value(wf, dummy).numpy()
    out['apply_in_place'] = ham.apply_in_place(wf, dummy).numpy()
  shared = {}
  ops = training.EnergyGradientOptimizer().build_opt_ops(
      wavefunction=wf, hamiltonian=ham, hparams=hp, shared_resources=shared)
  out['eg_configs'] = shared[graph_builders.ResourceName.CONFIGS].detach().numpy().copy()
  out['eg_gradient'] = torch.cat([
      torch.zeros_like(v).reshape(-1) if g is None else g.detach().reshape(-1)
      for g, v in zip(ops.apply_gradients, wf.get_trainable_variables())]).numpy()
  out['eg_mean_energy'] = np.float32(ops.metrics.detach())
  out['case_json'] = np.array(json.dumps(dict(hparams={k: (list(v) if isinstance(v, tuple) else v)
                                                       for k, v in overrides.items()},
                                              leaf_specs=specs)))
  return out


def main():
  for k, (name, overrides) in enumerate(sorted(CASES.items())):
    out = make_case(name, overrides, seed=500 + k)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-28s %6.1f KB  psi[:3]=%s' % (name, os.path.getsize(path) / 1024.0, out['psi'][:3]))


if __name__ == '__main__':
  main()

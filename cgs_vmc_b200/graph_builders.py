"""Mirror of the reference's graph_builders.py on the CUDA library.

Same names and argument orders (graph_builders.py:16-151); tensors are torch
CUDA tensors and "ops" are `session.Op` objects executed by `Session.run`.
"""
import enum

import torch

from . import _native, engine
from .session import Op


class ResourceName(enum.Enum):
  """Type of sharable resources, serves as key in `shared_resources`
  (graph_builders.py:16-22)."""
  CONFIGS = 'CONFIGS'
  TARGET_CONFIGS = 'TARGET_CONFIGS'
  TARGET_PSI = 'TARGET_PSI'
  TRAINING_PSI = 'TRAINING_PSI'
  MONTE_CARLO_SAMPLING = 'MONTE_CARLO_SAMPLING'


class ConfigsVariable:
  """The non-trainable [batch_size, n_sites] variable of get_configs
  (graph_builders.py:119-122).  The walkers live bit-packed on the device
  (`state.packed`); `value()` materialises the reference's float32 +-1 view."""

  def __init__(self, state):
    self.state = state

  @property
  def shape(self):
    return (self.state.batch_size, self.state.n_sites)

  @property
  def packed(self):
    return self.state.packed

  def value(self):
    return self.state.configs()

  def assign(self, configs):
    self.state.set_configs(torch.as_tensor(configs))
    return self


def as_packed(inputs, n_sites=None):
  """Accepts a ConfigsVariable or a float [B, N] tensor of +-1."""
  if isinstance(inputs, ConfigsVariable):
    return inputs.packed
  inputs = torch.as_tensor(inputs)
  if inputs.dim() != 2:
    raise ValueError('inputs must have shape (batch, num_sites)')
  if n_sites is not None and inputs.shape[1] != n_sites:
    raise ValueError('Input tensor has wrong shape.')
  _native.require_cuda()
  return _native.pack_configs(inputs.to('cuda', torch.float32).contiguous())


_NUM_EPOCHS = {'value': None}


def get_or_create_num_epochs():
  """graph_builders.py:25-35: a process-wide int32 counter."""
  if _NUM_EPOCHS['value'] is None:
    _NUM_EPOCHS['value'] = torch.zeros((), dtype=torch.int32)
  return _NUM_EPOCHS['value']


def reset_num_epochs():
  _NUM_EPOCHS['value'] = None


def build_monte_carlo_sampling(inputs, wavefunction, psi=None):
  """graph_builders.py:38-89.  Returns (mc_step, acceptance_count) ops.

  `session.run(mc_step)` performs ONE Metropolis exchange step for every
  walker like the reference; `mc_step(n_steps=k)` runs k steps in a single
  persistent-kernel launch.  `psi` is accepted for signature compatibility
  (the kernels cache the current amplitude themselves)."""
  del psi
  if not isinstance(inputs, ConfigsVariable):
    raise ValueError('inputs must be the variable returned by get_configs')
  state = inputs.state
  last = {'before': 0}

  def generic_steps(n_steps):
    """Signed / composite amplitudes: propose on the device, evaluate
    (log|psi|, sign) of the proposals through the parts' kernels, accept."""
    n = state.n_sites
    wavefunction.connect(n)
    logabs, sign = wavefunction.amplitudes(state.packed)
    logabs, sign = logabs.float().contiguous().clone(), sign.float().contiguous().clone()
    for _ in range(int(n_steps)):
      proposed, u_acc = _native.propose_exchange(state.packed, n, state.seed, state.walker_id0,
                                                 state.step)
      la_new, s_new = wavefunction.amplitudes(proposed)
      _native.accept_exchange(state.packed, proposed, n, logabs, sign, la_new.float().contiguous(),
                              s_new.float().contiguous(), u_acc, state.accept_count)
      state.step += 1
    state.proposed += int(n_steps) * state.batch_size

  def mc_step(n_steps=1):
    last['before'] = None
    last['mark'] = state.accept_count.clone()
    if wavefunction.fast_path:
      state.mc_steps(wavefunction.native(state.n_sites), n_steps)
    else:
      generic_steps(n_steps)
    return inputs

  def acceptance_count():
    if last.get('mark') is None:
      return 0.0
    return float((state.accept_count - last['mark']).item())

  return Op(mc_step, 'mc_step'), Op(acceptance_count, 'acceptance_count')


def get_configs(shared_resources, batch_size, n_sites, include=True,
                configs_id=ResourceName.CONFIGS, seed=0xC65, walker_id0=0):
  """graph_builders.py:92-125 (same ValueError on a shape mismatch)."""
  if configs_id in shared_resources:
    configs = shared_resources[configs_id]
    if list(configs.shape) != [batch_size, n_sites]:
      raise ValueError('Size of existing variable does not match.')
    return configs
  _native.require_cuda()
  state = engine.WalkerState(batch_size, n_sites, seed=seed, walker_id0=walker_id0)
  configs = ConfigsVariable(state)
  if include:
    shared_resources[configs_id] = configs
  return configs


def get_monte_carlo_sampling(shared_resources, inputs, wavefunction, include=True):
  """graph_builders.py:128-151."""
  if ResourceName.MONTE_CARLO_SAMPLING in shared_resources:
    return shared_resources[ResourceName.MONTE_CARLO_SAMPLING]
  mc_step, acc_rate = build_monte_carlo_sampling(inputs, wavefunction)
  if include:
    shared_resources[ResourceName.MONTE_CARLO_SAMPLING] = (mc_step, acc_rate)
  return mc_step, acc_rate

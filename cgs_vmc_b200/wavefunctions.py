"""Mirror of the reference's wavefunctions.py for the in-scope ansaetze
(fully_connected, rbm, conv_1d, conv_2d) on the CUDA library.

Same class names, constructor arguments, `from_hparams`, registry and error
behaviour (wavefunctions.py:21-615, 1157-1211).  Amplitudes are evaluated in
the log domain on the device; `__call__` returns the reference's float32
psi = exp(z - exp_norm_shift) (wavefunctions.py:206-232).
"""
import copy
import math

import numpy as np
import torch

from . import _native, layers
from .session import Op

_UNBUILT = ('res_net_1d', 'res_net_2d', 'mps', 'pbdg', 'fully_connected_nnb',
            'ed_vector', 'gnn')


def _sonnet_init(shapes, generator):
  """Sonnet v1 defaults: weights truncated normal (+-2 sigma), sigma =
  1/sqrt(fan_in); biases zero."""
  out = []
  for shape in shapes:
    if len(shape) == 1:
      out.append(torch.zeros(shape))
    else:
      fan_in = int(np.prod(shape[:-1]))
      t = torch.empty(shape)
      s = 1.0 / math.sqrt(fan_in)
      torch.nn.init.trunc_normal_(t, 0.0, s, -2 * s, 2 * s, generator=generator)
      out.append(t)
  return out


class Wavefunction:
  """Wavefunction interface (wavefunctions.py:21-297)."""

  _kind = None

  def __init__(self, name='wavefunction'):
    self._name = name
    self._unique_name = name
    self._sub_wavefunctions = []
    self._exp_norm_shift = None
    self._native = None
    self._n_sites = None
    self._seed = None
    self._variables = []

  # ---- to be provided by subclasses ------------------------------------
  def _native_args(self, n_sites):
    raise NotImplementedError

  def _param_shapes(self, n_sites):
    raise NotImplementedError

  def _build(self, inputs):
    raise NotImplementedError

  # ---- device handle -----------------------------------------------------
  def native(self, n_sites=None):
    """The cgsvmc ansatz handle; created (with Sonnet-default parameters) the
    first time the number of sites is known, like Sonnet creates variables at
    first connection."""
    if self._native is None:
      if n_sites is None:
        n_sites = self._n_sites
      if n_sites is None:
        raise ValueError('wavefunction has not been connected to inputs yet')
      self._n_sites = int(n_sites)
      self._native = _native.Ansatz(**self._native_args(self._n_sites))
      gen = torch.Generator()
      if self._seed is not None:
        gen.manual_seed(int(self._seed))
      shapes = self._param_shapes(self._n_sites)
      init = _sonnet_init(shapes, gen)
      self._native.set_params(torch.cat([t.reshape(-1) for t in init]))
      self._variables, off = [], 0
      for shape in shapes:
        n = int(np.prod(shape))
        self._variables.append(self._native.params[off:off + n].view(*shape))
        off += n
      self._exp_norm_shift = -10.0            # wavefunctions.py:209
    elif n_sites is not None and int(n_sites) != self._n_sites:
      raise ValueError('Input tensor has wrong shape.')
    return self._native

  def seed(self, value):
    """Seeds the parameter initialisation (the reference is unseeded)."""
    self._seed = value
    return self

  # ---- evaluation ----------------------------------------------------------
  def log_amplitude(self, inputs):
    """z - exp_norm_shift = log psi, float32 [B]."""
    from . import graph_builders
    n = inputs.shape[1]
    a = self.native(n)
    packed = graph_builders.as_packed(inputs, n)
    return a.log_amp(packed) - self._exp_norm_shift

  def __call__(self, inputs):
    return self._build(inputs)

  def get_trainable_variables(self):
    """wavefunctions.py:167-175: list of parameter tensors (views into the
    flat device buffer, layout of include/cgsvmc.h)."""
    if self._native is None:
      self.native()
    out = list(self._variables)
    for sub in self._sub_wavefunctions:
      out += sub.get_trainable_variables()
    return out

  @property
  def flat_parameters(self):
    return self.native().params

  def __deepcopy__(self, memo):
    """wavefunctions.py:177-204: a structurally identical module with its own
    (freshly initialised) variables, named 'dc_<name>'."""
    if id(self) in memo:
      return memo[id(self)]
    args = dict(self._init_args)
    args['name'] = 'dc_{}'.format(self._unique_name)
    new = type(self)(**args)
    new._n_sites = self._n_sites
    memo[id(self)] = new
    return new

  # ---- normalisation (wavefunctions.py:206-288) ---------------------------
  def normalize_batch(self, batch_of_amplitudes, max_value=1e10):
    if self._exp_norm_shift is None and self._native is None:
      return None

    def run():
      log_max = float(torch.log(torch.as_tensor(batch_of_amplitudes()
                                if callable(batch_of_amplitudes)
                                else batch_of_amplitudes).max()))
      self._exp_norm_shift += log_max - math.log(max_value)
      return self._exp_norm_shift
    return Op(run, 'normalize_batch')

  def update_norm(self, batch_of_amplitudes, max_value=1e10):
    """Op that raises exp_norm_shift iff max psi > max_value
    (wavefunctions.py:261-288).  `batch_of_amplitudes` may be a tensor or a
    callable returning the current amplitudes (graph semantics)."""
    def run():
      amps = batch_of_amplitudes() if callable(batch_of_amplitudes) else batch_of_amplitudes
      log_max = float(torch.log(torch.as_tensor(amps).max()))
      max_log = math.log(max_value)
      if log_max > max_log:
        self._exp_norm_shift += log_max - max_log
      return self._exp_norm_shift
    return Op(run, 'update_norm')

  def __add__(self, other):
    raise NotImplementedError('sum / difference / product wavefunctions need signed '
                              'amplitudes: not built in the CUDA path (SURVEY.md 8(f) rank 3)')
  __sub__ = __add__
  __mul__ = __add__

  @classmethod
  def from_hparams(cls, hparams, name=''):
    raise NotImplementedError


def module_transfer_ops(source_module, target_module):
  """wavefunctions.py:300-325: op copying every variable of source to target."""
  def run():
    src, dst = source_module.native(), target_module.native(source_module._n_sites)
    if src.num_params != dst.num_params:
      raise ValueError('`target_module` does not have the same structure as source.')
    dst.params.copy_(src.params)
  return Op(run, 'module_transfer')


def _check_exp(output_activation):
  if output_activation not in ('exp', None):
    raise NotImplementedError(
        'output_activation=%r: only exp (positive amplitudes, log domain) is built in '
        'the CUDA path (SURVEY.md 8(f) rank 3)' % (output_activation,))


class _ExpAnsatz(Wavefunction):
  def _build(self, inputs):
    return torch.exp(self.log_amplitude(inputs))


class FullyConnectedNetwork(_ExpAnsatz):
  """wavefunctions.py:328-388."""
  _kind = 'fully_connected'

  def __init__(self, num_layers, layer_size, nonlinearity='relu', output_activation='exp',
               name='fully_connected_network'):
    super().__init__(name=name)
    _check_exp(output_activation)
    self._num_layers, self._layer_size = num_layers, layer_size
    self._nonlinearity, self._output_activation = nonlinearity, output_activation
    self._init_args = dict(num_layers=num_layers, layer_size=layer_size,
                           nonlinearity=nonlinearity, output_activation=output_activation)

  def _native_args(self, n):
    return dict(kind='fully_connected', n_sites=n, num_layers=self._num_layers,
                layer_size=self._layer_size, nonlinearity=self._nonlinearity)

  def _param_shapes(self, n):
    shapes, n_in = [], n
    for _ in range(self._num_layers):
      shapes += [(n_in, self._layer_size), (self._layer_size,)]
      n_in = self._layer_size
    return shapes + [(n_in, 1), (1,)]

  @classmethod
  def from_hparams(cls, hparams, name=''):
    params = dict(num_layers=hparams.num_fc_layers, layer_size=hparams.fc_layer_size,
                  output_activation=layers.NONLINEARITIES[hparams.output_activation],
                  nonlinearity=layers.NONLINEARITIES[hparams.nonlinearity])
    if name:
      params['name'] = name
    wf = cls(**params)
    wf._n_sites = hparams.num_sites
    return wf


class RestrictedBoltzmannNetwork(_ExpAnsatz):
  """wavefunctions.py:391-452."""
  _kind = 'rbm'

  def __init__(self, num_layers, layer_size, nonlinearity='relu',
               name='restricted_boltzmann_network'):
    super().__init__(name=name)
    self._num_layers, self._layer_size, self._nonlinearity = num_layers, layer_size, nonlinearity
    self._init_args = dict(num_layers=num_layers, layer_size=layer_size, nonlinearity=nonlinearity)

  def _native_args(self, n):
    return dict(kind='rbm', n_sites=n, num_layers=self._num_layers,
                layer_size=self._layer_size, nonlinearity=self._nonlinearity)

  def _param_shapes(self, n):
    shapes, n_in = [(n, 1), (1,)], n
    for _ in range(self._num_layers):
      shapes += [(n_in, self._layer_size), (self._layer_size,)]
      n_in = self._layer_size
    return shapes + [(n_in, self._layer_size), (self._layer_size,)]

  @classmethod
  def from_hparams(cls, hparams, name=''):
    params = dict(num_layers=hparams.num_fc_layers, layer_size=hparams.fc_layer_size,
                  nonlinearity=layers.NONLINEARITIES[hparams.nonlinearity])
    if name:
      params['name'] = name
    wf = cls(**params)
    wf._n_sites = hparams.num_sites
    return wf


class Conv1DNetwork(_ExpAnsatz):
  """wavefunctions.py:454-528."""
  _kind = 'conv_1d'

  def __init__(self, num_layers, num_filters, kernel_size, nonlinearity='relu',
               output_activation='exp', name='conv_1d_network'):
    super().__init__(name=name)
    _check_exp(output_activation)
    self._num_layers, self._num_filters, self._kernel_size = num_layers, num_filters, kernel_size
    self._nonlinearity, self._output_activation = nonlinearity, output_activation
    self._components = [layers.Conv1dPeriodic(num_filters, kernel_size) for _ in range(num_layers)]
    self._init_args = dict(num_layers=num_layers, num_filters=num_filters, kernel_size=kernel_size,
                           nonlinearity=nonlinearity, output_activation=output_activation)

  def _native_args(self, n):
    return dict(kind='conv_1d', n_sites=n, num_layers=self._num_layers,
                num_filters=self._num_filters, kernel_size=self._kernel_size,
                nonlinearity=self._nonlinearity)

  def _param_shapes(self, n):
    shapes, c_in = [], 1
    for _ in range(self._num_layers):
      shapes += [(self._kernel_size, c_in, self._num_filters), (self._num_filters,)]
      c_in = self._num_filters
    return shapes

  @classmethod
  def from_hparams(cls, hparams, name=''):
    params = dict(num_layers=hparams.num_conv_layers, num_filters=hparams.num_conv_filters,
                  kernel_size=hparams.kernel_size,
                  output_activation=layers.NONLINEARITIES[hparams.output_activation],
                  nonlinearity=layers.NONLINEARITIES[hparams.nonlinearity])
    if name:
      params['name'] = name
    wf = cls(**params)
    wf._n_sites = hparams.num_sites
    return wf


class Conv2DNetwork(_ExpAnsatz):
  """wavefunctions.py:531-615."""
  _kind = 'conv_2d'

  def __init__(self, num_layers, num_filters, kernel_size, size_x, size_y, nonlinearity='relu',
               output_activation='exp', name='conv_2d_network'):
    super().__init__(name=name)
    _check_exp(output_activation)
    self._num_layers, self._num_filters, self._kernel_size = num_layers, num_filters, kernel_size
    self._size_x, self._size_y = size_x, size_y
    self._nonlinearity, self._output_activation = nonlinearity, output_activation
    self._components = [layers.Conv2dPeriodic(num_filters, kernel_size) for _ in range(num_layers)]
    self._n_sites = size_x * size_y
    self._init_args = dict(num_layers=num_layers, num_filters=num_filters, kernel_size=kernel_size,
                           size_x=size_x, size_y=size_y, nonlinearity=nonlinearity,
                           output_activation=output_activation)

  def _native_args(self, n):
    return dict(kind='conv_2d', n_sites=n, num_layers=self._num_layers,
                num_filters=self._num_filters, kernel_size=self._kernel_size,
                size_x=self._size_x, size_y=self._size_y, nonlinearity=self._nonlinearity)

  def _param_shapes(self, n):
    shapes, c_in, k = [], 1, self._kernel_size
    for _ in range(self._num_layers):
      shapes += [(k, k, c_in, self._num_filters), (self._num_filters,)]
      c_in = self._num_filters
    return shapes

  @classmethod
  def from_hparams(cls, hparams, name=''):
    params = dict(num_layers=hparams.num_conv_layers, num_filters=hparams.num_conv_filters,
                  kernel_size=hparams.kernel_size, size_x=hparams.size_x, size_y=hparams.size_y,
                  output_activation=layers.NONLINEARITIES[hparams.output_activation],
                  nonlinearity=layers.NONLINEARITIES[hparams.nonlinearity])
    if name:
      params['name'] = name
    return cls(**params)


def build_wavefunction(hparams):
  """wavefunctions.py:1157-1196."""
  wavefunction_type = hparams.wavefunction_type
  if wavefunction_type in WAVEFUNCTION_TYPES:
    return WAVEFUNCTION_TYPES[wavefunction_type].from_hparams(hparams)
  if wavefunction_type in _UNBUILT or wavefunction_type in ('sum', 'diff', 'prod'):
    raise NotImplementedError(
        'wavefunction_type=%r exists in the reference but is outside the CUDA hot path '
        '(SURVEY.md section 2 rows 12-17)' % wavefunction_type)
  raise ValueError('Provided wavefunction_type is not registered.')


WAVEFUNCTION_TYPES = {
    'fully_connected': FullyConnectedNetwork,
    'rbm': RestrictedBoltzmannNetwork,
    'conv_1d': Conv1DNetwork,
    'conv_2d': Conv2DNetwork,
}

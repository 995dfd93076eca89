"""Mirror of the reference's wavefunctions.py for the in-scope ansaetze
(fully_connected, rbm, conv_1d, conv_2d, res_net_1d, res_net_2d) on the CUDA
library, including the
signed output activations (layers.py:13-21) and the sum / difference / product
composites (wavefunctions.py:61-165, 1178-1194).

Same class names, constructor arguments, `from_hparams`, registry and error
behaviour (wavefunctions.py:21-615, 1157-1211).  Amplitudes are carried as
(log|psi|, sign psi) on the device; `__call__` returns the reference's float32
psi (= exp(z - exp_norm_shift) for output_activation = exp,
wavefunctions.py:206-232).

Two evaluation routes: a wavefunction with `fast_path` (one ansatz, exp output)
runs on the fused sampler / local-energy / estimator kernels; everything else
goes through `amplitudes()` -- cgsvmc_log_amp of the parts, combined here -- and
the amplitude-agnostic entry points cgsvmc_propose_exchange /
cgsvmc_accept_exchange / cgsvmc_local_energy_from_amps.
"""
import copy
import math

import numpy as np
import torch

from . import _native, layers
from .session import Op

_UNBUILT = ('mps', 'pbdg', 'fully_connected_nnb', 'ed_vector', 'gnn')

# The reference initialises every module from TensorFlow's unseeded random
# stream: two wavefunctions never start with the same parameters.  Here an
# unseeded wavefunction draws its initial parameters from a stream keyed by
# (torch.initial_seed(), number of wavefunctions initialised so far in this
# process): distinct per instance, identical on every rank of a sharded run
# (all ranks build the same wavefunctions in the same order; the optimizers
# additionally broadcast rank 0's parameters, training._Model).
_INIT_COUNTER = [0]


def _next_init_seed():
  _INIT_COUNTER[0] += 1
  mixed = (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + _INIT_COUNTER[0] * 0xD1B54A32D192ED03)
  return mixed & 0x7FFFFFFFFFFFFFFF


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
      gen.manual_seed(int(self._seed) if self._seed is not None else _next_init_seed())
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

  # ---- (log|psi|, sign) interface used by the amplitude-agnostic route ---------
  fast_path = False        # one native ansatz with exp output: fused kernels apply

  def connect(self, n_sites):
    """Creates the device parameters for `n_sites` (Sonnet creates variables
    at first connection) and returns self."""
    raise NotImplementedError

  def leaves(self):
    """The parameterised ansaetze of this wavefunction, in
    get_trainable_variables order."""
    raise NotImplementedError

  def amplitudes(self, packed):
    """(log|psi| float32 [B], sign float32 [B] of +-1 (0 where psi = 0))."""
    raise NotImplementedError

  def weighted_grad_sum(self, packed, weights):
    """[K, P_total] = sum_b weights[k, b] d log psi_b / d params, concatenated
    over leaves()."""
    raise NotImplementedError

  @property
  def num_params(self):
    return sum(leaf.native().num_params for leaf in self.leaves())

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
  def _global_log_max(self, batch_of_amplitudes, log_amplitudes):
    """log of the largest amplitude over the WHOLE batch: tf.reduce_max of
    wavefunctions.py:255 / 283 runs over every walker, so under walker
    sharding the local maximum is all-reduced (MAX) -- every rank then applies
    the same shift.  With `log_amplitudes` (a tensor or callable giving
    log psi = z - shift) the maximum is taken in the log domain, where
    exp overflow cannot turn the shift into inf."""
    from . import distributed
    if log_amplitudes is not None:
      logs = log_amplitudes() if callable(log_amplitudes) else log_amplitudes
      log_max = torch.as_tensor(logs).max().reshape(1).double()
    else:
      amps = batch_of_amplitudes() if callable(batch_of_amplitudes) else batch_of_amplitudes
      log_max = torch.log(torch.as_tensor(amps).max()).reshape(1).double()
    distributed.allreduce_(log_max, op='max')
    return float(log_max.item())

  def normalize_batch(self, batch_of_amplitudes, max_value=1e10, log_amplitudes=None):
    """Op mapping the batch onto (0, max_value] (wavefunctions.py:234-257)."""
    if self._exp_norm_shift is None and self._native is None:
      return None

    def run():
      log_max = self._global_log_max(batch_of_amplitudes, log_amplitudes)
      self._exp_norm_shift += log_max - math.log(max_value)
      return self._exp_norm_shift
    return Op(run, 'normalize_batch')

  def update_norm(self, batch_of_amplitudes, max_value=1e10, log_amplitudes=None):
    """Op that raises exp_norm_shift iff max psi > max_value
    (wavefunctions.py:261-288).  `batch_of_amplitudes` may be a tensor or a
    callable returning the current amplitudes (graph semantics);
    `log_amplitudes` optionally gives log psi of the same batch."""
    if not self.fast_path:          # no exp_norm_shift without the exp output
      return None

    def run():
      log_max = self._global_log_max(batch_of_amplitudes, log_amplitudes)
      max_log = math.log(max_value)
      if log_max > max_log:
        self._exp_norm_shift += log_max - max_log
      return self._exp_norm_shift
    return Op(run, 'update_norm')

  def __add__(self, other):
    """wavefunctions.py:61-99."""
    return SumOfWavefunctions(self, other)

  def __mul__(self, other):
    """wavefunctions.py:101-161: other is a Wavefunction or a float."""
    return ProductOfWavefunctions(self, other)

  def __sub__(self, other):
    """wavefunctions.py:163-165."""
    return self.__add__(other * -1.)

  @classmethod
  def from_hparams(cls, hparams, name=''):
    raise NotImplementedError


def module_transfer_ops(source_module, target_module):
  """wavefunctions.py:300-325: op copying every variable of source to target."""
  def run():
    target_module.connect(source_module._n_sites)
    src, dst = source_module.leaves(), target_module.leaves()
    if len(src) != len(dst) or any(a.native().num_params != b.native().num_params
                                   for a, b in zip(src, dst)):
      raise ValueError('`target_module` does not have the same structure as source.')
    for a, b in zip(src, dst):
      b.native().params.copy_(a.native().params)
      # The reference copies the trainable variables only and leaves the copy's
      # exp_norm_shift at its initial -10.  psi_target / psi then carries the
      # constant factor e^(shift - shift_copy), which cancels in the log-overlap
      # gradient; copying the shift as well keeps that ratio O(1) in float32.
      b._exp_norm_shift = a._exp_norm_shift
  return Op(run, 'module_transfer')


def _check_exp(output_activation):
  if output_activation not in layers.NONLINEARITIES:
    raise ValueError('unknown output_activation %r' % (output_activation,))


def _output_value_and_slope(name, z):
  """f(z) and f'(z) for the output activations of layers.py:13-21."""
  if name == 'identity':
    return z, torch.ones_like(z)
  if name == 'tanh':
    v = torch.tanh(z)
    return v, 1.0 - v * v
  if name == 'sigmoid':
    v = torch.sigmoid(z)
    return v, v * (1.0 - v)
  if name == 'relu':
    return torch.relu(z), (z > 0).to(z.dtype)
  if name == 'cos':
    return torch.cos(z), -torch.sin(z)
  if name == 'tan':
    v = torch.tan(z)
    return v, 1.0 + v * v
  raise ValueError('unknown output_activation %r' % (name,))


class _ExpAnsatz(Wavefunction):
  """One native ansatz z(sigma) followed by the output activation:
  psi = exp(z - exp_norm_shift) (wavefunctions.py:206-232, 350-353) or
  psi = f(z) without a shift for the other activations."""

  @property
  def fast_path(self):
    return getattr(self, '_output_activation', 'exp') in ('exp', None)

  def connect(self, n_sites):
    self.native(n_sites)
    return self

  def leaves(self):
    return [self]

  def _z(self, packed):
    return self.native().log_amp(packed)

  def amplitudes(self, packed):
    z = self._z(packed)
    if self.fast_path:
      return z - self._exp_norm_shift, torch.ones_like(z)
    v, _ = _output_value_and_slope(self._output_activation, z)
    return torch.log(torch.abs(v)), torch.sign(v)

  def weighted_grad_sum(self, packed, weights):
    weights = weights.reshape(-1, packed.shape[0])
    if not self.fast_path:        # d log psi = f'(z) / f(z) dz
      v, dv = _output_value_and_slope(self._output_activation, self._z(packed))
      weights = weights * (dv / v)
    return self.native().weighted_grad_sum(packed, weights.float().contiguous())

  def _build(self, inputs):
    if self.fast_path:
      return torch.exp(self.log_amplitude(inputs))
    from . import graph_builders
    n = inputs.shape[1]
    self.native(n)
    v, _ = _output_value_and_slope(self._output_activation,
                                   self._z(graph_builders.as_packed(inputs, n)))
    return v


class _Composite(Wavefunction):
  """Common part of the sum / product wrappers (wavefunctions.py:61-161)."""

  def connect(self, n_sites):
    self._n_sites = int(n_sites)
    for sub in self._sub_wavefunctions:
      sub.connect(n_sites)
    return self

  def native(self, n_sites=None):
    raise NotImplementedError('a composite wavefunction has no single device ansatz; '
                              'use leaves() / amplitudes()')

  def leaves(self):
    out = []
    for sub in self._sub_wavefunctions:
      out += sub.leaves()
    return out

  def get_trainable_variables(self):
    out = []
    for sub in self._sub_wavefunctions:
      out += sub.get_trainable_variables()
    return out

  def _build(self, inputs):
    from . import graph_builders
    n = inputs.shape[1]
    self.connect(n)
    logabs, sign = self.amplitudes(graph_builders.as_packed(inputs, n))
    return sign * torch.exp(logabs)

  def update_norm(self, batch_of_amplitudes, max_value=1e10, log_amplitudes=None):
    return None          # no exp_norm_shift of its own (wavefunctions.py:261-270)

  def __deepcopy__(self, memo):
    raise NotImplementedError('deepcopy of composite wavefunctions is not supported '
                              '(the reference falls back to inspect-based copying, wavefunctions.py:177-204)')

  @classmethod
  def from_hparams(cls, hparams, name=''):
    raise ValueError('Hparams initialization is not supported for %s.' % cls._what)


class SumOfWavefunctions(_Composite):
  """wavefunctions.py:64-97: psi = psi_a + psi_b."""
  _what = 'sum'

  def __init__(self, wf_a, wf_b, name='sum_of_wavefunctions'):
    name = '_plus_'.join([wf_a._unique_name, wf_b._unique_name])
    super().__init__(name=name)
    self._wf_a, self._wf_b = wf_a, wf_b
    self._sub_wavefunctions += [wf_a, wf_b]

  def _parts(self, packed):
    la, sa = self._wf_a.amplitudes(packed)
    lb, sb = self._wf_b.amplitudes(packed)
    m = torch.maximum(la, lb)
    m = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    va, vb = sa * torch.exp(la - m), sb * torch.exp(lb - m)
    return va, vb, m

  def amplitudes(self, packed):
    va, vb, m = self._parts(packed)
    v = va + vb
    return m + torch.log(torch.abs(v)), torch.sign(v)

  def weighted_grad_sum(self, packed, weights):
    # d log(psi_a + psi_b) = (psi_a / psi) d log psi_a + (psi_b / psi) d log psi_b
    weights = weights.reshape(-1, packed.shape[0])
    va, vb, _ = self._parts(packed)
    v = va + vb
    return torch.cat([self._wf_a.weighted_grad_sum(packed, weights * (va / v)),
                      self._wf_b.weighted_grad_sum(packed, weights * (vb / v))], dim=1)


class ProductOfWavefunctions(_Composite):
  """wavefunctions.py:104-159: psi = psi_a * psi_b with psi_b a Wavefunction
  or a float factor."""
  _what = 'product'

  def __init__(self, wf_a, wf_b, name='product_of_wavefunctions'):
    if isinstance(wf_b, Wavefunction):
      name = '_times_'.join([wf_b._unique_name, wf_a._unique_name])
      components = [wf_a, wf_b]
    elif isinstance(wf_b, (float, int)):
      name = '_times_'.join([str(float(wf_b)).replace('-', 'neg_'), wf_a._unique_name])
      components = [wf_a]
    else:
      raise ValueError('Type of other is not supported.')
    super().__init__(name=name)
    self._wf_a, self._wf_b = wf_a, wf_b
    self._sub_wavefunctions += components

  def amplitudes(self, packed):
    la, sa = self._wf_a.amplitudes(packed)
    if isinstance(self._wf_b, Wavefunction):
      lb, sb = self._wf_b.amplitudes(packed)
      return la + lb, sa * sb
    c = float(self._wf_b)
    return la + (math.log(abs(c)) if c != 0.0 else -math.inf), sa * float(np.sign(c))

  def weighted_grad_sum(self, packed, weights):
    out = [self._wf_a.weighted_grad_sum(packed, weights)]
    if isinstance(self._wf_b, Wavefunction):
      out.append(self._wf_b.weighted_grad_sum(packed, weights))
    return torch.cat(out, dim=1)


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


class ResNet1D(_ExpAnsatz):
  """wavefunctions.py:617-711: initial periodic convolution, num_blocks
  residual blocks x + conv(selu(conv(x))) (layers.py:231-296), sum."""
  _kind = 'res_net_1d'

  def __init__(self, num_blocks, num_filters, kernel_size, conv_stride=1, bottleneck=False,
               output_activation='exp', name='res_net_1d'):
    super().__init__(name=name)
    _check_exp(output_activation)
    if bottleneck:
      raise NotImplementedError('BottleneckResBlock1d raises AttributeError in the reference itself '
                                '(undefined self._output_channels, layers.py:348)')
    if conv_stride != 1:
      raise NotImplementedError('residual blocks are built for stride 1 (the skip connection '
                                'requires it, layers.py:216-218)')
    self._num_blocks, self._num_filters, self._kernel_size = num_blocks, num_filters, kernel_size
    self._conv_stride, self._bottleneck, self._output_activation = conv_stride, bottleneck, output_activation
    self._init_args = dict(num_blocks=num_blocks, num_filters=num_filters, kernel_size=kernel_size,
                           conv_stride=conv_stride, bottleneck=bottleneck,
                           output_activation=output_activation)

  def _native_args(self, n):
    return dict(kind=self._kind, n_sites=n, num_layers=self._num_blocks,
                num_filters=self._num_filters, kernel_size=self._kernel_size, nonlinearity='selu')

  def _spatial(self):
    return (self._kernel_size,)

  def _param_shapes(self, n):
    sp, f = self._spatial(), self._num_filters
    shapes = [sp + (1, f), (f,)]
    for _ in range(self._num_blocks):
      shapes += [sp + (f, f), (f,), sp + (f, f), (f,)]
    return shapes

  @classmethod
  def from_hparams(cls, hparams, name=''):
    params = dict(num_blocks=hparams.num_resnet_blocks, num_filters=hparams.num_conv_filters,
                  kernel_size=hparams.kernel_size, conv_stride=hparams.conv_strides,
                  output_activation=layers.NONLINEARITIES[hparams.output_activation])
    if name:
      params['name'] = name
    wf = cls(**params)
    wf._n_sites = hparams.num_sites
    return wf


class ResNet2D(ResNet1D):
  """wavefunctions.py:713-809."""
  _kind = 'res_net_2d'

  def __init__(self, num_blocks, num_filters, kernel_size, conv_stride, size_x, size_y,
               bottleneck=False, output_activation='exp', name='res_net_2d'):
    super().__init__(num_blocks, num_filters, kernel_size, conv_stride, bottleneck,
                     output_activation, name=name)
    self._size_x, self._size_y = size_x, size_y
    self._n_sites = size_x * size_y
    self._init_args.update(size_x=size_x, size_y=size_y)

  def _native_args(self, n):
    return dict(super()._native_args(n), size_x=self._size_x, size_y=self._size_y)

  def _spatial(self):
    return (self._kernel_size, self._kernel_size)

  @classmethod
  def from_hparams(cls, hparams, name=''):
    params = dict(num_blocks=hparams.num_resnet_blocks, num_filters=hparams.num_conv_filters,
                  kernel_size=hparams.kernel_size, conv_stride=hparams.conv_strides,
                  size_x=hparams.size_x, size_y=hparams.size_y,
                  output_activation=layers.NONLINEARITIES[hparams.output_activation])
    if name:
      params['name'] = name
    return cls(**params)


def build_wavefunction(hparams):
  """wavefunctions.py:1157-1196."""
  wavefunction_type = hparams.wavefunction_type
  if wavefunction_type in WAVEFUNCTION_TYPES:
    return WAVEFUNCTION_TYPES[wavefunction_type].from_hparams(hparams)
  if wavefunction_type in ('sum', 'diff', 'prod'):       # wavefunctions.py:1178-1194
    wf_type_a, wf_type_b = hparams.composite_wavefunction_types
    activation_a, activation_b = hparams.composite_output_activations
    wf_a_hparams, wf_b_hparams = copy.copy(hparams), copy.copy(hparams)
    wf_a_hparams.set_hparam('output_activation', activation_a)
    wf_a_hparams.set_hparam('wavefunction_type', wf_type_a)
    wf_b_hparams.set_hparam('output_activation', activation_b)
    wf_b_hparams.set_hparam('wavefunction_type', wf_type_b)
    wf_a = WAVEFUNCTION_TYPES[wf_type_a].from_hparams(wf_a_hparams)
    wf_b = WAVEFUNCTION_TYPES[wf_type_b].from_hparams(wf_b_hparams)
    if wavefunction_type == 'sum':
      return wf_a + wf_b
    if wavefunction_type == 'diff':
      return wf_a - wf_b
    return wf_a * wf_b
  if wavefunction_type in _UNBUILT:
    raise NotImplementedError(
        'wavefunction_type=%r exists in the reference but is outside the CUDA hot path '
        '(SURVEY.md section 2 rows 12-17)' % wavefunction_type)
  raise ValueError('Provided wavefunction_type is not registered.')


WAVEFUNCTION_TYPES = {
    'fully_connected': FullyConnectedNetwork,
    'rbm': RestrictedBoltzmannNetwork,
    'conv_1d': Conv1DNetwork,
    'conv_2d': Conv2DNetwork,
    'res_net_1d': ResNet1D,
    'res_net_2d': ResNet2D,
}

"""Eager stand-in for the Sonnet v1 modules used by cgs-vmc.

TEST INFRASTRUCTURE (see tests/golden/tf_shim/tensorflow/__init__.py).
Implements the documented Sonnet v1 behaviour the reference relies on:
  * AbstractModule: `__call__` -> `_build`; variables live in the module's
    own variable scope (fixed at construction, nested under the scope that was
    current then) and are shared between calls (template semantics);
    `_enter_variable_scope()`, `_unique_name`.
  * Linear: y = x @ w + b, w:[in,out] truncated normal stddev 1/sqrt(in),
    b zeros.  Conv1D / Conv2D: NWC / NHWC cross-correlation, w:[k..., in, out]
    truncated normal stddev 1/sqrt(prod(k) * in), b zeros, VALID padding.
    The convolutions are written as explicit sums over kernel offsets, on
    purpose a different code path from the oracle's torch conv call.
  * Sequential.
"""
import math

import tensorflow as tf

VALID = 'VALID'
SAME = 'SAME'


class AbstractModule:
  def __init__(self, name='module'):
    self._scope_name = tf._unique_scope(name)
    self._unique_name = self._scope_name.split('/')[-1]
    self._ctor_counters = {}

  def _enter_variable_scope(self):
    return tf._frame(self._scope_name, self._ctor_counters)

  def __call__(self, *args, **kwargs):
    # template semantics: sub-module uniquification restarts on every call so
    # that re-connecting the module reuses the same variables.
    with tf._frame(self._scope_name, {}):
      return self._build(*args, **kwargs)

  def _build(self, *args, **kwargs):
    raise NotImplementedError


class Linear(AbstractModule):
  def __init__(self, output_size, name='linear'):
    super().__init__(name=name)
    self._output_size = output_size

  def _build(self, inputs):
    n_in = int(inputs.shape[-1])
    w = tf.get_variable(
        'w', (n_in, self._output_size),
        tf.truncated_normal_initializer(stddev=1.0 / math.sqrt(n_in)))
    b = tf.get_variable('b', (self._output_size,), tf.zeros_initializer())
    return inputs @ w + b


class _ConvND(AbstractModule):
  _rank = None

  def __init__(self, output_channels, kernel_shape, stride=1, padding=SAME,
               name='conv'):
    super().__init__(name=name)
    if padding != VALID:
      raise NotImplementedError('shim implements VALID padding only')
    ks = kernel_shape if isinstance(kernel_shape, (tuple, list)) else \
        (kernel_shape,) * self._rank
    st = stride if isinstance(stride, (tuple, list)) else (stride,) * self._rank
    self._kernel = tuple(int(k) for k in ks)
    self._stride = tuple(int(s) for s in st)
    self._output_channels = output_channels

  def _variables(self, c_in):
    fan_in = c_in
    for k in self._kernel:
      fan_in *= k
    w = tf.get_variable(
        'w', self._kernel + (c_in, self._output_channels),
        tf.truncated_normal_initializer(stddev=1.0 / math.sqrt(fan_in)))
    b = tf.get_variable('b', (self._output_channels,), tf.zeros_initializer())
    return w, b


class Conv1D(_ConvND):
  _rank = 1

  def __init__(self, output_channels, kernel_shape, stride=1, padding=SAME,
               name='conv_1d'):
    super().__init__(output_channels, kernel_shape, stride, padding, name)

  def _build(self, inputs):          # [B, L, C]
    w, b = self._variables(int(inputs.shape[-1]))
    k, = self._kernel
    s, = self._stride
    n_out = (inputs.shape[1] - k) // s + 1
    out = None
    for d in range(k):
      term = inputs[:, d:d + (n_out - 1) * s + 1:s, :] @ w[d]
      out = term if out is None else out + term
    return out + b


class Conv2D(_ConvND):
  _rank = 2

  def __init__(self, output_channels, kernel_shape, stride=1, padding=SAME,
               name='conv_2d'):
    super().__init__(output_channels, kernel_shape, stride, padding, name)

  def _build(self, inputs):          # [B, H, W, C]
    w, b = self._variables(int(inputs.shape[-1]))
    kh, kw = self._kernel
    sh, sw = self._stride
    n_h = (inputs.shape[1] - kh) // sh + 1
    n_w = (inputs.shape[2] - kw) // sw + 1
    out = None
    for dh in range(kh):
      for dw in range(kw):
        patch = inputs[:, dh:dh + (n_h - 1) * sh + 1:sh,
                       dw:dw + (n_w - 1) * sw + 1:sw, :]
        term = patch @ w[dh, dw]
        out = term if out is None else out + term
    return out + b


class Sequential(AbstractModule):
  def __init__(self, layers, name='sequential'):
    super().__init__(name=name)
    self._layers = list(layers)

  def _build(self, inputs):
    net = inputs
    for layer in self._layers:
      net = layer(net)
    return net

"""Eager stand-in for the TensorFlow 1.x op set used by cgs-vmc.

TEST INFRASTRUCTURE, used only by tests/golden/make_golden.py (in the authoring
container, where /root/reference exists) to execute the UNMODIFIED reference
Python modules and record golden input/output vectors.  TensorFlow 1.x has no
Python 3.12 build and cannot be installed here; this module implements the
documented semantics of exactly the ops the reference calls, eagerly, on
torch-CPU float32 tensors (autograd supplies tf.gradients).  It is not a
general TensorFlow replacement and nothing in the product imports it.

Graph-mode caveats of running eagerly:
  * every op executes once, when the reference's graph-construction code runs;
  * tf.metrics.* return (value, value): one `accumulate` == one batch;
  * optimizers do not update variables: `apply_gradients` / `minimize` return
    the gradient list so the generator can record it.
"""
import contextlib
import re
import types

import numpy as np
import torch

Tensor = torch.Tensor
float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
AUTO_REUSE = 'AUTO_REUSE'


class TensorShape(tuple):
  def as_list(self):
    return list(self)


class Variable(torch.Tensor):
  """torch tensor whose `.shape` also answers `.as_list()`."""

  @property
  def shape(self):
    return TensorShape(torch.Tensor.shape.__get__(self))

  @property
  def name(self):
    return getattr(self, '_tf_name', None)


def _get_shape(self):
  return TensorShape(tuple(torch.Tensor.shape.__get__(self)))


torch.Tensor.get_shape = _get_shape   # tf.Tensor.get_shape().as_list()

# ----------------------------------------------------------------------------
# variable store, scopes (Sonnet template semantics: see sonnet shim)
# ----------------------------------------------------------------------------
_VARIABLES = {}          # full name -> tensor
_TRAINABLE = []          # names, creation order
_FRAMES = []             # [{'scope': str, 'counters': dict}]
_GLOBAL_COUNTERS = {}


def _reset_shim_state():
  _VARIABLES.clear()
  del _TRAINABLE[:]
  del _FRAMES[:]
  _GLOBAL_COUNTERS.clear()
  del _UNIFORM_QUEUE[:]


def _current_scope():
  return _FRAMES[-1]['scope'] if _FRAMES else ''


def _join(scope, name):
  return name if not scope else scope + '/' + name


def _unique_scope(name):
  """tf.variable_scope(None, default_name=name) uniquification."""
  counters = _FRAMES[-1]['counters'] if _FRAMES else _GLOBAL_COUNTERS
  key = _join(_current_scope(), name)
  n = counters.get(key, 0)
  counters[key] = n + 1
  return key if n == 0 else '%s_%d' % (key, n)


@contextlib.contextmanager
def _frame(scope, counters):
  _FRAMES.append({'scope': scope, 'counters': counters})
  try:
    yield
  finally:
    _FRAMES.pop()


@contextlib.contextmanager
def variable_scope(name, reuse=None, default_name=None):
  scope = _join(_current_scope(), name) if name else _current_scope()
  with _frame(scope, {}):
    yield


def zeros_initializer(dtype=float32):
  return lambda shape, dtype=dtype: torch.zeros(shape, dtype=dtype)


def constant_initializer(value):
  return lambda shape, dtype=float32: torch.full(tuple(shape), float(value), dtype=dtype)


def truncated_normal_initializer(stddev=1.0, seed=None):
  def init(shape, dtype=float32):
    out = torch.empty(tuple(shape), dtype=torch.float64)
    torch.nn.init.trunc_normal_(out, 0.0, stddev, -2 * stddev, 2 * stddev)
    return out.to(dtype)
  return init


def get_variable(name, shape=None, initializer=None, dtype=float32,
                 trainable=True):
  full = _join(_current_scope(), name)
  if full in _VARIABLES:
    return _VARIABLES[full]
  if initializer is None:
    raise ValueError('shim: get_variable needs an initializer for ' + full)
  if callable(initializer):
    value = initializer(tuple(shape), dtype)
  else:
    value = torch.as_tensor(np.asarray(initializer)).clone()
  value = value.detach().clone().as_subclass(Variable)
  if trainable:
    value.requires_grad_(True)
    _TRAINABLE.append(full)
  value._tf_name = full + ':0'
  _VARIABLES[full] = value
  return value


class GraphKeys:
  TRAINABLE_VARIABLES = 'trainable_variables'


def get_collection(key, scope=None):
  assert key == GraphKeys.TRAINABLE_VARIABLES
  return [_VARIABLES[n] for n in _TRAINABLE
          if scope is None or re.match(scope, n)]


def local_variables():
  return []


def variables_initializer(var_list):
  return None


def global_variables_initializer():
  return None


def local_variables_initializer():
  return None


def group(*args):
  return list(args)


# ----------------------------------------------------------------------------
# random numbers: replayable
# ----------------------------------------------------------------------------
_UNIFORM_QUEUE = []      # tensors returned (in order) by tf.random_uniform
_UNIFORM_LOG = []        # every draw handed out, for recording


def random_uniform(shape, dtype=float32, minval=0, maxval=None, seed=None):
  shape = tuple(int(s) for s in shape)
  if _UNIFORM_QUEUE:
    out = _UNIFORM_QUEUE.pop(0)
    assert tuple(out.shape) == shape, (tuple(out.shape), shape)
  else:
    out = torch.rand(shape, dtype=dtype)
  _UNIFORM_LOG.append(out.clone())
  return out


# ----------------------------------------------------------------------------
# ops
# ----------------------------------------------------------------------------
def _t(x, like=None):
  if isinstance(x, torch.Tensor):
    return x
  t = torch.as_tensor(np.asarray(x))
  if t.dtype == torch.float64:     # numpy constants adopt the graph's float32
    t = t.to(torch.float32)
  return t


def convert_to_tensor(x, dtype=None):
  return _t(x)


def constant(x, dtype=None):
  t = _t(x)
  return t.to(dtype) if dtype is not None else t


def identity(x):
  return x


def multiply(a, b):
  return _t(a) * _t(b)


def add_n(xs):
  out = _t(xs[0])
  for x in xs[1:]:
    out = out + _t(x)
  return out


def stack(xs, axis=0):
  xs = [_t(x).to(torch.int64) if not _t(x).is_floating_point() else _t(x)
        for x in xs]
  return torch.stack(xs, dim=axis)


def argmin(x, axis):
  # TF returns the smallest index on ties; make that explicit.
  mn = x.min(dim=axis, keepdim=True).values
  idx = torch.arange(x.shape[axis]).expand_as(x)
  return torch.where(x == mn, idx, torch.full_like(idx, x.shape[axis])).min(dim=axis).values


def argmax(x, axis):
  mx = x.max(dim=axis, keepdim=True).values
  idx = torch.arange(x.shape[axis]).expand_as(x)
  return torch.where(x == mx, idx, torch.full_like(idx, x.shape[axis])).min(dim=axis).values


def scatter_nd(indices, updates, shape):
  """Zeros of `shape` with updates added at indices (duplicates accumulate)."""
  indices = _t(indices).to(torch.int64)
  updates = _t(updates)
  if updates.dtype == torch.float64:
    updates = updates.to(torch.float32)
  out = torch.zeros(tuple(shape), dtype=updates.dtype)
  return out.index_put(tuple(indices[:, d] for d in range(indices.shape[1])),
                       updates, accumulate=True)


def slice(x, begin, size):   # pylint: disable=redefined-builtin
  idx = tuple(builtins_slice(b, None if s == -1 else b + s)
              for b, s in zip(begin, size))
  return x[idx]


import builtins as _builtins   # noqa: E402
builtins_slice = _builtins.slice


def squeeze(x, axis=None):
  return torch.squeeze(x) if axis is None else torch.squeeze(x, axis)


def expand_dims(x, axis):
  return torch.unsqueeze(x, axis)


def reshape(x, shape):
  return torch.reshape(x, tuple(shape))


def concat(xs, axis):
  return torch.cat(list(xs), dim=axis)


def cast(x, dtype):
  return _t(x).to(dtype)


def greater(a, b):
  return _t(a) > _t(b)


def less(a, b):
  return _t(a) < _t(b)


def abs(x):   # pylint: disable=redefined-builtin
  return torch.abs(x)


sqrt = torch.sqrt
exp = torch.exp
log = torch.log
cosh = torch.cosh
cos = torch.cos
tan = torch.tan
tanh = torch.tanh
sigmoid = torch.sigmoid
square = torch.square


def squared_difference(a, b):
  return (a - b) ** 2


def _axes(axis):
  if axis is None:
    return None
  return tuple(axis) if isinstance(axis, (list, tuple)) else axis


def reduce_sum(x, axis=None):
  return torch.sum(x) if axis is None else torch.sum(x, dim=_axes(axis))


def reduce_mean(x, axis=None):
  return torch.mean(x) if axis is None else torch.mean(x, dim=_axes(axis))


def reduce_max(x, axis=None):
  return torch.max(x) if axis is None else torch.amax(x, dim=_axes(axis))


def stop_gradient(x):
  return x.detach() if isinstance(x, torch.Tensor) else x


def assign_add(ref, value):
  with torch.no_grad():
    ref.add_(_t(value).to(ref.dtype))
  return ref


def assign(ref, value):
  with torch.no_grad():
    ref.copy_(_t(value).to(ref.dtype))
  return ref


def cond(pred, true_fn, false_fn):
  return true_fn() if bool(pred) else false_fn()


def gradients(ys, xs):
  """d sum(ys) / d xs, like tf.gradients."""
  return list(torch.autograd.grad(ys.sum(), list(xs), retain_graph=True,
                                  allow_unused=True))


nn = types.SimpleNamespace(relu=torch.relu, selu=torch.selu)


def _mean_tensor(x):
  return x, x


def _mean(x):
  m = torch.mean(x)
  return m, m


metrics = types.SimpleNamespace(mean_tensor=_mean_tensor, mean=_mean)


class _Optimizer:
  def __init__(self, learning_rate, **kwargs):
    self.learning_rate = learning_rate
    self.kwargs = kwargs

  def apply_gradients(self, grads_and_vars):
    return [g for g, _ in grads_and_vars]

  def minimize(self, loss, var_list=None):
    return gradients(loss, var_list)


def _piecewise_constant(x, boundaries, values):
  x = int(x)
  for b, v in zip(boundaries, values):
    if x <= b:
      return v
  return values[-1]


train = types.SimpleNamespace(
    AdamOptimizer=_Optimizer, GradientDescentOptimizer=_Optimizer,
    RMSPropOptimizer=_Optimizer, MomentumOptimizer=_Optimizer,
    piecewise_constant=_piecewise_constant,
    ExponentialMovingAverage=object, Saver=object)


class Session:
  def run(self, fetches):
    return fetches


class _HParams:
  """tf.contrib.training.HParams: attribute bag with the methods cgs-vmc uses."""

  def __init__(self, hparam_def=None, **kwargs):
    self._names = []
    for k, v in kwargs.items():
      self.add_hparam(k, v)

  def add_hparam(self, name, value):
    self._names.append(name)
    setattr(self, name, value)

  def set_hparam(self, name, value):
    if name not in self._names:
      raise ValueError('Unknown hparam ' + name)
    setattr(self, name, value)

  def override_from_dict(self, values):
    for k, v in values.items():
      self.set_hparam(k, v)
    return self

  def values(self):
    return {k: getattr(self, k) for k in self._names}


from . import contrib   # noqa: E402,F401
contrib.training.HParams = _HParams
gfile = types.SimpleNamespace(GFile=open)
linalg = types.SimpleNamespace(slogdet=torch.linalg.slogdet, det=torch.linalg.det)

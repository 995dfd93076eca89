"""Ansatz forward passes, torch-on-CPU (oracle side; test infrastructure).

Restates wavefunctions.py:328-615 and layers.py:24-160 for the four in-scope
ansaetze.  Everything works on z = log(psi) + exp_norm_shift, i.e. the tensor
that the reference feeds to `add_exp_normalization` followed by `tf.exp`
(wavefunctions.py:206-232, 350-351, 416, 436, 491, 574); `psi()` materialises
the reference's float amplitude `exp(z - shift)`.

Flat parameter layout (shared with the CUDA library, include/cgsvmc.h):
  fully_connected : W_1[in,out], b_1[out], ..., W_L, b_L, W_out[in,1], b_out[1]
  rbm             : a[N], a0[1], (W_l, b_l) hidden relu layers ...,
                    W[in,H], c[H]
  conv_1d         : per layer w[k, cin, cout], b[cout]
  conv_2d         : per layer w[k, k, cin, cout], b[cout]
  res_net_1d / 2d : initial conv w, b (1 -> F), then per block first_conv w, b
                    and second_conv w, b (F -> F)
Matrices are row-major with the Sonnet shapes (snt.Linear w:[in,out];
snt.Conv w:[spatial..., in, out]).
"""
import dataclasses
import math
from typing import List, Tuple

import numpy as np
import torch

NONLINEARITIES = {   # layers.py:13-21
    'relu': torch.relu,
    'exp': torch.exp,
    'cos': torch.cos,
    'tan': torch.tan,
    'tanh': torch.tanh,
    'sigmoid': torch.sigmoid,
    'identity': lambda x: x,
    'selu': torch.selu,      # fixed inside the ResNet blocks (layers.py:225, 293)
}


@dataclasses.dataclass
class AnsatzSpec:
  kind: str                 # 'fully_connected' | 'rbm' | 'conv_1d' | 'conv_2d' | 'res_net_1d' | 'res_net_2d'
  n_sites: int
  num_layers: int = 3       # num_fc_layers / num_conv_layers / num_resnet_blocks
  layer_size: int = 80      # fc_layer_size
  num_filters: int = 16
  kernel_size: int = 5
  size_x: int = 1
  size_y: int = 1
  nonlinearity: str = 'relu'

  def __post_init__(self):
    if self.kind in ('conv_2d', 'res_net_2d') and self.size_x * self.size_y != self.n_sites:
      raise ValueError('size_x * size_y must equal n_sites for %s' % self.kind)


def param_shapes(spec: AnsatzSpec) -> List[Tuple[str, Tuple[int, ...]]]:
  """Ordered (name, shape) list defining the flat parameter layout."""
  out = []
  if spec.kind == 'fully_connected':       # wavefunctions.py:345-349
    n_in = spec.n_sites
    for l in range(spec.num_layers):
      out += [('w%d' % l, (n_in, spec.layer_size)), ('b%d' % l, (spec.layer_size,))]
      n_in = spec.layer_size
    out += [('w_out', (n_in, 1)), ('b_out', (1,))]
  elif spec.kind == 'rbm':                 # wavefunctions.py:410-417
    out += [('a', (spec.n_sites, 1)), ('a0', (1,))]
    n_in = spec.n_sites
    for l in range(spec.num_layers):
      out += [('w%d' % l, (n_in, spec.layer_size)), ('b%d' % l, (spec.layer_size,))]
      n_in = spec.layer_size
    out += [('w_rbm', (n_in, spec.layer_size)), ('c', (spec.layer_size,))]
  elif spec.kind in ('conv_1d', 'conv_2d'):  # wavefunctions.py:483-489, 566-572
    c_in = 1
    k = spec.kernel_size
    for l in range(spec.num_layers):
      shape = (k, c_in, spec.num_filters) if spec.kind == 'conv_1d' else \
          (k, k, c_in, spec.num_filters)
      out += [('w%d' % l, shape), ('b%d' % l, (spec.num_filters,))]
      c_in = spec.num_filters
  elif spec.kind in ('res_net_1d', 'res_net_2d'):   # wavefunctions.py:651-671, 753-769
    # initial periodic conv (1 -> F), then num_layers ResBlocks of two F -> F convs
    k, f = spec.kernel_size, spec.num_filters
    sp = (k,) if spec.kind == 'res_net_1d' else (k, k)
    out += [('w_init', sp + (1, f)), ('b_init', (f,))]
    for l in range(spec.num_layers):
      out += [('w%d_1' % l, sp + (f, f)), ('b%d_1' % l, (f,)),
              ('w%d_2' % l, sp + (f, f)), ('b%d_2' % l, (f,))]
  else:
    raise ValueError('Provided wavefunction_type is not registered.')
  return out


def num_params(spec: AnsatzSpec) -> int:
  return sum(int(np.prod(s)) for _, s in param_shapes(spec))


def init_params(spec: AnsatzSpec, seed: int = 1234, bias_scale: float = 0.0,
                dtype=torch.float32) -> List[torch.Tensor]:
  """Sonnet v1 default initialisers: weights truncated-normal (+-2 sigma) with
  sigma = 1/sqrt(fan_in), biases zero.  `bias_scale` > 0 draws non-zero biases
  so that tests exercise them."""
  gen = torch.Generator().manual_seed(seed)
  params = []
  for name, shape in param_shapes(spec):
    if len(shape) == 1:
      p = bias_scale * torch.randn(shape, generator=gen, dtype=torch.float64)
    else:
      fan_in = int(np.prod(shape[:-1]))
      p = torch.empty(shape, dtype=torch.float64)
      torch.nn.init.trunc_normal_(p, mean=0.0, std=1.0 / math.sqrt(fan_in),
                                  a=-2.0 / math.sqrt(fan_in),
                                  b=2.0 / math.sqrt(fan_in), generator=gen)
    params.append(p.to(dtype))
  return params


def flatten(params: List[torch.Tensor]) -> torch.Tensor:
  return torch.cat([p.reshape(-1) for p in params])


def unflatten(spec: AnsatzSpec, flat: torch.Tensor) -> List[torch.Tensor]:
  out, off = [], 0
  for _, shape in param_shapes(spec):
    n = int(np.prod(shape))
    out.append(flat[off:off + n].reshape(shape))
    off += n
  return out


def log_cosh(x: torch.Tensor, literal: bool = False) -> torch.Tensor:
  """log(cosh(x)).  The reference applies tf.cosh then tf.log literally
  (wavefunctions.py:415) which overflows float32 for |x| > ~89; the default
  here is the algebraically identical |x| + log1p(exp(-2|x|)) - ln 2."""
  if literal:
    return torch.log(torch.cosh(x))
  ax = torch.abs(x)
  return ax + torch.log1p(torch.exp(-2.0 * ax)) - math.log(2.0)


def _pad_periodic_1d(x, k):
  """layers.py:51-74 on [B, L, C]: odd k pads (k-1)/2 both sides; even k pads
  k/2 on the left and k/2-1 on the right."""
  size = x.shape[1]
  if k % 2 == 1:
    left = right = (k - 1) // 2
  else:
    left, right = k // 2, k // 2 - 1
  return torch.cat([x[:, size - left:], x, x[:, :right]], dim=1)


def _pad_periodic_2d(x, k):
  """layers.py:117-148 on [B, X, Y, C] (NHWC): odd k pads (k-1)/2 on all four
  sides; even k pads k/2-1 before and k/2 after on both axes."""
  if k % 2 == 1:
    before = after = (k - 1) // 2
  else:
    before, after = k // 2 - 1, k // 2
  size2 = x.shape[2]
  size1 = x.shape[1]
  x = torch.cat([x[:, :, size2 - before:], x, x[:, :, :after]], dim=2)
  x = torch.cat([x[:, size1 - before:], x, x[:, :after]], dim=1)
  return x


def _conv_valid(x, w, b):
  """snt.Conv1D / snt.Conv2D with padding VALID, stride 1: NHWC
  cross-correlation, w:[spatial..., in, out] (layers.py:46-49, 113-115)."""
  if w.dim() == 3:
    y = torch.nn.functional.conv1d(x.permute(0, 2, 1), w.permute(2, 1, 0), b)
    return y.permute(0, 2, 1)
  y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b)
  return y.permute(0, 2, 3, 1)


def log_amp(spec: AnsatzSpec, params: List[torch.Tensor],
            configs: torch.Tensor, literal_log_cosh: bool = False) -> torch.Tensor:
  """z(sigma) = log psi + shift for a [B, N] batch of +-1 configurations."""
  act = NONLINEARITIES[spec.nonlinearity]
  x = configs.to(params[0].dtype)
  if spec.kind == 'fully_connected':        # wavefunctions.py:345-353, 370-371
    h = x
    for l in range(spec.num_layers):
      h = act(h @ params[2 * l] + params[2 * l + 1])
    return (h @ params[-2] + params[-1]).reshape(-1)
  if spec.kind == 'rbm':                    # wavefunctions.py:410-417, 434-436
    onsite = (x @ params[0] + params[1]).reshape(-1)
    h = x
    for l in range(spec.num_layers):
      h = act(h @ params[2 + 2 * l] + params[3 + 2 * l])
    theta = h @ params[-2] + params[-1]
    return onsite + log_cosh(theta, literal_log_cosh).sum(dim=1)
  if spec.kind == 'conv_1d':                # wavefunctions.py:483-491, 510
    h = x.unsqueeze(2)
    for l in range(spec.num_layers):
      h = _conv_valid(_pad_periodic_1d(h, spec.kernel_size),
                      params[2 * l], params[2 * l + 1])
      if l + 1 != spec.num_layers:
        h = act(h)
    return h.sum(dim=(1, 2))
  if spec.kind == 'conv_2d':                # wavefunctions.py:566-574, 593-595
    h = x.reshape(-1, spec.size_x, spec.size_y, 1)
    for l in range(spec.num_layers):
      h = _conv_valid(_pad_periodic_2d(h, spec.kernel_size),
                      params[2 * l], params[2 * l + 1])
      if l + 1 != spec.num_layers:
        h = act(h)
    return h.sum(dim=(1, 2, 3))
  if spec.kind in ('res_net_1d', 'res_net_2d'):
    # initial conv without nonlinearity, then x <- x + conv2(selu(conv1(x)))
    # per block (layers.py:203-228, 271-296), then the sum over sites and channels
    one_d = spec.kind == 'res_net_1d'
    pad = _pad_periodic_1d if one_d else _pad_periodic_2d
    h = x.unsqueeze(2) if one_d else x.reshape(-1, spec.size_x, spec.size_y, 1)
    k = spec.kernel_size
    h = _conv_valid(pad(h, k), params[0], params[1])
    for l in range(spec.num_layers):
      w1, b1, w2, b2 = params[2 + 4 * l:6 + 4 * l]
      r = torch.selu(_conv_valid(pad(h, k), w1, b1))
      h = h + _conv_valid(pad(r, k), w2, b2)
    return h.sum(dim=(1, 2) if one_d else (1, 2, 3))
  raise ValueError('Provided wavefunction_type is not registered.')


def psi(spec, params, configs, shift=-10.0, **kw):
  """The reference's float amplitude exp(z - exp_norm_shift)
  (wavefunctions.py:232 with the initial shift of -10)."""
  return torch.exp(log_amp(spec, params, configs, **kw) - shift)


def update_norm_shift(shift, psi_batch, max_value=1e10):
  """Wavefunction.update_norm (wavefunctions.py:261-288): bump the shift by
  log(max psi) - log(max_value) iff max psi exceeds max_value."""
  log_max = float(torch.log(psi_batch.max()))
  max_log = math.log(max_value)
  return shift + (log_max - max_log) if log_max > max_log else shift

"""Signed output activations and composite wavefunctions (oracle side; test
infrastructure).

Restates FullyConnectedNetwork / Conv1DNetwork / Conv2DNetwork with an output
activation other than exp (wavefunctions.py:350-353, 490-493, 573-576: no
exp_norm_shift, psi = f(z)), the sum / difference / product wrappers
(wavefunctions.py:61-165) and build_wavefunction's composite branch
(wavefunctions.py:1178-1194) in float64 torch, in amplitude form like the
reference; gradients by autograd (training.py:545-548).
"""
import torch

from . import ansatz as _ansatz
from . import hamiltonian as _hamiltonian

OUTPUT_ACTIVATIONS = {       # layers.py:13-21
    'relu': torch.relu, 'cos': torch.cos, 'tan': torch.tan, 'tanh': torch.tanh,
    'sigmoid': torch.sigmoid, 'identity': lambda z: z,
}


class Leaf:
  """One parameterised ansatz: spec, parameter list, output activation name
  and (for exp) the exp_norm_shift."""

  def __init__(self, spec, params, activation='exp', shift=-10.0):
    self.spec, self.params, self.activation, self.shift = spec, params, activation, shift

  def psi(self, configs, params=None):
    z = _ansatz.log_amp(self.spec, self.params if params is None else params, configs)
    if self.activation == 'exp':
      return torch.exp(z - self.shift)
    return OUTPUT_ACTIVATIONS[self.activation](z)


def psi(kind, leaves, configs, params=None):
  """kind: 'single' | 'sum' | 'diff' | 'prod' (diff = a + (-1.) * b,
  wavefunctions.py:163-165)."""
  params = [None] * len(leaves) if params is None else params
  a = leaves[0].psi(configs, params[0])
  if kind == 'single':
    return a
  b = leaves[1].psi(configs, params[1])
  if kind == 'sum':
    return a + b
  if kind == 'diff':
    return a - b
  if kind == 'prod':
    return a * b
  raise ValueError(kind)


def local_energy(kind, leaves, configs, bonds_ij, jx, jz):
  """local_value (operators.py:249-259): (diag psi + offdiag) / psi."""
  fn = lambda c: psi(kind, leaves, c)
  diag, off = _hamiltonian.build(configs, bonds_ij, jx, jz, fn)
  return diag + off / fn(configs)


def energy_gradient(kind, leaves, configs, bonds_ij, jx, jz):
  """One accumulate + apply of EnergyGradientOptimizer (training.py:539-564):
  returns (flat gradient over all leaves, mean energy)."""
  e = local_energy(kind, leaves, configs, bonds_ij, jx, jz).detach()
  leaf_params = [[p.detach().clone().requires_grad_(True) for p in leaf.params] for leaf in leaves]
  value = psi(kind, leaves, configs, leaf_params)
  scaled = value / value.detach()
  flat = [p for ps in leaf_params for p in ps]
  g1 = torch.autograd.grad(scaled.sum(), flat, retain_graph=True, allow_unused=True)
  g2 = torch.autograd.grad((e * scaled).sum(), flat, allow_unused=True)
  cat = lambda gs: torch.cat([torch.zeros_like(p).reshape(-1) if g is None else g.reshape(-1)
                              for g, p in zip(gs, flat)])
  return cat(g2) - e.mean() * cat(g1), float(e.mean())

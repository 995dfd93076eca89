"""CPU-side checks of the host logic behind signed / composite wavefunctions
(cgs_vmc_b200.wavefunctions): the (log|psi|, sign) algebra of the sum / product
wrappers and the chain-rule weights of their gradients, on stand-in leaves
that need no GPU; the product-side lattice helpers against the oracle's."""
import math

import numpy as np
import pytest
import torch

from cgs_vmc_b200 import lattices, wavefunctions
from oracle import lattices as olattices


class FakeLeaf(wavefunctions.Wavefunction):
  """psi_b = s_b * exp(l_b) with d log psi_b / d theta = O[b, :] (fixed tables)."""

  def __init__(self, logabs, sign, o_matrix, name='fake'):
    super().__init__(name=name)
    self._l, self._s, self._o = logabs, sign, o_matrix

  def connect(self, n_sites):
    return self

  def leaves(self):
    return [self]

  def amplitudes(self, packed):
    return self._l, self._s

  def weighted_grad_sum(self, packed, weights):
    return weights.reshape(-1, self._l.shape[0]).double() @ self._o


def _leaf(seed, batch=7, n_params=3):
  g = torch.Generator().manual_seed(seed)
  logabs = torch.randn(batch, generator=g, dtype=torch.float64)
  sign = torch.where(torch.rand(batch, generator=g) < 0.4, -1.0, 1.0).double()
  o = torch.randn(batch, n_params, generator=g, dtype=torch.float64)
  return FakeLeaf(logabs, sign, o, name='fake%d' % seed)


PACKED = torch.zeros(7, 1, dtype=torch.int64)     # only its batch dimension is looked at


def _psi(leaf):
  return leaf._s * torch.exp(leaf._l)


@pytest.mark.parametrize('name', ['identity', 'tanh', 'sigmoid', 'relu', 'cos', 'tan'])
def test_output_activation_value_and_slope(name):
  z = torch.linspace(-1.3, 1.4, 23, dtype=torch.float64).requires_grad_(True)
  v, dv = wavefunctions._output_value_and_slope(name, z)
  (g,) = torch.autograd.grad(v.sum(), z)
  np.testing.assert_allclose(dv.detach().numpy(), g.numpy(), rtol=1e-12, atol=1e-12)
  ref = {'identity': lambda t: t, 'tanh': torch.tanh, 'sigmoid': torch.sigmoid, 'relu': torch.relu,
         'cos': torch.cos, 'tan': torch.tan}[name](z)
  np.testing.assert_allclose(v.detach().numpy(), ref.detach().numpy(), rtol=1e-12)


def test_sum_of_wavefunctions_algebra():
  a, b = _leaf(1), _leaf(2)
  s = a + b
  logabs, sign = s.amplitudes(None)
  psi = _psi(a) + _psi(b)
  np.testing.assert_allclose((sign * torch.exp(logabs)).numpy(), psi.numpy(), rtol=1e-12)
  # d log(psi_a + psi_b) = (psi_a O_a + psi_b O_b) / psi, leaf by leaf
  w = torch.randn(2, 7, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
  got = s.weighted_grad_sum(PACKED, w)
  ref = torch.cat([(w * (_psi(a) / psi)) @ a._o, (w * (_psi(b) / psi)) @ b._o], dim=1)
  np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-12)
  assert [l._unique_name for l in s.leaves()] == ['fake1', 'fake2']


def test_difference_and_scalar_product():
  a, b = _leaf(4), _leaf(5)
  d = a - b                                      # a + (b * -1.), wavefunctions.py:163-165
  logabs, sign = d.amplitudes(None)
  psi = _psi(a) - _psi(b)
  np.testing.assert_allclose((sign * torch.exp(logabs)).numpy(), psi.numpy(), rtol=1e-12)
  w = torch.ones(1, 7, dtype=torch.float64)
  got = d.weighted_grad_sum(PACKED, w)
  ref = torch.cat([(w * (_psi(a) / psi)) @ a._o, (w * (-_psi(b) / psi)) @ b._o], dim=1)
  np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-12)
  scaled = a * 2.5
  l2, s2 = scaled.amplitudes(None)
  np.testing.assert_allclose(l2.numpy(), (a._l + math.log(2.5)).numpy(), rtol=1e-12)
  assert torch.equal(s2, a._s) and len(scaled.leaves()) == 1


def test_product_of_wavefunctions_algebra():
  a, b = _leaf(6), _leaf(7)
  p = a * b
  logabs, sign = p.amplitudes(None)
  np.testing.assert_allclose((sign * torch.exp(logabs)).numpy(), (_psi(a) * _psi(b)).numpy(), rtol=1e-12)
  w = torch.randn(1, 7, generator=torch.Generator().manual_seed(8), dtype=torch.float64)
  got = p.weighted_grad_sum(PACKED, w)               # d log(psi_a psi_b) = O_a + O_b
  np.testing.assert_allclose(got.numpy(), torch.cat([w @ a._o, w @ b._o], dim=1).numpy(), rtol=1e-12)
  with pytest.raises(ValueError, match='not supported'):
    wavefunctions.ProductOfWavefunctions(a, 'x')
  assert p.update_norm(None) is None
  with pytest.raises(NotImplementedError):
    p.native()


def test_sum_survives_cancellation_and_zero():
  """A leaf that is exactly zero (log = -inf, sign 0) drops out; exact
  cancellation gives sign 0 rather than NaN signs."""
  a = _leaf(9)
  zero = FakeLeaf(torch.full((7,), -math.inf, dtype=torch.float64), torch.zeros(7, dtype=torch.float64),
                  torch.zeros(7, 3, dtype=torch.float64), name='zero')
  logabs, sign = (a + zero).amplitudes(None)
  np.testing.assert_allclose(logabs.numpy(), a._l.numpy(), rtol=1e-12)
  assert torch.equal(sign, a._s)
  _, sign0 = (a - a).amplitudes(None)
  assert torch.all(sign0 == 0)


def test_product_side_lattices_match_the_oracle():
  assert lattices.chain_bonds(20) == olattices.chain_bonds(20)
  for size in (4, 6, 10):
    assert lattices.square_nn_bonds(size) == olattices.square_nn_bonds(size)
    assert lattices.square_nnn_bonds(size) == olattices.square_nnn_bonds(size)
    for got, ref in zip(lattices.j1j2_couplings(size, 0.5), olattices.j1j2_couplings(size, 0.5)):
      assert np.array_equal(got, ref)
  assert lattices.square_nn_bonds(4, 6) == olattices.square_nn_bonds(4, 6)
  for got, ref in zip(lattices.heisenberg_couplings(lattices.chain_bonds(8)),
                      olattices.heisenberg_couplings(olattices.chain_bonds(8))):
    assert np.array_equal(got, ref)

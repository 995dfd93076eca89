"""Mirror of the reference's operators.py on the CUDA library
(operators.py:13-287): Operator interface, HeisenbergBond and
HeisenbergHamiltonian.  j_x / j_z may be scalars (the reference) or one value
per bond (J1-J2; the reference expresses that as a sum of HeisenbergBond)."""
import numpy as np
import torch

from . import _native


class Operator:
  """Operator base class (operators.py:13-87)."""

  def build(self, wavefunction, inputs, psi=None):
    raise NotImplementedError

  def local_value(self, wavefunction, inputs, psi=None):
    raise NotImplementedError

  def apply_in_place(self, wavefunction, inputs, psi=None):
    raise NotImplementedError

  def apply(self, wavefunction):
    raise NotImplementedError


class _BondListOperator(Operator):
  def __init__(self, bonds, j_x, j_z):
    self._bonds_list = [(int(a), int(b)) for a, b in bonds]
    n = len(self._bonds_list)
    self._j_x = j_x
    self._j_z = j_z
    self._jx_array = np.ascontiguousarray(np.broadcast_to(np.asarray(j_x, dtype=np.float32), (n,)))
    self._jz_array = np.ascontiguousarray(np.broadcast_to(np.asarray(j_z, dtype=np.float32), (n,)))
    self._native = {}

  def native(self, n_sites):
    if n_sites not in self._native:
      self._native[n_sites] = _native.Hamiltonian(
          np.asarray(self._bonds_list, dtype=np.int32).reshape(-1, 2), self._jx_array,
          self._jz_array, n_sites)
    return self._native[n_sites]

  def _evaluate(self, wavefunction, inputs):
    from . import graph_builders
    n = inputs.shape[1]
    a = wavefunction.native(n)
    packed = graph_builders.as_packed(inputs, n)
    e, z, diag, off = a.local_energy(self.native(n), packed, want_parts=True)
    return e, z - wavefunction._exp_norm_shift, diag, off

  def build(self, wavefunction, inputs, psi=None):
    """(diagonal matrix element, off-diagonal term) of <R|O|psi>
    (operators.py:137-169, 227-247)."""
    _, logpsi, diag, off = self._evaluate(wavefunction, inputs)
    return diag, off * torch.exp(logpsi)

  def local_value(self, wavefunction, inputs, psi=None):
    """<R|O|psi> / <R|psi> (operators.py:171-181, 249-259)."""
    e, _, _, _ = self._evaluate(wavefunction, inputs)
    return e

  def apply_in_place(self, wavefunction, inputs, psi=None):
    """<R|O|psi> (operators.py:183-193, 261-271)."""
    e, logpsi, _, _ = self._evaluate(wavefunction, inputs)
    return e * torch.exp(logpsi)

  def apply(self, wavefunction):
    raise NotImplementedError('Operator.apply builds a TransformedWavefunction (operators.py:90-125); '
                              'no driver uses it and it is outside the CUDA hot path')


class HeisenbergBond(_BondListOperator):
  """S_i . S_j on one bond (operators.py:128-210)."""

  def __init__(self, bond, j_x, j_z):
    super().__init__([bond], j_x, j_z)
    self._bond = tuple(bond)


class HeisenbergHamiltonian(_BondListOperator):
  """Heisenberg Hamiltonian on a bond list (operators.py:212-287)."""

  def __init__(self, bonds, j_x, j_z):
    super().__init__(bonds, j_x, j_z)
    self._heisenberg_bonds = self._bonds_list

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
    """(E_loc, log|psi|, sign psi, diag, offdiag / psi) on the device."""
    from . import graph_builders
    n = inputs.shape[1]
    packed = graph_builders.as_packed(inputs, n)
    ham = self.native(n)
    if wavefunction.fast_path:          # fused local-energy kernels
      a = wavefunction.native(n)
      e, z, diag, off = a.local_energy(ham, packed, want_parts=True)
      return e, z - wavefunction._exp_norm_shift, torch.ones_like(z), diag, off
    # signed / composite amplitudes: psi on the bond-flipped configurations of
    # cgsvmc_flip_enum through the parts' kernels, combined on the device
    wavefunction.connect(n)
    b = packed.shape[0]
    logabs, sign = wavefunction.amplitudes(packed)
    _, flipped = ham.flip_enum(packed, want_flipped=True)
    fl, fs = wavefunction.amplitudes(flipped.reshape(b * ham.n_bonds, -1))
    logabs, sign = logabs.float().contiguous(), sign.float().contiguous()
    e, diag, off = _native.local_energy_from_amps(
        ham, packed, logabs, sign, fl.float().reshape(b, ham.n_bonds).contiguous(),
        fs.float().reshape(b, ham.n_bonds).contiguous(), want_parts=True)
    return e, logabs, sign, diag, off

  def build(self, wavefunction, inputs, psi=None):
    """(diagonal matrix element, off-diagonal term) of <R|O|psi>
    (operators.py:137-169, 227-247)."""
    _, logpsi, sign, diag, off = self._evaluate(wavefunction, inputs)
    return diag, off * sign * torch.exp(logpsi)

  def local_value(self, wavefunction, inputs, psi=None):
    """<R|O|psi> / <R|psi> (operators.py:171-181, 249-259)."""
    return self._evaluate(wavefunction, inputs)[0]

  def apply_in_place(self, wavefunction, inputs, psi=None):
    """<R|O|psi> (operators.py:183-193, 261-271)."""
    e, logpsi, sign, _, _ = self._evaluate(wavefunction, inputs)
    return e * sign * torch.exp(logpsi)

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

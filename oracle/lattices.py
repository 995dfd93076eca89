"""Bond lists for the synthetic lattices (oracle side; test infrastructure).

The reference only ever sees a list of ``(i, j)`` pairs: either the 1-D
periodic default ``[(i, (i + 1) % n_sites)]`` (run_training.py:109) or the two
integer columns of ``J.txt`` (run_training.py:105-107).  Site numbering of the
square lattices is ``s = x * size_y + y`` so that it matches the
``tf.reshape(inputs, [-1, size_x, size_y, 1])`` of wavefunctions.py:593.
"""
import numpy as np


def chain_bonds(n_sites):
  """1-D periodic chain, run_training.py:109."""
  return [(i, (i + 1) % n_sites) for i in range(n_sites)]


def square_nn_bonds(size_x, size_y=None):
  """Nearest-neighbour bonds of a periodic size_x x size_y square lattice."""
  size_y = size_x if size_y is None else size_y
  bonds = []
  for x in range(size_x):
    for y in range(size_y):
      s = x * size_y + y
      bonds.append((s, ((x + 1) % size_x) * size_y + y))
      bonds.append((s, x * size_y + (y + 1) % size_y))
  return bonds


def square_nnn_bonds(size_x, size_y=None):
  """Next-nearest-neighbour (diagonal) bonds of the periodic square lattice."""
  size_y = size_x if size_y is None else size_y
  bonds = []
  for x in range(size_x):
    for y in range(size_y):
      s = x * size_y + y
      bonds.append((s, ((x + 1) % size_x) * size_y + (y + 1) % size_y))
      bonds.append((s, ((x + 1) % size_x) * size_y + (y - 1) % size_y))
  return bonds


def heisenberg_couplings(bonds, j_x=-1.0, j_z=1.0):
  """Uniform couplings as in HeisenbergHamiltonian(bonds, j_x, j_z)
  (operators.py:215-225).  Returns (ij[int32 n,2], jx[f32 n], jz[f32 n])."""
  ij = np.asarray(bonds, dtype=np.int32).reshape(-1, 2)
  n = ij.shape[0]
  return (ij, np.full(n, j_x, dtype=np.float32),
          np.full(n, j_z, dtype=np.float32))


def j1j2_couplings(size, j2=0.5, marshall=True):
  """J1-J2 model on the periodic size x size lattice as a per-bond list.

  Each bond is a HeisenbergBond(bond, j_x, j_z) (operators.py:128-135); the
  NN bonds carry (j_x, j_z) = (-1, 1) when Marshall-rotated, the NNN bonds
  (+j2, j2) (they connect the same sublattice so the rotation leaves them).
  """
  nn = square_nn_bonds(size)
  nnn = square_nnn_bonds(size)
  ij = np.asarray(nn + nnn, dtype=np.int32)
  jx = np.concatenate([np.full(len(nn), -1.0 if marshall else 1.0),
                       np.full(len(nnn), j2)]).astype(np.float32)
  jz = np.concatenate([np.full(len(nn), 1.0),
                       np.full(len(nnn), j2)]).astype(np.float32)
  return ij, jx, jz

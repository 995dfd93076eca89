"""Bond lists of the lattices the drivers and benchmarks use:
periodic chain (run_training.py:109), periodic L x L square lattice with site
index s = x * size_y + y (the reshape of wavefunctions.py:593), nearest and
next-nearest neighbours, and the coupling arrays of the Heisenberg / J1-J2
models in the reference's matrix-element convention (operators.py:165-169)."""
import numpy as np


def chain_bonds(n_sites):
  return [(i, (i + 1) % n_sites) for i in range(n_sites)]


def square_nn_bonds(size_x, size_y=None):
  size_y = size_x if size_y is None else size_y
  bonds = []
  for x in range(size_x):
    for y in range(size_y):
      s = x * size_y + y
      bonds.append((s, ((x + 1) % size_x) * size_y + y))
      bonds.append((s, x * size_y + (y + 1) % size_y))
  return bonds


def square_nnn_bonds(size_x, size_y=None):
  size_y = size_x if size_y is None else size_y
  bonds = []
  for x in range(size_x):
    for y in range(size_y):
      s = x * size_y + y
      xp = (x + 1) % size_x
      bonds.append((s, xp * size_y + (y + 1) % size_y))
      bonds.append((s, xp * size_y + (y - 1) % size_y))
  return bonds


def heisenberg_couplings(bonds, j_x=-1.0, j_z=1.0):
  """(ij int32 [n, 2], jx float32 [n], jz float32 [n]) with uniform couplings."""
  ij = np.asarray(bonds, dtype=np.int32).reshape(-1, 2)
  n = ij.shape[0]
  return ij, np.full(n, j_x, dtype=np.float32), np.full(n, j_z, dtype=np.float32)


def j1j2_couplings(size, j2=0.5):
  """NN bonds with (jx, jz) = (-1, 1) (Marshall-rotated) and NNN bonds with
  (+j2, j2): the rotation leaves same-sublattice bonds unchanged."""
  nn, nnn = square_nn_bonds(size), square_nnn_bonds(size)
  ij = np.asarray(nn + nnn, dtype=np.int32).reshape(-1, 2)
  jx = np.concatenate([np.full(len(nn), -1.0), np.full(len(nnn), j2)]).astype(np.float32)
  jz = np.concatenate([np.full(len(nn), 1.0), np.full(len(nnn), j2)]).astype(np.float32)
  return ij, jx, jz

"""Small exact diagonalisation in the Sz = 0 sector (oracle side; test
infrastructure).  Matrix elements follow operators.py:169:
  <s|H|s>   = sum_b jz_b / 4 * s_i s_j
  <s'|H|s>  = jx_b / 2 for s' = s with the antiparallel pair (i, j) exchanged.
Provides the known answers of SURVEY.md appendix C and a FullVector-style
lookup amplitude (wavefunctions.py:1001-1080 uses Lin tables; a dict does the
same job here) for the zero-variance test E_loc(s) == E0.
"""
import itertools

import numpy as np
import scipy.sparse
import scipy.sparse.linalg


def sz0_basis(n_sites):
  """All bit patterns with n_sites // 2 bits cleared (= spins down), sorted."""
  n_up = n_sites - n_sites // 2
  states = []
  for ups in itertools.combinations(range(n_sites), n_up):
    s = 0
    for i in ups:
      s |= 1 << i
    states.append(s)
  return np.array(sorted(states), dtype=np.int64)


def hamiltonian_matrix(n_sites, bonds_ij, jx, jz):
  basis = sz0_basis(n_sites)
  index = {int(s): k for k, s in enumerate(basis)}
  rows, cols, vals = [], [], []
  for k, s in enumerate(basis):
    s = int(s)
    diag = 0.0
    for b, (i, j) in enumerate(np.asarray(bonds_ij).reshape(-1, 2)):
      bi, bj = (s >> int(i)) & 1, (s >> int(j)) & 1
      diag += 0.25 * float(jz[b]) * (1.0 if bi == bj else -1.0)
      if bi != bj:
        t = s ^ ((1 << int(i)) | (1 << int(j)))
        rows.append(index[t]); cols.append(k); vals.append(0.5 * float(jx[b]))
    rows.append(k); cols.append(k); vals.append(diag)
  dim = len(basis)
  h = scipy.sparse.coo_matrix((vals, (rows, cols)), shape=(dim, dim)).tocsr()
  return basis, h


def ground_state(n_sites, bonds_ij, jx, jz):
  """Returns (E0, basis[int64 dim], vector[float64 dim]) of the lowest state."""
  basis, h = hamiltonian_matrix(n_sites, bonds_ij, jx, jz)
  if len(basis) <= 2000:
    w, v = np.linalg.eigh(h.toarray())
    return float(w[0]), basis, v[:, 0]
  w, v = scipy.sparse.linalg.eigsh(h, k=1, which='SA', tol=1e-12)
  return float(w[0]), basis, v[:, 0]


def lookup_amplitude(basis, vector):
  """Returns psi_fn(configs[B, N] of +-1) -> amplitudes, a stand-in for the
  reference's FullVector ansatz."""
  import torch
  index = {int(s): k for k, s in enumerate(basis)}

  def psi_fn(configs):
    c = np.asarray(configs)
    n = c.shape[1]
    weights = (1 << np.arange(n, dtype=np.int64))
    keys = ((c > 0).astype(np.int64) * weights).sum(axis=1)
    out = np.array([vector[index[int(k)]] if int(k) in index else 0.0
                    for k in keys])
    return torch.from_numpy(out)
  return psi_fn

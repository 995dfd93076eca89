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


def configs_of(basis, n_sites):
  """float64 [dim, N] of +-1 for the bit patterns of `basis` (bit i = site i up)."""
  bits = (np.asarray(basis)[:, None] >> np.arange(n_sites)[None, :]) & 1
  return 2.0 * bits.astype(np.float64) - 1.0


def exact_expectation(n_sites, bonds_ij, jx, jz, psi_fn):
  """<psi|H|psi> / <psi|psi> over the Sz = 0 sector: what the mean local value
  of evaluation.py:98-102 converges to when the walkers sample |psi|^2.
  psi_fn maps float64 configurations [dim, N] to amplitudes [dim]."""
  basis, h = hamiltonian_matrix(n_sites, bonds_ij, jx, jz)
  psi = np.asarray(psi_fn(configs_of(basis, n_sites)), dtype=np.float64)
  return float(psi @ (h @ psi) / (psi @ psi))


def exact_acceptance_rate(n_sites, psi_fn):
  """Stationary acceptance probability of the exchange sampler of
  graph_builders.py:54-89: a uniformly random up site and a uniformly random
  down site are exchanged and the move is accepted with min(1, (psi'/psi)^2):
    A = sum_s pi(s) / (n_up n_dn) sum_{u, d} min(1, pi(s^{ud}) / pi(s))."""
  basis = sz0_basis(n_sites)
  index = {int(s): k for k, s in enumerate(basis)}
  psi = np.asarray(psi_fn(configs_of(basis, n_sites)), dtype=np.float64)
  pi = psi * psi / np.sum(psi * psi)
  n_up = n_sites - n_sites // 2
  n_dn = n_sites // 2
  total = 0.0
  for k, s in enumerate(basis):
    s = int(s)
    ups = [i for i in range(n_sites) if (s >> i) & 1]
    dns = [i for i in range(n_sites) if not (s >> i) & 1]
    acc = 0.0
    for u in ups:
      for d in dns:
        t = s ^ ((1 << u) | (1 << d))
        acc += min(1.0, pi[index[t]] / pi[k]) if pi[k] > 0 else 0.0
    total += pi[k] * acc / (n_up * n_dn)
  return float(total)

"""Heisenberg local energy (oracle side; test infrastructure).

Restates operators.HeisenbergBond.build (operators.py:137-169) and
HeisenbergHamiltonian.build / local_value / apply_in_place
(operators.py:227-271) with per-bond couplings (a HeisenbergHamiltonian with
uniform (j_x, j_z) is the special case of constant arrays).
"""
import numpy as np
import torch


def bond_terms(configs, bond, j_x, j_z):
  """HeisenbergBond.build, operators.py:154-169, without the amplitude call:
  returns (0.25 * j_z * s_i s_j, 0.25 * j_x * 2 * mask, updated_config); the
  caller multiplies the second term by the amplitude of updated_config."""
  i, j = int(bond[0]), int(bond[1])
  si = configs[:, i].clone()
  sj = configs[:, j].clone()
  updated = configs.clone()
  updated[:, i] += sj - si                      # :158-164 (scatter + add_n)
  updated[:, j] += si - sj
  sz = si * sj                                  # :165
  mask = (sz < 0).to(configs.dtype)             # :166-167
  return 0.25 * j_z * sz, 0.25 * j_x * 2.0 * mask, updated


def build(configs, bonds_ij, jx, jz, psi_fn):
  """HeisenbergHamiltonian.build (operators.py:227-247) in amplitude form:
  returns (diag[B], offdiag[B]) with offdiag = sum_b 0.5 jx mask psi(flipped).
  Evaluates psi on every bond like the reference (operators.py:168)."""
  diag = torch.zeros(configs.shape[0], dtype=configs.dtype)
  off = None
  for k in range(len(bonds_ij)):
    d, pref, updated = bond_terms(configs, bonds_ij[k], float(jx[k]),
                                  float(jz[k]))
    term = pref * psi_fn(updated)
    diag = diag + d
    off = term if off is None else off + term
  return diag, off


def local_energy(configs, bonds_ij, jx, jz, log_amp_fn):
  """local_value (operators.py:249-259) in log form:
  E_loc = sum_b [ jz/4 s_i s_j + jx/2 [s_i != s_j] exp(z(flipped) - z) ].
  Only antiparallel bonds are evaluated (the others are multiplied by a zero
  mask in the reference)."""
  dtype = log_amp_fn(configs[:1]).dtype
  cfg = configs.to(dtype)
  z = log_amp_fn(cfg)
  e = torch.zeros(cfg.shape[0], dtype=dtype)
  for k in range(len(bonds_ij)):
    d, pref, updated = bond_terms(cfg, bonds_ij[k], float(jx[k]),
                                  float(jz[k]))
    e = e + d
    act = pref != 0
    if bool(act.any()):
      zf = log_amp_fn(updated[act])
      e[act] = e[act] + pref[act] * torch.exp(zf - z[act])
  return e


def apply_in_place(configs, bonds_ij, jx, jz, psi_fn):
  """operators.py:261-271: diag * psi + offdiag."""
  diag, off = build(configs, bonds_ij, jx, jz, psi_fn)
  return diag * psi_fn(configs) + off


def n_active(configs, bonds_ij):
  """Number of antiparallel bonds per walker (the measured `n_active` that
  the roofline bookkeeping of SURVEY.md 8(d) asks for)."""
  c = np.asarray(configs)
  ij = np.asarray(bonds_ij)
  return (c[:, ij[:, 0]] * c[:, ij[:, 1]] < 0).sum(axis=1)

"""Philox4x32-10 and the sampler's fast-mode proposal rule in numpy (oracle
side; test infrastructure).

The CUDA sampler (cgsvmc_mc_steps) does not draw B*N uniforms per step like
graph_builders.py:59; it draws one Philox block per (walker, step) and picks
"the k-th up site" and "the k-th down site", which has the same distribution
as the argmax / argmin of sigma * u (a uniformly random up site and an
independent uniformly random down site).  This module restates that rule so
that the GPU trajectories can be checked move by move.
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
  """Vectorised over numpy arrays of uint64 holding 32-bit values."""
  c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)]
  k0 = np.uint64(k0) & _MASK
  k1 = np.uint64(k1) & _MASK
  for _ in range(10):
    p0 = _M0 * c0
    p1 = _M1 * c2
    n0 = ((p1 >> _S32) ^ c1 ^ k0) & _MASK
    n1 = p1 & _MASK
    n2 = ((p0 >> _S32) ^ c3 ^ k1) & _MASK
    n3 = p0 & _MASK
    c0, c1, c2, c3 = n0, n1, n2, n3
    k0 = (k0 + _W0) & _MASK
    k1 = (k1 + _W1) & _MASK
  return c0, c1, c2, c3


def walker_step_random(seed, walker_ids, step):
  """Counter = (step_lo, step_hi, walker_lo, walker_hi), key = seed."""
  walker_ids = np.asarray(walker_ids, dtype=np.uint64)
  step = np.uint64(step)
  seed = np.uint64(seed)
  return philox4x32_10(np.full_like(walker_ids, step & _MASK),
                       np.full_like(walker_ids, step >> _S32),
                       walker_ids & _MASK, walker_ids >> _S32,
                       seed & _MASK, seed >> _S32)


def u32_to_unit(r):
  return ((np.asarray(r, dtype=np.uint64) >> np.uint64(8)).astype(np.float64)
          / 16777216.0)


def fast_proposal(configs, seed, walker_ids, step):
  """Returns (down_site, up_site, u_accept) for every walker.

  up_site  = k_up-th site (ascending) with spin +1, k_up = (r0 * n_up) >> 32
  down_site = k_dn-th site with spin -1,            k_dn = (r1 * n_dn) >> 32
  u_accept = (r2 >> 8) / 2^24; the move (+2 at down_site, -2 at up_site) is
  accepted iff |psi'/psi|^2 > u_accept (graph_builders.py:75-79 squared).
  """
  configs = np.asarray(configs)
  r0, r1, r2, _ = walker_step_random(seed, walker_ids, step)
  b = configs.shape[0]
  up_site = np.empty(b, dtype=np.int64)
  down_site = np.empty(b, dtype=np.int64)
  for w in range(b):
    ups = np.flatnonzero(configs[w] > 0)
    dns = np.flatnonzero(configs[w] < 0)
    k_up = int((int(r0[w]) * len(ups)) >> 32)
    k_dn = int((int(r1[w]) * len(dns)) >> 32)
    up_site[w] = ups[k_up]
    down_site[w] = dns[k_dn]
  return down_site, up_site, u32_to_unit(r2)

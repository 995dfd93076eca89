"""Integer / bit work of the hot path, numpy (oracle side; test infrastructure).

Bit convention shared with the CUDA library (include/cgsvmc.h): a walker's
configuration is ``W = ceil(N / 64)`` little-endian uint64 words; site ``i``
lives in word ``i >> 6`` at bit ``i & 63``; bit = 1 means spin +1 (up), bit = 0
means spin -1 (down).  Unused high bits are 0.
"""
import numpy as np


def n_words(n_sites):
  return (n_sites + 63) // 64


def pack(configs):
  """[B, N] array of +-1 -> uint64 [B, W]."""
  configs = np.asarray(configs)
  b, n = configs.shape
  w = n_words(n)
  out = np.zeros((b, w), dtype=np.uint64)
  up = configs > 0
  for i in range(n):
    out[:, i >> 6] |= up[:, i].astype(np.uint64) << np.uint64(i & 63)
  return out


def unpack(packed, n_sites, dtype=np.float32):
  """uint64 [B, W] -> [B, N] array of +-1."""
  packed = np.asarray(packed, dtype=np.uint64)
  b = packed.shape[0]
  out = np.empty((b, n_sites), dtype=dtype)
  for i in range(n_sites):
    bit = (packed[:, i >> 6] >> np.uint64(i & 63)) & np.uint64(1)
    out[:, i] = np.where(bit == 1, 1, -1)
  return out


def random_sz0_configs(n_sites, batch_size, rng):
  """Uniformly random configurations with n_sites // 2 spins down.

  Same distribution as utils.random_configurations (utils.py:169-192): start
  from all +1 and set n_sites // 2 distinct, uniformly chosen sites to -1 (the
  reference draws them by rejection, we draw a permutation prefix).
  """
  configs = np.ones((batch_size, n_sites), dtype=np.float32)
  for b in range(batch_size):
    down = rng.permutation(n_sites)[: n_sites // 2]
    configs[b, down] = -1.0
  return configs


def flip_enum(packed, bonds_ij, n_sites):
  """Bond enumeration + flipped configurations, operators.py:154-167.

  For every walker b and bond k = (i, j):
    active[b, k]  = 1 iff s_i * s_j < 0                 (operators.py:165-167)
    flipped[b, k] = configuration with s_i and s_j exchanged
                    (operators.py:158-164; for an antiparallel pair that is the
                    XOR of both bits, for a parallel pair it is the identity).
  Returns (active_mask uint32 [B, ceil(n_bonds/32)], flipped uint64
  [B, n_bonds, W]).  Bond k is bit ``k & 31`` of mask word ``k >> 5``.
  """
  packed = np.asarray(packed, dtype=np.uint64)
  b, w = packed.shape
  ij = np.asarray(bonds_ij, dtype=np.int64).reshape(-1, 2)
  nb = ij.shape[0]
  mask = np.zeros((b, (nb + 31) // 32), dtype=np.uint32)
  flipped = np.repeat(packed[:, None, :], nb, axis=1).copy()
  one = np.uint64(1)
  for k in range(nb):
    i, j = int(ij[k, 0]), int(ij[k, 1])
    bi = (packed[:, i >> 6] >> np.uint64(i & 63)) & one
    bj = (packed[:, j >> 6] >> np.uint64(j & 63)) & one
    anti = (bi ^ bj).astype(np.uint64)
    mask[:, k >> 5] |= (anti.astype(np.uint32) << np.uint32(k & 31))
    flipped[:, k, i >> 6] ^= anti << np.uint64(i & 63)
    flipped[:, k, j >> 6] ^= anti << np.uint64(j & 63)
  return mask, flipped


def flip_enum_dense(configs, bonds_ij):
  """The same enumeration stated literally on the +-1 float layout, i.e. the
  scatter form of operators.py:154-167, used to cross-check `flip_enum`."""
  configs = np.asarray(configs, dtype=np.float32)
  ij = np.asarray(bonds_ij, dtype=np.int64).reshape(-1, 2)
  b, n = configs.shape
  nb = ij.shape[0]
  active = np.zeros((b, nb), dtype=bool)
  updated = np.repeat(configs[:, None, :], nb, axis=1).copy()
  for k in range(nb):
    i, j = int(ij[k, 0]), int(ij[k, 1])
    si = configs[:, i].copy()
    sj = configs[:, j].copy()
    updated[:, k, i] += sj - si
    updated[:, k, j] += si - sj
    active[:, k] = (si * sj) < 0
  return active, updated

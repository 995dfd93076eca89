"""Shared pieces of the three command-line drivers (run_training.py,
run_supervised_training.py, run_energy_evaluation.py at the repository root),
which re-host the reference's drivers on the CUDA path."""
import os

import numpy as np
import torch


def init_distributed():
  """One process per GPU under torchrun; a no-op otherwise.  Returns rank."""
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if torch.cuda.is_available():
    torch.cuda.set_device(local_rank)
  if world > 1:
    import torch.distributed as dist
    if not dist.is_initialized():
      dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    return dist.get_rank()
  return 0


def load_bonds(checkpoint_dir, n_sites, heisenberg_jx):
  """Bonds from <checkpoint_dir>/J.txt or the 1-D periodic default
  (run_training.py:103-109).  J.txt holds two integer columns; an optional
  third (and fourth) column gives a per-bond j_x (and j_z) -- the extension
  needed for J1-J2.  Returns (bonds, j_x, j_z)."""
  path = os.path.join(checkpoint_dir, 'J.txt')
  if os.path.exists(path):
    data = np.atleast_2d(np.genfromtxt(path))
    bonds = [(int(r[0]), int(r[1])) for r in data]
    if data.shape[1] >= 3:
      j_x = data[:, 2].astype(np.float32)
      j_z = data[:, 3].astype(np.float32) if data.shape[1] >= 4 else np.abs(j_x)
      return bonds, j_x, j_z
    return bonds, heisenberg_jx, 1.0
  return [(i, (i + 1) % n_sites) for i in range(n_sites)], heisenberg_jx, 1.0

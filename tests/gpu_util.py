"""Helpers shared by the -m gpu parity tests (CUDA library vs oracle)."""
import numpy as np
import torch

from oracle import ansatz as oansatz
from oracle import bits


def make_native(spec, flat_params):
  from cgs_vmc_b200 import _native
  a = _native.Ansatz(spec.kind, spec.n_sites, num_layers=spec.num_layers,
                     layer_size=spec.layer_size, num_filters=spec.num_filters,
                     kernel_size=spec.kernel_size, size_x=spec.size_x,
                     size_y=spec.size_y, nonlinearity=spec.nonlinearity)
  a.set_params(torch.as_tensor(np.asarray(flat_params, dtype=np.float32)))
  return a


def packed_cuda(configs_np):
  """+-1 numpy [B, N] -> packed int64 CUDA tensor via the oracle packer."""
  p = bits.pack(configs_np).view(np.int64)
  return torch.from_numpy(p).cuda()


def unpack_np(packed_t, n_sites):
  return bits.unpack(packed_t.cpu().numpy().view(np.uint64), n_sites)


def oracle_params(spec, flat_params, dtype=torch.float64):
  return oansatz.unflatten(spec, torch.as_tensor(np.asarray(flat_params)).to(dtype))


def amp_scale(spec, params, cfg):
  """Sum of |terms| entering z: the natural scale of float32 rounding error of
  a forward pass (used to state tolerances)."""
  z = oansatz.log_amp(spec, [p.abs() for p in params], cfg.abs())
  return z.abs().numpy() + 1.0

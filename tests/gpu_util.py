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


def relu_kink_band(spec, params, cfg64, weights64, eps=2e-5, rel=1e-5):
  """Kink-aware tolerance for gradients of relu networks evaluated with
  float32-grade (not float64) forward passes: d relu / dx jumps at 0, so a
  pre-activation within rounding of 0 may legitimately be switched either way
  and then changes the gradient by that unit's whole contribution.  Returns
  |g(+eps) - g(-eps)| per entry, where g(s) is the float64 oracle gradient
  with the kink moved to s * eps: zero unless some unit lies within eps of its
  kink, and then exactly the spread such units can cause.  The kink distance
  is max(eps, rel * max |x|) per pre-activation tensor: the forward error of
  a 22-bit-plane product grows with the magnitude of the layer."""
  import torch
  from oracle import estimators
  if spec.nonlinearity != 'relu':
    return 0.0
  saved = oansatz.NONLINEARITIES['relu']
  outs = []
  try:
    for sgn in (1.0, -1.0):
      oansatz.NONLINEARITIES['relu'] = lambda x, s=sgn: torch.where(
          x > s * max(eps, rel * float(x.abs().max())), x, torch.zeros_like(x))
      outs.append(estimators.weighted_grad_sum(spec, params, cfg64, weights64).numpy())
  finally:
    oansatz.NONLINEARITIES['relu'] = saved
  return np.abs(outs[0] - outs[1])


def kink_free_configs(spec, params, n, rng, margin=1e-4, chunk=1024, max_chunks=40):
  """n random Sz = 0 configurations on which no relu pre-activation of the
  float64 oracle lies within margin * max |x| (per layer) of the kink: there
  the gradient does not depend on how the forward pass rounds, and the strict
  float32 tolerances apply to any forward precision."""
  import torch
  from oracle import bits
  assert spec.nonlinearity == 'relu'
  saved = oansatz.NONLINEARITIES['relu']
  keep = []
  try:
    for _ in range(max_chunks):
      cand = bits.random_sz0_configs(spec.n_sites, chunk, rng)
      clear = torch.ones(chunk, dtype=torch.bool)

      def probe(x):
        nonlocal clear
        a = x.detach().abs().reshape(chunk, -1)
        clear &= a.min(dim=1).values > margin * float(a.max())
        return saved(x)
      oansatz.NONLINEARITIES['relu'] = probe
      with torch.no_grad():
        oansatz.log_amp(spec, params, torch.from_numpy(cand).to(params[0].dtype))
      keep.extend(cand[clear.numpy()])
      if len(keep) >= n:
        break
  finally:
    oansatz.NONLINEARITIES['relu'] = saved
  assert len(keep) >= n, 'only %d kink-free configurations found' % len(keep)
  return np.stack(keep[:n])

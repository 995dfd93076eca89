"""Metropolis exchange step (oracle side; test infrastructure).

Restates graph_builders.build_monte_carlo_sampling (graph_builders.py:38-89)
in replay mode: the two `tf.random_uniform` draws (lines 59 and 77) are
explicit inputs so the CUDA replay kernel can be compared move by move.
"""
import torch


def propose(configs: torch.Tensor, u_sites: torch.Tensor):
  """graph_builders.py:59-73.  configs, u_sites: [B, N].  Returns
  (down_site[B], up_site[B], updated_config[B, N]): `down_site` is a uniformly
  random site holding -1 (argmin of sigma * u), `up_site` a uniformly random
  site holding +1 (argmax); the proposal adds +2 at down_site and -2 at
  up_site.  argmin / argmax return the first occurrence on ties."""
  swap_choice = configs * u_sites
  # torch.argmin/argmax do not guarantee first-occurrence on ties; use the
  # explicit definition.
  mn = swap_choice.min(dim=1, keepdim=True).values
  mx = swap_choice.max(dim=1, keepdim=True).values
  n = configs.shape[1]
  idx = torch.arange(n).expand_as(swap_choice)
  big = torch.full_like(idx, n)
  down = torch.where(swap_choice == mn, idx, big).min(dim=1).values
  up = torch.where(swap_choice == mx, idx, big).min(dim=1).values
  updated = configs.clone()
  rows = torch.arange(configs.shape[0])
  updated[rows, down] += 2.0
  updated[rows, up] -= 2.0
  return down, up, updated


def mc_step(configs, u_sites, u_acc, log_amp_fn):
  """One full step, graph_builders.py:54-89.

  log_amp_fn maps [B, N] -> z[B] (log-amplitudes); the reference compares
  |psi'| / |psi| with sqrt(u) (lines 75-79, strict `>`), evaluated here as
  exp(z' - z) > sqrt(u) in the dtype of z.  Returns (new_configs, accept_mask
  [B] bool, log_ratio[B], down, up).
  """
  down, up, updated = propose(configs, u_sites)
  z = log_amp_fn(configs)
  z_new = log_amp_fn(updated)
  log_ratio = z_new - z
  ratios = torch.exp(log_ratio)
  accept = ratios > torch.sqrt(u_acc.to(ratios.dtype))
  new_configs = torch.where(accept[:, None], updated, configs)
  return new_configs, accept, log_ratio, down, up


def mc_step_reference_form(configs, u_sites, u_acc, psi_fn):
  """The same step written with amplitudes (not logs), exactly the op order of
  graph_builders.py:54-89; this is the form the CPU baseline times."""
  batch = configs.shape[0]
  psi = psi_fn(configs)                                       # :54-55
  swap_choice = configs * u_sites                             # :60
  rows = torch.arange(batch)
  down = torch.argmin(swap_choice, 1)                         # :62-63
  up = torch.argmax(swap_choice, 1)                           # :64-65
  spin_down_update = torch.zeros_like(configs)                # :67-68
  spin_down_update[rows, down] = 2.0
  spin_up_update = torch.zeros_like(configs)                  # :70-71
  spin_up_update[rows, up] = -2.0
  updated = configs + spin_down_update + spin_up_update       # :73
  new_psi = psi_fn(updated)                                   # :74
  ratios = new_psi.abs() / psi.abs()                          # :75
  rnd = torch.sqrt(u_acc)                                     # :76-77
  mask = (ratios > rnd).to(configs.dtype)                     # :79
  acc_down = torch.zeros_like(configs)                        # :81-84
  acc_down[rows, down] = 2.0 * mask
  acc_up = torch.zeros_like(configs)
  acc_up[rows, up] = -2.0 * mask
  return configs + acc_down + acc_up, mask.sum()              # :86-88

"""Reference-equivalent CPU path (oracle side; the timed CPU baseline).

TensorFlow 1.x cannot be installed here, so the reference's CPU path is timed
through an op-for-op torch-CPU float32 restatement of the graph the reference
builds -- deliberately INCLUDING its redundancies (BASELINE.md
section 3):
  * two full forward passes per Metropolis step (graph_builders.py:54-55, 74),
  * B*N + B uniforms per step (graph_builders.py:59, 77),
  * dense [B, N] scatter updates (graph_builders.py:67-84),
  * a full forward pass on every bond before masking (operators.py:164-168),
  * two backward passes per gradient accumulation (training.py:545-548),
  * one Python dispatch per step (training.py:608-609, 616-617).
All host threads are used (torch intra-op pool).  Only bench.py's
cpu_baseline / --impl reference legs call this module.
"""
import time

import torch

from . import ansatz as _ansatz
from . import sampler as _sampler


class ReferenceEquivalentVMC:
  """One walker batch + ansatz + Hamiltonian, stepped the reference's way."""

  def __init__(self, spec, flat_params, bonds_ij, jx, jz, configs, shift=-10.0,
               seed=0):
    self.spec = spec
    self.params = [p.clone().float().requires_grad_(True)
                   for p in _ansatz.unflatten(spec, torch.as_tensor(flat_params).float())]
    self.bonds = [(int(a), int(b)) for a, b in bonds_ij]
    self.jx = [float(x) for x in jx]
    self.jz = [float(x) for x in jz]
    self.configs = torch.as_tensor(configs).float().clone()
    self.shift = shift
    self.gen = torch.Generator().manual_seed(seed)

  def psi(self, configs):
    return torch.exp(_ansatz.log_amp(self.spec, self.params, configs,
                                     literal_log_cosh=True) - self.shift)

  def mc_step(self):
    """session.run(mc_step), graph_builders.py:38-89."""
    with torch.no_grad():
      b, n = self.configs.shape
      u_sites = torch.rand((b, n), generator=self.gen)
      u_acc = torch.rand((b,), generator=self.gen)
      self.configs, count = _sampler.mc_step_reference_form(
          self.configs, u_sites, u_acc, self.psi)
    return count

  def accumulate(self):
    """session.run(accumulate_gradients), training.py:539-558 for one batch:
    psi, local energy (forward on every bond), two tf.gradients."""
    configs = self.configs
    psi = self.psi(configs)
    psi_ng = psi.detach()
    with torch.no_grad():     # local_energy is wrapped in stop_gradient
      rows = torch.arange(configs.shape[0])
      diag = torch.zeros_like(psi_ng)
      off = torch.zeros_like(psi_ng)
      for (i, j), jx, jz in zip(self.bonds, self.jx, self.jz):
        si, sj = configs[:, i], configs[:, j]
        upd_i = torch.zeros_like(configs)
        upd_i[rows, i] = sj - si
        upd_j = torch.zeros_like(configs)
        upd_j[rows, j] = si - sj
        updated = configs + upd_i + upd_j
        sz = si * sj
        mask = (sz < 0).float()
        diag = diag + 0.25 * jz * sz
        off = off + 0.25 * jx * 2.0 * mask * self.psi(updated)
      e_loc = diag + off / psi_ng
    g1 = torch.autograd.grad((psi / psi_ng).sum(), self.params, retain_graph=True)
    g2 = torch.autograd.grad((psi / psi_ng * e_loc).sum(), self.params)
    return e_loc, g1, g2

  def step(self, n_mc_steps):
    """One batch iteration of training.py:614-617."""
    e_loc, g1, g2 = self.accumulate()
    for _ in range(n_mc_steps):
      self.mc_step()
    return float(e_loc.mean())


def time_steps(vmc, n_mc_steps, steps, warmup):
  for _ in range(warmup):
    vmc.step(n_mc_steps)
  times = []
  for _ in range(steps):
    t0 = time.perf_counter()
    vmc.step(n_mc_steps)
    times.append(time.perf_counter() - t0)
  return times

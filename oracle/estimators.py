"""Gradient / loss estimators (oracle side; test infrastructure).

Restates EnergyGradientOptimizer.build_opt_ops (training.py:531-586) and
SupervisedWavefunctionOptimizer.build_opt_ops (training.py:141-189) on top of
one primitive, S_k = sum_b w_kb * O_b with O_b = d log psi_b / d params
(SURVEY.md appendix A.5), obtained here by torch autograd.
"""
import math

import torch

from . import ansatz as _ansatz


def weighted_grad_sum(spec, params, configs, weights):
  """S[K, P] = sum_b weights[k, b] * d z_b / d params (flat layout).

  training.py:545-548 computes tf.gradients(psi / stop_gradient(psi) * w) which
  sums over the batch: exactly sum_b w_b * d log psi_b."""
  leaves = [p.detach().clone().requires_grad_(True) for p in params]
  z = _ansatz.log_amp(spec, leaves, configs)
  weights = weights.to(z.dtype).reshape(-1, z.shape[0])
  rows = []
  for k in range(weights.shape[0]):
    grads = torch.autograd.grad((weights[k] * z).sum(), leaves,
                                retain_graph=True, allow_unused=True)
    grads = [torch.zeros_like(p) if g is None else g
             for g, p in zip(grads, leaves)]
    rows.append(_ansatz.flatten(grads))
  return torch.stack(rows)


class EnergyGradientAccumulator:
  """The local-variable accumulators of training.py:550-568.

  `accumulate` adds one batch: running mean over *batches* of G1 = sum_b O_b
  and G2 = sum_b E_b O_b (tf.metrics.mean_tensor), and a per-sample running
  mean of E (tf.metrics.mean, training.py:555).  `gradient` returns
  mean(G2) - E_mean * mean(G1) (training.py:562-564) -- B times the textbook
  covariance, kept on purpose."""

  def __init__(self, n_params, dtype=torch.float64):
    self.reset(n_params, dtype)

  def reset(self, n_params=None, dtype=None):
    if n_params is not None:
      self._p, self._dtype = n_params, dtype
    self.g1 = torch.zeros(self._p, dtype=self._dtype)
    self.g2 = torch.zeros(self._p, dtype=self._dtype)
    self.e_sum = 0.0
    self.e_count = 0
    self.n_batches = 0

  def accumulate(self, spec, params, configs, e_loc):
    ones = torch.ones_like(e_loc)
    s = weighted_grad_sum(spec, params, configs, torch.stack([ones, e_loc]))
    self.g1 += s[0].to(self._dtype)
    self.g2 += s[1].to(self._dtype)
    self.e_sum += float(e_loc.sum())
    self.e_count += e_loc.numel()
    self.n_batches += 1

  @property
  def mean_energy(self):
    return self.e_sum / self.e_count

  def gradient(self):
    return self.g2 / self.n_batches - self.mean_energy * self.g1 / self.n_batches


def swo_loss_and_grad(spec, params, configs, psi_target, shift=-10.0):
  """training.py:166-175: loss = mean_b (psi_b - t_b)^2 / sg(psi_b)^2 with
  t = psi_target * sqrt(2^N); d loss = mean_b 2 (1 - t_b / psi_b) O_b."""
  n_sites = configs.shape[1]
  z = _ansatz.log_amp(spec, params, configs)
  psi = torch.exp(z - shift)
  t = psi_target.to(psi.dtype) * math.sqrt(2.0 ** n_sites)
  loss = torch.mean((psi - t) ** 2 / psi ** 2)
  w = 2.0 * (1.0 - t / psi) / configs.shape[0]
  grad = weighted_grad_sum(spec, params, configs, w[None, :])[0]
  return loss, grad


def energy_stats(e_loc):
  """Sum E, sum E^2, count: what K5 packs next to the gradient sums."""
  e = e_loc.to(torch.float64)
  return float(e.sum()), float((e * e).sum()), e.numel()


def log_overlap_grad(spec, params, configs, ratio):
  """LogOverlapSWO / LogOverlapImaginaryTimeSWO gradient for one batch
  (training.py:336-360, 672-699): tf.gradients sums over the batch,
  tf.metrics.mean(ratio) is a per-sample mean, so
    g = sum_b O_b - (sum_b r_b O_b) / mean_b(r_b)."""
  ones = torch.ones_like(ratio)
  s = weighted_grad_sum(spec, params, configs, torch.stack([ones, ratio]))
  return s[0] - s[1] / ratio.mean()


def dual_sampling_loss_and_grad(spec, params, configs, psi_target, shift=-10.0):
  """DualSamplingSWO (training.py:452-466): loss = mean_b (psi_b - t_b)^2 with
  t = psi_target * sqrt(2^N), NOT divided by psi^2; d loss = mean_b 2 (psi_b -
  t_b) psi_b O_b.  `configs` is concat([psi walkers, target walkers])."""
  n_sites = configs.shape[1]
  z = _ansatz.log_amp(spec, params, configs)
  psi = torch.exp(z - shift)
  t = psi_target.to(psi.dtype) * math.sqrt(2.0 ** n_sites)
  loss = torch.mean((psi - t) ** 2)
  w = 2.0 * (psi - t) * psi / configs.shape[0]
  grad = weighted_grad_sum(spec, params, configs, w[None, :])[0]
  return loss, grad


def imaginary_time_ratio(psi, psi_omega, e_loc_omega, beta):
  """training.py:655-670: ratio = (psi_O - beta H psi_O) / psi with
  H psi_O = E_loc[psi_O] psi_O (apply_in_place, operators.py:261-271); the
  supervisor energy estimate is mean(E_loc[psi_O]) (training.py:662, 681)."""
  return psi_omega * (1.0 - beta * e_loc_omega) / psi

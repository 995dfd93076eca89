"""Device-resident state of one walker shard and the per-batch estimator pass.

This is the host-side glue under the reference-shaped API (graph_builders,
operators, training): it owns the torch buffers and calls the C-ABI.  Nothing
here computes on the CPU.
"""
import torch

from . import _native


class WalkerState:
  """The `configs` variable of graph_builders.get_configs
  (graph_builders.py:92-125), kept bit-packed on the device, plus the Philox
  stream position.  `walker_id0` is the global id of the first walker of this
  shard, so trajectories do not depend on the sharding."""

  def __init__(self, batch_size, n_sites, seed=0xC65, walker_id0=0, device='cuda',
               packed=None):
    self.batch_size, self.n_sites = batch_size, n_sites
    self.seed, self.walker_id0 = int(seed), int(walker_id0)
    self.step = 0
    if packed is None:
      packed = _native.random_configs(batch_size, n_sites, self.seed, self.walker_id0,
                                      device=device)
    self.packed = packed
    self.accept_count = torch.zeros(1, dtype=torch.int64, device=packed.device)
    self.proposed = 0
    # device-side copy of the Philox step offset for CUDA-graph replays
    self.step_dev = torch.zeros(1, dtype=torch.int64, device=packed.device)

  def configs(self):
    """float32 [B, N] of +-1: the reference's view of the state."""
    return _native.unpack_configs(self.packed, self.n_sites)

  def set_configs(self, configs):
    if tuple(configs.shape) != (self.batch_size, self.n_sites):
      raise ValueError('Size of existing variable does not match.')   # graph_builders.py:118
    # in place: captured CUDA graphs keep pointing at the same buffer
    _native.pack_configs(configs.to(self.packed.device, torch.float32).contiguous(), out=self.packed)

  def mc_steps(self, ansatz, n_steps, log_amp_out=None):
    """n_steps x session.run(mc_step) (graph_builders.py:38-89) in one launch."""
    ansatz.mc_steps(self.packed, n_steps, self.seed, self.walker_id0, self.step,
                    self.accept_count, log_amp_out)
    self.step += int(n_steps)
    self.proposed += int(n_steps) * self.batch_size

  def mc_steps_graph(self, ansatz, n_steps):
    """The same with the step offset read from `step_dev` on the device
    (capturable); the caller keeps `step_dev` and `step` in sync."""
    ansatz.mc_steps_graph(self.packed, n_steps, self.seed, self.walker_id0, self.step_dev,
                          self.accept_count)


class EnergyGradientSums:
  """The local-variable accumulators of training.py:550-568 as one packed
  float buffer [2, P] (sum_b O_b | sum_b E_b O_b) plus float64 energy
  statistics [sum E, sum E^2, n, 0]; this is also the all-reduce payload of
  the walker-sharded run."""

  def __init__(self, ansatz, batch_size, device='cuda'):
    self.ansatz = ansatz
    self.sums = torch.zeros(2, ansatz.num_params, dtype=torch.float32, device=device)
    self.stats = torch.zeros(4, dtype=torch.float64, device=device)
    self.weights = torch.ones(2, batch_size, dtype=torch.float32, device=device)
    self.log_amp = torch.empty(batch_size, dtype=torch.float32, device=device)
    self.n_batches = 0

  def reset(self):
    """session.run(reset_gradients), training.py:568."""
    self.sums.zero_()
    self.stats.zero_()
    self.n_batches = 0

  def accumulate(self, ham, packed):
    """session.run(accumulate_gradients), training.py:539-558 for one batch in
    one library call: E_loc, S = [sum O, sum E O] and the energy statistics."""
    e_row = self.weights[1]
    self.ansatz.accumulate(ham, packed, self.sums, self.stats, e_loc_out=e_row,
                           log_amp_out=self.log_amp)
    self.n_batches += 1
    return e_row

  def mean_energy(self):
    s = self.stats
    return s[0] / s[2]

  def gradient(self, n_batches=None):
    """mean_batches(sum E O) - mean(E) * mean_batches(sum O), training.py:562-564
    (B times the covariance: tf.gradients sums over the batch)."""
    nb = float(self.n_batches if n_batches is None else n_batches)
    return self.sums[1] / nb - self.mean_energy().float() * self.sums[0] / nb


class GraphedBatchStep:
  """One batch iteration of EnergyGradientOptimizer.run_optimization_epoch
  (training.py:614-617) -- accumulate_gradients followed by
  num_monte_carlo_sweeps * num_sites Metropolis steps -- captured once as a
  CUDA graph and replayed: one graph launch instead of a Python dispatch per
  kernel (the reference pays one session.run per Metropolis step).

  The parameter tables are rebuilt inside the graph, so in-place optimizer
  updates between replays are picked up; the Philox step offset lives in
  device memory and is advanced by the graph itself."""

  def __init__(self, state, ansatz, ham, sums, n_steps):
    self.state, self.ansatz, self.ham, self.sums, self.n_steps = state, ansatz, ham, sums, int(n_steps)
    state.step_dev.fill_(state.step)
    saved = (sums.sums.clone(), sums.stats.clone(), state.packed.clone(), state.accept_count.clone())
    side = torch.cuda.Stream(device=state.packed.device)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):           # warm-up: sizes every scratch buffer
      self._body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    # undo the warm-up so that capture + replays see the caller's state
    sums.sums.copy_(saved[0]); sums.stats.copy_(saved[1])
    state.packed.copy_(saved[2]); state.accept_count.copy_(saved[3])
    state.step_dev.fill_(state.step)
    self.graph = torch.cuda.CUDAGraph()
    _native.check(_native.load().cgsvmc_ansatz_params_changed(ansatz._handle))   # capture the table build
    with torch.cuda.graph(self.graph):
      self._body()
    # capture does not execute: nothing to undo

  def _body(self):
    self.sums.ansatz.accumulate(self.ham, self.state.packed, self.sums.sums, self.sums.stats,
                                e_loc_out=self.sums.weights[1], log_amp_out=self.sums.log_amp)
    self.state.mc_steps_graph(self.ansatz, self.n_steps)

  def replay(self):
    self.graph.replay()
    self.sums.n_batches += 1
    self.state.step += self.n_steps
    self.state.proposed += self.n_steps * self.state.batch_size

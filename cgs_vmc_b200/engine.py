"""Device-resident state of one walker shard and the per-batch estimator pass.

This is the host-side glue under the reference-shaped API (graph_builders,
operators, training): it owns the torch buffers and calls the C-ABI.  Nothing
here computes on the CPU.
"""
import os
import time

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

  def __init__(self, ansatz, batch_size, device='cuda', want_log_amp=False):
    self.ansatz = ansatz
    self.sums = torch.zeros(2, ansatz.num_params, dtype=torch.float32, device=device)
    self.stats = torch.zeros(4, dtype=torch.float64, device=device)
    self.weights = torch.ones(2, batch_size, dtype=torch.float32, device=device)
    # log psi of every walker is a by-product nobody on the training path
    # consumes (the kernels work with amplitude ratios); opt in to get it
    self.log_amp = (torch.empty(batch_size, dtype=torch.float32, device=device)
                    if want_log_amp else None)
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

  def batch_step(self, ham, state, n_steps, on_device_counter=False):
    """accumulate() followed by n_steps Metropolis steps of `state`
    (training.py:614-617) through cgsvmc_batch_step: one fused kernel for the
    pure RBM, the two calls otherwise."""
    e_row = self.weights[1]
    counter = dict(step_counter=state.step_dev) if on_device_counter else dict(step0=state.step)
    self.ansatz.batch_step(ham, state.packed, self.sums, self.stats, n_steps, state.seed,
                           state.walker_id0, accept_count=state.accept_count, e_loc_out=e_row,
                           log_amp_out=self.log_amp, **counter)
    if on_device_counter:      # capturable: the caller keeps state.step in sync
      return e_row
    self.n_batches += 1
    state.step += int(n_steps)
    state.proposed += int(n_steps) * state.batch_size
    return e_row

  def batch_steps(self, ham, state, n_steps, n_batches, on_device_counter=False, e_loc_out=None):
    """n_batches x batch_step() in one library call (cgsvmc_batch_steps): the
    inner loop of run_optimization_epoch (training.py:614-617) as ONE
    persistent kernel for the pure RBM.  e_loc_out: optional float32
    [n_batches, B]."""
    counter = dict(step_counter=state.step_dev) if on_device_counter else dict(step0=state.step)
    self.ansatz.batch_steps(ham, state.packed, n_batches, self.sums, self.stats, n_steps, state.seed,
                            state.walker_id0, accept_count=state.accept_count, e_loc_out=e_loc_out,
                            **counter)
    if on_device_counter:      # capturable: the caller keeps state.step in sync
      return
    self.n_batches += int(n_batches)
    state.step += int(n_steps) * int(n_batches)
    state.proposed += int(n_steps) * int(n_batches) * state.batch_size

  def mean_energy(self):
    s = self.stats
    return s[0] / s[2]

  def gradient(self, n_batches=None):
    """mean_batches(sum E O) - mean(E) * mean_batches(sum O), training.py:562-564
    (B times the covariance: tf.gradients sums over the batch)."""
    nb = float(self.n_batches if n_batches is None else n_batches)
    return self.sums[1] / nb - self.mean_energy().float() * self.sums[0] / nb


class _CapturedStep:
  """Shared machinery of the captured batch steps: warm-up on a side stream
  (sizes every scratch buffer), undo, two captures per variant -- with the
  parameter-table build (replayed after a parameter update, detected through
  the torch version counter of the parameter buffer) and without it -- and the
  bookkeeping of the device-side Philox step counter."""

  def _prepare(self, state, ansatz, ham, sums, n_steps, variants):
    self.state, self.ansatz, self.ham, self.sums, self.n_steps = state, ansatz, ham, sums, int(n_steps)
    dev = state.packed.device
    state.step_dev.fill_(state.step)
    saved = (sums.sums.clone(), sums.stats.clone(), state.packed.clone(), state.accept_count.clone())
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      self._body(variants[0])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    # undo the warm-up so that capture + replays see the caller's state
    sums.sums.copy_(saved[0]); sums.stats.copy_(saved[1])
    state.packed.copy_(saved[2]); state.accept_count.copy_(saved[3])
    state.step_dev.fill_(state.step)
    self.graphs = {}
    for rebuild in (True, False):
      for v in variants:
        if rebuild:
          _native.check(_native.load().cgsvmc_ansatz_params_changed(ansatz._handle))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
          self._body(v)
        self.graphs[(rebuild, v)] = g
    # capture does not execute: nothing to undo, but the tables the handle now
    # believes valid were never built -- the first replay must be a rebuilding one
    self._params_version = None
    self._expected_step = state.step

  def _replay(self, variant):
    if self.state.step != self._expected_step:
      # steps were taken outside the graph (equilibration): resync the device counter
      self.state.step_dev.fill_(self.state.step)
    version = self.ansatz.params._version
    rebuild = version != self._params_version
    self._params_version = version
    self.graphs[(rebuild, variant)].replay()
    self._expected_step = self.state.step + self.n_steps
    self.sums.n_batches += 1
    self.state.step += self.n_steps
    self.state.proposed += self.n_steps * self.state.batch_size


class GraphedEpoch(_CapturedStep):
  """The whole inner loop of run_optimization_epoch (training.py:614-617),
  `n_batches` batch iterations, captured as one CUDA graph holding
  cgsvmc_batch_steps: for the pure RBM one persistent cooperative kernel per
  epoch (tables loaded once, walkers resident in registers across the
  iterations, one cross-CTA reduction)."""

  def __init__(self, state, ansatz, ham, sums, n_steps, n_batches):
    self.n_batches = int(n_batches)
    self._prepare(state, ansatz, ham, sums, n_steps, (0,))

  def _body(self, variant):
    self.sums.batch_steps(self.ham, self.state, self.n_steps, self.n_batches, on_device_counter=True)

  def replay(self):
    self._replay(0)
    # _replay booked one iteration
    extra = self.n_batches - 1
    self.sums.n_batches += extra
    self.state.step += extra * self.n_steps
    self.state.proposed += extra * self.n_steps * self.state.batch_size
    self._expected_step = self.state.step


class GraphedBatchStep(_CapturedStep):
  """One batch iteration of EnergyGradientOptimizer.run_optimization_epoch
  (training.py:614-617) -- accumulate_gradients followed by
  num_monte_carlo_sweeps * num_sites Metropolis steps -- captured once as a
  CUDA graph and replayed: one graph launch instead of a Python dispatch per
  kernel (the reference pays one session.run per Metropolis step).

  The graph holds cgsvmc_batch_step: the fused estimator + sweep kernel and the
  deterministic reduction (which also advances the Philox step counter)."""

  def __init__(self, state, ansatz, ham, sums, n_steps):
    self._prepare(state, ansatz, ham, sums, n_steps, (0,))

  def _body(self, variant):
    self.sums.batch_step(self.ham, self.state, self.n_steps, on_device_counter=True)

  def replay(self):
    self._replay(0)


class HostFedBatchStep(_CapturedStep):
  """The batch step for a caller that keeps the reference's float32 [B, N]
  configuration tensor in HOST memory (graph_builders.py:92-125 viewed from
  outside the session): every submit() uploads one pinned host batch on a copy
  stream (double-buffered: the upload of batch k overlaps the compute of batch
  k - 1) and replays one captured graph per buffer slot holding
  cgsvmc_batch_step_fed -- the walker kernel bit-packs the float32
  configurations itself and stores the step's result, the energy statistics
  (sum E, sum E^2, n: the metric the reference reads back, training.py:619-620),
  straight into pinned host memory, so neither a packing launch nor a copy
  node sits between two steps.  result() blocks until the oldest outstanding
  batch has landed.  The [2, P] gradient sums are accumulated on the device
  like the reference's local variables; fetch_sums() copies them out."""

  def __init__(self, state, ansatz, ham, sums, n_steps, host_pack='auto'):
    dev = state.packed.device
    B, N, P = state.batch_size, state.n_sites, ansatz.num_params
    self.copy_stream = torch.cuda.Stream(device=dev)
    # float32 host input: either uploaded as it is (4 N bytes per walker, packed
    # by the walker kernel) or bit-packed by the host cores into a pinned staging
    # buffer first (cgsvmc_pack_configs_host: 8 bytes per 64 sites over PCIe).
    # 'auto' measures both on this box and packs unless the host cores are far slower than the link.
    self.host_staging = [torch.zeros(B, _native.n_words(N), dtype=torch.int64).pin_memory() for _ in range(2)]
    # packing threads: the ranks of one box share its cores (8 ranks x 8 threads on 16 cores
    # measured 98 us per e2e step against 74 us for one rank)
    ranks_here = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1')))
    self.pack_threads = max(1, min(8, (os.cpu_count() or 8) // ranks_here))
    self.host_pack_probe = None
    if host_pack == 'auto':
      host_pack = self._probe_host_pack(B, N, dev)
    self.host_pack = bool(host_pack)
    self.dev_cfg = [torch.empty(B, N, dtype=torch.float32, device=dev) for _ in range(2)]
    self.host_stats = [torch.zeros(4, dtype=torch.float64).pin_memory() for _ in range(2)]
    self.host_sums = torch.empty(2, P, dtype=torch.float32).pin_memory()
    self.uploaded = [torch.cuda.Event() for _ in range(2)]
    self.landed = [torch.cuda.Event() for _ in range(2)]
    self._h2d_bytes = (B * N * 4, B * _native.n_words(N) * 8)      # float32 upload, host-packed upload
    self.d2h_bytes_stats = 32
    self.d2h_bytes_sums = 2 * P * 4
    self._submitted = 0
    self._collected = 0
    for c in self.dev_cfg:
      c.copy_(state.configs())
    # one walker buffer per slot (after a submit `state.packed` is rebound to
    # the buffer just stepped); variants 0 / 1 step a slot from its float32
    # upload, 2 / 3 from a bit-packed upload
    self.slot_packed = [state.packed.clone() for _ in range(2)]
    self._prepare(state, ansatz, ham, sums, n_steps, (0, 1, 2, 3))
    for p in self.slot_packed:               # the warm-up stepped slot 0
      p.copy_(state.packed)
    main = torch.cuda.current_stream()
    for ev in self.landed:
      ev.record(main)

  def _body(self, variant):
    slot, from_packed = variant & 1, variant >= 2
    st = self.state
    self.ansatz.batch_step_fed(self.ham, None if from_packed else self.dev_cfg[slot],
                               self.slot_packed[slot], self.sums.sums, self.sums.stats, self.n_steps,
                               st.seed, st.walker_id0, st.step_dev, accept_count=st.accept_count,
                               e_loc_out=self.sums.weights[1], stats_out=self.host_stats[slot])

  @property
  def h2d_bytes(self):
    return self._h2d_bytes[1 if self.host_pack else 0]

  def _probe_host_pack(self, B, N, dev):
    """True when packing a float32 [B, N] batch on the host cores takes less
    time than the extra bytes of its float32 upload on this box."""
    host = torch.ones(B, N, dtype=torch.float32).pin_memory()
    dst = torch.empty(B, N, dtype=torch.float32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(4):
      dst.copy_(host, non_blocking=True)
    e1.record()
    e1.synchronize()
    t_h2d = e0.elapsed_time(e1) * 1e-3 / 4
    # packing threads: the fastest of 1 / 2 / 4 / 8 within this rank's share of the cores (a
    # container with a CPU quota below its visible core count is faster single-threaded)
    timings = {}
    for threads in sorted({1, 2, 4, 8, self.pack_threads}):
      if threads > self.pack_threads:
        continue
      _native.pack_configs_host(host, self.host_staging[0], threads)
      t0 = time.perf_counter()
      for _ in range(6):
        _native.pack_configs_host(host, self.host_staging[0], threads)
      timings[threads] = (time.perf_counter() - t0) / 6
    self.pack_threads = min(timings, key=timings.get)
    t_pack = timings[self.pack_threads]
    self.host_pack_probe = {'h2d_float32_us': t_h2d * 1e6, 'host_pack_us': t_pack * 1e6,
                            'h2d_float32_gbps': B * N * 4 / t_h2d / 1e9,
                            'host_pack_us_by_threads': {str(k): v * 1e6 for k, v in timings.items()}}
    # Measured on this pool (profiles/r02A_bench_steps{20,200}.json): even on a
    # box whose link moves the float32 batch in 27 us the packed form wins (69
    # against 76 us per step in steady state, 73 against 98 us over the first
    # 20 steps after an idle period -- the link and the in-kernel packing of
    # 1.18 MB cost more than 46 us of host cores), so pack unless the host
    # cores are pathologically slow next to the link
    return t_pack < t_h2d + 60e-6

  def submit(self, host_configs):
    """host_configs: pinned host tensor, either the reference's float32 [B, N]
    of +-1 or the library's own walker layout, int64 [B, ceil(N / 64)]
    bit-packed (8 bytes per walker up to 64 sites instead of 4 N).
    Asynchronous."""
    slot = self._submitted & 1
    main = torch.cuda.current_stream()
    from_packed = host_configs.dtype == torch.int64 or self.host_pack
    self.copy_stream.wait_event(self.landed[slot])
    if host_configs.dtype == torch.int64:
      if tuple(host_configs.shape) != tuple(self.slot_packed[slot].shape):
        raise ValueError('Size of existing variable does not match.')
      with torch.cuda.stream(self.copy_stream):
        self.slot_packed[slot].copy_(host_configs, non_blocking=True)
    else:
      if (tuple(host_configs.shape) != tuple(self.dev_cfg[slot].shape) or host_configs.dtype != torch.float32 or
          host_configs.is_cuda or not host_configs.is_contiguous()):
        raise ValueError('Size of existing variable does not match.')
      if self.host_pack:
        self.uploaded[slot].synchronize()        # the copy engine is done with this staging buffer
        _native.upload_configs(host_configs, self.slot_packed[slot], self.copy_stream,
                               staging=self.host_staging[slot], n_threads=self.pack_threads)
      else:
        _native.upload_configs(host_configs, self.dev_cfg[slot], self.copy_stream)
    self.uploaded[slot].record(self.copy_stream)
    main.wait_event(self.uploaded[slot])
    self._replay(slot + (2 if from_packed else 0))
    self.state.packed = self.slot_packed[slot]
    # one event: the step's statistics have landed on the host AND the slot's
    # buffers may be refilled
    self.landed[slot].record(main)
    self._submitted += 1

  def outstanding(self):
    return self._submitted - self._collected

  def result(self):
    """Host energy statistics [sum E, sum E^2, n, 0] (cumulative since the last
    reset) as of the oldest outstanding batch."""
    if self._collected >= self._submitted:
      raise RuntimeError('no batch outstanding')
    slot = self._collected & 1
    self.landed[slot].synchronize()
    self._collected += 1
    return self.host_stats[slot]

  def fetch_sums(self):
    """The [2, P] gradient sums accumulated so far, in pinned host memory
    (synchronises)."""
    self.host_sums.copy_(self.sums.sums, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return self.host_sums

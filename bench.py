#!/usr/bin/env python
"""Benchmark of the VMC hot path (BASELINE.json metric: walker-steps/sec and
E_loc evals/sec), workload C2 = 6x6 Heisenberg, RBM (H = 144), 8192 walkers
per GPU.

A "step" is one batch iteration of EnergyGradientOptimizer.run_optimization_epoch
(training.py:614-617): `accumulate_gradients` (local energy of every walker,
the two gradient sums, energy statistics) followed by one Monte-Carlo sweep
(num_monte_carlo_sweeps * num_sites = 36 Metropolis steps per walker).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints one JSON line (see DESIGN.md section "Measurement" for every field).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOAD = 'C2: 6x6 square-lattice Heisenberg (72 NN bonds, jx=-1, jz=1), rbm H=144, 8192 walkers/GPU'
N_SITES, SIZE, HIDDEN, WALKERS = 36, 6, 144, 8192
SWEEP_STEPS = N_SITES            # num_monte_carlo_sweeps (1) * num_sites
SEED = 0xC65
EPOCH_BATCHES = 50               # num_batches_per_epoch default (utils.py:132-138)
L2_FLUSH_BYTES = 256 << 20


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=200)
  ap.add_argument('--warmup', type=int, default=20)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--walkers', type=int, default=WALKERS, help='walkers per GPU')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-cuda-graph', dest='cuda_graph', action='store_false',
                  help='launch the kernels of a step one by one instead of replaying a captured graph')
  return ap.parse_args()


def peaks():
  path = os.path.join(REPO, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    p = json.load(open(path))
    return dict(hbm_gbs=p['hbm_gbs'], sm_max_mhz=p.get('sm_max_mhz', 1965.0), source='measured')
  return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source='fallback')


def problem():
  """Synthetic inputs of SURVEY.md 8(d): Sonnet-default weights (truncated
  normal, sigma = 1 / sqrt(fan_in), zero biases; torch.Generator seed 1234) in
  the flat layout of include/cgsvmc.h, NN bonds with (jx, jz) = (-1, 1)."""
  from cgs_vmc_b200 import lattices, wavefunctions
  gen = torch.Generator().manual_seed(1234)
  shapes = [(N_SITES, 1), (1,), (N_SITES, HIDDEN), (HIDDEN,)]     # a, a0, W, c
  flat = torch.cat([t.reshape(-1) for t in wavefunctions._sonnet_init(shapes, gen)]).float()
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(SIZE), -1.0, 1.0)
  return flat, ij, jx, jz


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
  """Samples SM clock and throttle reasons through NVML from a background
  thread while the timed region runs (the region lasts milliseconds, far below
  nvidia-smi's own start-up time, so the CLI loop of the profiling recipe
  cannot be used here; the NVML fields are the same ones it prints)."""

  def __init__(self, index, period_s=0.002):
    import threading
    self.samples, self.max_mhz, self.reasons = [], None, set()
    self._stop = threading.Event()
    self._thread = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self._nv = pynvml
      self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
    except Exception:   # NVML unavailable: report nulls
      self._nv = None
      return
    self._period = period_s
    self._thread = threading.Thread(target=self._run, daemon=True)
    self._thread.start()

  def _run(self):
    nv = self._nv
    names = {
        'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown,
        'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
        'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown,
        'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap,
    }
    while not self._stop.is_set():
      try:
        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for k, bit in names.items():
          if mask & bit:
            self.reasons.add(k)
      except Exception:
        pass
      self._stop.wait(self._period)

  def stop(self):
    if self._thread is not None:
      self._stop.set()
      self._thread.join(timeout=2)
    sm = float(np.median(self.samples)) if self.samples else None
    return dict(sm_mhz=sm, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                samples=len(self.samples))


# ----------------------------------------------------------------------------
# CPU arm: the reference-equivalent op sequence on the host cores
# ----------------------------------------------------------------------------
def cpu_arm(walkers, steps, warmup):
  from oracle import ansatz as oansatz
  from oracle import bits, cpu_baseline
  flat, ij, jx, jz = problem()
  spec = oansatz.AnsatzSpec('rbm', N_SITES, num_layers=0, layer_size=HIDDEN, size_x=SIZE, size_y=SIZE)
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  cfg = bits.random_sz0_configs(N_SITES, walkers, np.random.default_rng(1234))
  vmc = cpu_baseline.ReferenceEquivalentVMC(spec, flat, ij, jx, jz, cfg, seed=SEED)
  times = cpu_baseline.time_steps(vmc, SWEEP_STEPS, steps, warmup)
  t = float(np.mean(times))
  return dict(walker_steps_per_sec=walkers * SWEEP_STEPS / t, eloc_evals_per_sec=walkers / t,
              ms_per_step=t * 1e3, cores=cores, walkers=walkers, steps=steps)


def run_reference(args, rank):
  """--impl reference: the reference's CPU implementation of the same step.
  TensorFlow 1.x / Sonnet v1 are not installable in this image, so this is
  the op-for-op torch-CPU restatement (oracle/cpu_baseline.py, kind "port").
  Under torchrun only rank 0 works."""
  if rank != 0:
    return
  # bounded sample: the full batch per step unless K steps of it would exceed
  # ~4 minutes on this host (probed with one step); then 2048 walkers per step
  # and per-walker throughput is what is reported
  probe = cpu_arm(args.walkers, 1, 0)
  budget_s = 240.0
  fits = probe['ms_per_step'] * 1e-3 * (args.steps + args.warmup) <= budget_s
  sample_walkers = args.walkers if fits else min(args.walkers, 2048)
  r = cpu_arm(sample_walkers, args.steps, args.warmup)
  line = {
      'impl': 'reference', 'metric': 'walker_steps_per_sec', 'value': r['walker_steps_per_sec'],
      'unit': 'walker-steps/s', 'eloc_evals_per_sec': r['eloc_evals_per_sec'],
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': WORKLOAD, 'walkers_per_gpu': args.walkers,
                 'mc_steps_per_step': SWEEP_STEPS, 'n_bonds': 72},
      'cpu_baseline': {'value': r['walker_steps_per_sec'], 'unit': 'walker-steps/s',
                       'cores': r['cores'], 'kind': 'port',
                       'sample': '%d steps of the same workload on %d of the %d walkers per step, '
                                 'torch-CPU float32 restatement of the TF graph, all host threads'
                                 % (args.steps, sample_walkers, args.walkers)},
      'e2e': {'value': r['walker_steps_per_sec'], 'unit': 'walker-steps/s',
              'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line))


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
  import torch.distributed as dist
  from cgs_vmc_b200 import _native, engine
  _native.require_cuda()
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  flat, ij, jx, jz = problem()
  B = args.walkers
  ansatz = _native.Ansatz('rbm', N_SITES, num_layers=0, layer_size=HIDDEN, device=dev)
  ansatz.set_params(flat)
  ham = _native.Hamiltonian(ij, jx, jz, N_SITES, device=dev)
  state = engine.WalkerState(B, N_SITES, seed=SEED, walker_id0=rank * B, device=dev)
  sums = engine.EnergyGradientSums(ansatz, B, device=dev)
  P = ansatz.num_params
  payload = torch.zeros(2 * P + 4, dtype=torch.float32, device=dev)   # all-reduce buffer
  flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
  state.mc_steps(ansatz, 20 * N_SITES)                                 # equilibrate

  launches = [0]
  counter = [0]
  ev = lambda: torch.cuda.Event(enable_timing=True)

  graphed = engine.GraphedBatchStep(state, ansatz, ham, sums, SWEEP_STEPS) if args.cuda_graph else None

  def step(events=None):
    """accumulate_gradients + one sweep (+ the packed all-reduce when sharded).
    Timed as one captured CUDA graph (default) or, with --no-cuda-graph /
    when per-phase events are wanted, as separate launches."""
    if graphed is not None and events is None:
      graphed.replay()
      launches[0] += 2   # fused estimator + sweep kernel, reduction (+ table build after a parameter update)
    else:
      if events: events[0].record()
      sums.accumulate(ham, state.packed)     # E_loc + both gradient sums + energy statistics
      if events: events[2].record()
      state.mc_steps(ansatz, SWEEP_STEPS)
      state.step_dev.fill_(state.step)
      if events: events[3].record()
      launches[0] += 3   # walker kernel, reduce, mc kernel (parameter tables are cached)
    counter[0] += 1
    if world > 1 and counter[0] % EPOCH_BATCHES == 0:
      epoch_end()

  def epoch_end():
    # epoch end (training.py:619-620): the only exchange of the sharded run --
    # accumulation is linear, so the [2P + 4] sums are all-reduced once per
    # epoch, not once per batch
    payload[:2 * P].copy_(sums.sums.reshape(-1))
    payload[2 * P:].copy_(sums.stats.float())
    dist.all_reduce(payload)
    sums.reset()

  for _ in range(max(args.warmup, 3)):
    step()
    flush.zero_()
  if world > 1:
    for _ in range(3):               # one-time NCCL set-up (first collective after graph replays: ~10 ms,
      epoch_end()                    # measured) happens here, not in the timed region
      step()
      flush.zero_()
  counter[0] = 0
  # NVML start-up (milliseconds, rank 0 only) before the barrier: a rank that
  # enters the timed loop late makes the others wait in the first all-reduce
  clock = ClockSampler(local_rank) if rank == 0 else None
  launches[0] = 0
  marks = [[ev(), ev()] for _ in range(args.steps)]
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()

  # ---- timed region: K steps, device-resident inputs -----------------------
  torch.cuda.synchronize()
  wall0 = time.perf_counter()
  for k in range(args.steps):
    marks[k][0].record()
    step()
    marks[k][1].record()
    flush.zero_()                      # L2 flush, outside the per-step event pair
  torch.cuda.synchronize()
  wall = time.perf_counter() - wall0
  if world > 1:
    dist.barrier()
  clocks = clock.stop() if clock else None
  n_launch = launches[0]
  t_step = np.array([m[0].elapsed_time(m[1]) for m in marks]) * 1e-3
  if os.environ.get('CGSVMC_BENCH_DEBUG'):
    order = np.argsort(t_step)[::-1][:8]
    sys.stderr.write('rank %d: step time median %.1f us, max %.1f us; slowest steps %s\n' % (
        rank, np.median(t_step) * 1e6, t_step.max() * 1e6,
        [(int(i), round(float(t_step[i]) * 1e6, 1)) for i in order]))
  # per-phase device times (kernel shares, roofline kernel time): the same step
  # launched kernel by kernel with events between the phases, outside the timed region
  phase = [[ev() for _ in range(5)] for _ in range(20)]
  for k in range(20):
    step(phase[k])
    phase[k][4].record()
    flush.zero_()
  torch.cuda.synchronize()
  t_acc = np.array([m[0].elapsed_time(m[2]) for m in phase]) * 1e-3
  t_mc = np.array([m[2].elapsed_time(m[3]) for m in phase]) * 1e-3
  total = torch.tensor([t_step.sum()], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(total, op=dist.ReduceOp.MAX)
  total_s = float(total.item())

  # ---- e2e: the same step through host buffers -----------------------------
  # engine.HostFedBatchStep: the float32 [B, N] configuration tensor of the
  # reference (graph_builders.py:92-125) comes from pinned host memory every
  # step (copy stream, double-buffered: the upload of batch k overlaps the
  # compute of batch k - 1); pack + batch step + the device->host copy of the
  # [2, P] sums and the energy statistics are one captured graph per slot.
  # Every step's copies are inside the timed region.
  host_cfg = torch.empty(B, N_SITES, dtype=torch.float32).pin_memory()
  host_cfg.copy_(state.configs().cpu())
  fed = engine.HostFedBatchStep(state, ansatz, ham, sums, SWEEP_STEPS)

  def e2e_run(n, host_input=None):
    host_input = host_cfg if host_input is None else host_input
    for k in range(n):
      fed.submit(host_input)
      if fed.outstanding() > 1:
        fed.result()                                              # host consumes the energy of step k - 1
      if (k + 1) % EPOCH_BATCHES == 0:                            # epoch end: the [2, P] gradient sums
        if world > 1:                                             # (all-reduced over the walker shards first)
          payload[:2 * P].copy_(sums.sums.reshape(-1))
          payload[2 * P:].copy_(sums.stats.float())
          dist.all_reduce(payload)
        fed.fetch_sums()
        sums.reset()
    while fed.outstanding():
      fed.result()

  e2e_run(4)
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  e0, e1 = ev(), ev()
  e0.record()
  e2e_run(args.steps)
  e1.record()
  torch.cuda.synchronize()
  e2e_total = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
  e2e_s = float(e2e_total.item())
  # pinned host -> device rate of this box for the 1.18 MB configuration tensor
  # (explains how far e2e sits above the device-resident step)
  dst_probe = torch.empty_like(host_cfg, device=dev)
  c0, c1 = ev(), ev()
  dst_probe.copy_(host_cfg, non_blocking=True)
  torch.cuda.synchronize()
  c0.record()
  for _ in range(20):
    dst_probe.copy_(host_cfg, non_blocking=True)
  c1.record()
  torch.cuda.synchronize()
  h2d_gbps = 20 * host_cfg.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
  # the same with the host holding the walkers in the library's bit-packed layout
  # (8 B per walker instead of 144 B): shows how much of e2e is the PCIe upload
  host_packed = state.packed.cpu().pin_memory()
  e2e_run(4, host_packed)
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  p0, p1 = ev(), ev()
  p0.record()
  e2e_run(args.steps, host_packed)
  p1.record()
  torch.cuda.synchronize()
  e2e_packed_total = torch.tensor([p0.elapsed_time(p1) * 1e-3], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(e2e_packed_total, op=dist.ReduceOp.MAX)
  e2e_packed_s = float(e2e_packed_total.item())

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  # ---- bookkeeping ----------------------------------------------------------
  pk = peaks()
  cfg_now = state.configs()
  bonds_t = torch.as_tensor(np.asarray(ij), device=dev, dtype=torch.long)
  n_act = float((cfg_now[:, bonds_t[:, 0]] * cfg_now[:, bonds_t[:, 1]] < 0).sum(dim=1).float().mean().item())
  walkers_total = B * world
  value = walkers_total * SWEEP_STEPS * args.steps / total_s
  eloc_rate = walkers_total * args.steps / total_s
  # The step is one fused kernel (estimators + sweep) plus the reduction; the
  # split launches below are timed only to show how the fused time divides.
  shares = {'split launch: rbm2::walker_kernel (accumulate) + reduce': float(t_acc.mean()),
            'split launch: rbm2::mc_kernel (36 Metropolis steps)': float(t_mc.mean())}
  H, N = HIDDEN, N_SITES
  f_inc = 4 * H + 4                                            # flop per ratio
  f_fwd = 2 * N * H + 2 * N + 6 * H                            # 10,440 + lncosh arithmetic
  f_grad = 2 * 2 * (N + 1) * (H + 1)                           # two weight columns, FMA = 2 flop
  tab_bytes = 2 * H * 4                                        # two table rows per ratio
  kernel = 'rbm2::walker_kernel<MC> (cgsvmc_batch_step: E_loc + gradient sums + 36 Metropolis steps)'
  ratios = B * (n_act + SWEEP_STEPS)                           # E_loc bond flips + sampler proposals
  flop = B * (f_fwd + f_grad) + ratios * f_inc
  mufu = ratios * (H // 4 + 1) + B * 3 * H                     # lg2 per 4 units; ex2, rcp, lg2 per unit of the state build
  bytes_alg = B * (8 + 8 + 4 + 4) + 148 * 2 * P * 4            # configs in/out, E_loc, z; per-CTA partial sums
  # device time of the captured step (fused kernel + reduction + launch gap):
  # an upper bound of the fused kernel's own duration, so `frac` is conservative
  t_k = total_s / args.steps if graphed is not None else float(t_acc.mean() + t_mc.mean())
  f_hz = pk['sm_max_mhz'] * 1e6
  fp32_peak = 148 * 128 * 2 * f_hz / 1e12                      # TFLOP/s, derived
  mufu_peak = 148 * 16 * f_hz / 1e12                           # T transcendental/s, derived
  smem_peak = 148 * 128 * f_hz / 1e12                          # TB/s, derived (128 B/clk/SM)
  # shared-memory bytes the algorithm needs per launch: two table rows per
  # amplitude ratio (E_loc bond flips + sampler proposals) and one 2W row per up
  # site for the state build; measured on B200: 128 B / clk / SM
  # (profiles/r01u_fp32_pipes_microbench.txt)
  smem_alg = ratios * tab_bytes + B * (N // 2) * H * 4
  roofline = {
      'kernel': kernel,
      'bound': 'shared-memory bandwidth (walker state and ratio tables are SM-resident; neither HBM nor '
               'the tensor pipe bounds this kernel, SURVEY.md 8(d))',
      'achieved': smem_alg / t_k / 1e12, 'peak': smem_peak, 'unit': 'TB/s',
      'frac': smem_alg / t_k / 1e12 / smem_peak, 'traffic': None,
      'peak_source': 'derived: 148 SM x 128 B/clk x sm_max_mhz (%s clock; 128 B/clk/SM confirmed by '
                     'profiles/microbench/fp32_pipes.cu)' % pk['source'],
      'what': 'algorithmic shared-memory bytes: 2 x H x 4 B per amplitude ratio + N/2 x H x 4 B per state build',
      'fp32': {'achieved': flop / t_k / 1e12, 'peak': fp32_peak, 'unit': 'TFLOP/s',
               'frac': flop / t_k / 1e12 / fp32_peak,
               'peak_source': 'derived: 148 SM x 128 FP32 lanes x 2 x sm_max_mhz'},
      'mufu': {'achieved': mufu / t_k / 1e12, 'peak': mufu_peak, 'unit': 'Ttranscendental/s',
               'frac': mufu / t_k / 1e12 / mufu_peak},
      'hbm': {'achieved': bytes_alg / t_k / 1e9, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
              'frac': bytes_alg / t_k / 1e9 / pk['hbm_gbs'], 'peak_source': pk['source']},
      'kernel_ms': t_k * 1e3, 'n_active_bonds_mean': n_act,
      'kernel_ms_is': 'CUDA-event time of the whole captured step on the launching stream (fused kernel + '
                      'reduction); ncu share of the fused kernel: profiles/ launch list',
  }
  traffic_path = os.path.join(REPO, 'profiles', 'dram_traffic.json')
  if os.path.exists(traffic_path):          # dram__bytes_read + write of one ncu --set full capture
    tr = json.load(open(traffic_path))
    roofline['traffic'] = tr.get('bytes_per_launch')
    roofline['traffic_source'] = tr.get('source')
  line = {
      'metric': 'walker_steps_per_sec', 'value': value, 'unit': 'walker-steps/s',
      'eloc_evals_per_sec': eloc_rate,
      'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
      'ms_per_step': total_s / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': WORKLOAD, 'walkers_per_gpu': B, 'mc_steps_per_step': SWEEP_STEPS,
                 'n_bonds': 72, 'n_params': P,
                 'l2': 'flushed: %d MiB memset between steps, outside the per-step CUDA-event pairs' % (L2_FLUSH_BYTES >> 20),
                 'launch': ('one captured CUDA graph per step: cgsvmc_batch_step = fused estimator + sweep '
                            'kernel and the deterministic reduction (the table build is replayed only '
                            'after a parameter update)' if args.cuda_graph else 'kernel by kernel'),
                 'parallelism': 'walkers sharded, params replicated' + (
                     ', one all-reduce of [2P+4] floats per epoch of %d steps' % EPOCH_BATCHES
                     if world > 1 else '')},
      'kernel_ms': {k: v * 1e3 for k, v in shares.items()},
      'kernel_rates': {'sampler_walker_steps_per_sec': B * SWEEP_STEPS / float(t_mc.mean()),
                       'accumulate_eloc_evals_per_sec': B / float(t_acc.mean())},
      'roofline': roofline,
      'e2e': {'value': walkers_total * SWEEP_STEPS * args.steps / e2e_s, 'unit': 'walker-steps/s',
              'eloc_evals_per_sec': walkers_total * args.steps / e2e_s,
              'ms_per_step': e2e_s / args.steps * 1e3,
              'h2d_bytes_per_step': fed.h2d_bytes,
              'd2h_bytes_per_step': fed.d2h_bytes_stats + fed.d2h_bytes_sums / EPOCH_BATCHES,
              'd2h': 'energy statistics every step; the [2, P] gradient sums once per epoch of %d steps '
                     '(training.py:562-568 reads them once per epoch)' % EPOCH_BATCHES,
              'input': 'float32 [B, N] +-1 configurations (the reference layout) from pinned host memory',
              'h2d_gbps_this_box': h2d_gbps,
              'packed_host_input': {
                  'value': walkers_total * SWEEP_STEPS * args.steps / e2e_packed_s,
                  'ms_per_step': e2e_packed_s / args.steps * 1e3,
                  'h2d_bytes_per_step': int(host_packed.numel() * 8),
                  'what': 'same loop with the host holding the walkers bit-packed (uint64 [B, ceil(N/64)])'}},
      'gpu_launches': n_launch,
      'clocks': clocks,
      'wall_s_timed_region': wall,
  }
  if world == 1 and not args.no_cpu_baseline:
    c = cpu_arm(B, 3, 1)
    line['cpu_baseline'] = {
        'value': c['walker_steps_per_sec'], 'unit': 'walker-steps/s',
        'eloc_evals_per_sec': c['eloc_evals_per_sec'], 'cores': c['cores'], 'kind': 'port',
        'ms_per_step': c['ms_per_step'],
        'sample': '3 full steps (after 1 warm-up) of the same workload at the full %d walkers: '
                  'reference-equivalent torch-CPU float32 op sequence (TF1 not installable)' % B}
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def _claim_stdout():
  """Returns a file object on the real stdout and points fd 1 at stderr, so
  that library banners (NCCL prints its version on stdout) cannot get in front
  of the one JSON line the driver parses."""
  sys.stdout.flush()
  real = os.fdopen(os.dup(1), 'w')
  os.dup2(2, 1)
  return real


def main():
  global print
  out = _claim_stdout()
  _print = print
  print = lambda *a, **k: (_print(*a, **dict(k, file=out)), out.flush())
  args = parse_args()
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if args.impl == 'reference':
    run_reference(args, rank)
  else:
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
  main()

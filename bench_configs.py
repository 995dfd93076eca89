#!/usr/bin/env python
"""Device-time measurements of the other BASELINE.json configurations (C1, C3,
C4 shape, C5) on one GPU; `bench.py` is the contract benchmark (C2).  Prints
one JSON line per configuration:

  python bench_configs.py [--configs c1,c3,c5rbm,c5conv] [--reps 5] [--walker-sweep 16384,...]
  torchrun --nproc-per-node 8 ... bench_configs.py --configs c3,c5rbm     # walkers sharded over 8 GPUs

Each line reports the sampler (one sweep = N Metropolis steps per walker), the
local energy and the fused accumulate (E_loc + both gradient sums) with CUDA
events on the launching stream, inputs resident in HBM, after warm-up.
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)


def square_bonds(size, nnn=False):
  """NN bonds with (jx, jz) = (-1, 1); with nnn the J1-J2 model at J2 = 0.5."""
  from cgs_vmc_b200 import lattices
  if nnn:
    return lattices.j1j2_couplings(size, 0.5)
  return lattices.heisenberg_couplings(lattices.square_nn_bonds(size), -1.0, 1.0)


CONFIGS = {
    # name: (ansatz kwargs, lattice, walkers, flop per forward)
    'c1': dict(kind='fully_connected', n=20, kw=dict(num_layers=3, layer_size=80), chain=True,
               walkers=1024, f_fwd=28960, desc='C1: chain-20 Heisenberg, fully_connected 20-80-80-80-1, 1024 walkers'),
    'c3': dict(kind='conv_2d', n=100, kw=dict(num_layers=5, num_filters=16, kernel_size=5, size_x=10, size_y=10),
               size=10, nnn=True, walkers=8192, f_fwd=5.2e6,
               desc='C3: 10x10 J1-J2 (J2=0.5), conv_2d 5x16x5x5, 8192 walkers (one GPU share of 65536)'),
    'c4': dict(kind='conv_2d', n=36, kw=dict(num_layers=5, num_filters=16, kernel_size=5, size_x=6, size_y=6),
               size=6, nnn=True, walkers=8192, f_fwd=1.872e6,
               desc='C4 shape: 6x6 J1-J2, conv_2d trainee, 8192 walkers'),
    'c5rbm': dict(kind='rbm', n=256, kw=dict(num_layers=0, layer_size=256), size=16, nnn=False,
                  walkers=131072, f_fwd=131584,
                  desc='C5: 16x16 Heisenberg, rbm H=256, 131072 walkers (one GPU share of 1M)'),
    'c5conv': dict(kind='conv_2d', n=256, kw=dict(num_layers=5, num_filters=16, kernel_size=5, size_x=16, size_y=16),
                   size=16, nnn=False, walkers=4096, f_fwd=13.312e6,
                   desc='C5: 16x16 Heisenberg, conv_2d 5x16x5x5, 4096 walkers'),
}


def time_call(fn, reps):
  fn()
  torch.cuda.synchronize()
  times = []
  for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) * 1e-3)
  return float(np.median(times))


def _dist():
  """(rank, world) under torchrun (one process per GPU, walkers sharded), else (0, 1)."""
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if world == 1:
    return 0, 1
  import torch.distributed as dist
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if not dist.is_initialized():
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  return dist.get_rank(), world


def _max_over_ranks(t, world):
  if world == 1:
    return t
  import torch.distributed as dist
  v = torch.tensor([t], dtype=torch.float64, device='cuda')
  dist.all_reduce(v, op=dist.ReduceOp.MAX)
  return float(v.item())


def run(name, reps, walkers=None, sweep_fraction=1.0):
  """`walkers` is the PER-GPU count; under torchrun every rank takes its own
  shard (global walker ids rank * B ...), times are the max over ranks and the
  rates are whole-job."""
  from cgs_vmc_b200 import _native, engine
  rank, world = _dist()
  c = CONFIGS[name]
  n, B = c['n'], walkers or c['walkers']
  a = _native.Ansatz(c['kind'], n, **c['kw'])
  gen = torch.Generator().manual_seed(1234)
  flat = torch.randn(a.num_params, generator=gen) * (1.0 / math.sqrt(n if c['kind'] != 'conv_2d' else 25 * 16))
  a.set_params(flat)
  if c.get('chain'):
    ij = np.asarray([(i, (i + 1) % n) for i in range(n)], np.int32)
    jx, jz = np.full(n, -1.0, np.float32), np.ones(n, np.float32)
  else:
    ij, jx, jz = square_bonds(c['size'], c.get('nnn', False))
  ham = _native.Hamiltonian(ij, jx, jz, n)
  state = engine.WalkerState(B, n, seed=0xC65, walker_id0=rank * B)
  sums = engine.EnergyGradientSums(a, B)
  steps = max(1, int(round(n * sweep_fraction)))
  state.mc_steps(a, steps)                                  # warm-up / equilibrate a little
  if world > 1:
    torch.distributed.barrier()
  t_mc = _max_over_ranks(time_call(lambda: state.mc_steps(a, steps), reps), world)
  t_eloc = _max_over_ranks(time_call(lambda: a.local_energy(ham, state.packed), reps), world)
  t_acc = _max_over_ranks(time_call(lambda: sums.accumulate(ham, state.packed), reps), world)
  if world > 1:      # the exchange of the sharded run: [2P + 4] floats, once per epoch
    payload = torch.cat([sums.sums.reshape(-1), sums.stats.float()])
    t_ar = _max_over_ranks(time_call(lambda: torch.distributed.all_reduce(payload), reps), world)
  else:
    t_ar = None
  mask, _ = ham.flip_enum(state.packed, want_flipped=False)
  n_act = float(sum(bin(int(v) & 0xffffffff).count('1') for v in mask[:256].cpu().numpy().reshape(-1))) / min(B, 256)
  if rank != 0:
    return
  Bt = B * world
  line = {
      'config': c['desc'], 'n_gpus': world, 'walkers_per_gpu': B, 'walkers': Bt, 'n_sites': n,
      'n_bonds': int(len(ij)), 'n_params': a.num_params,
      'mc_steps_per_launch': steps,
      'sampler_ms': t_mc * 1e3, 'walker_steps_per_sec': Bt * steps / t_mc,
      'local_energy_ms': t_eloc * 1e3, 'eloc_evals_per_sec': Bt / t_eloc,
      'accumulate_ms': t_acc * 1e3, 'accumulate_evals_per_sec': Bt / t_acc,
      'allreduce_sums_ms': None if t_ar is None else t_ar * 1e3,
      'n_active_bonds_mean': n_act,
      'sampler_tflops_algorithmic': Bt * steps * c['f_fwd'] / t_mc / 1e12 if c['kind'] != 'rbm' else None,
      'eloc_tflops_algorithmic': Bt * (1 + n_act) * c['f_fwd'] / t_eloc / 1e12 if c['kind'] != 'rbm' else None,
  }
  print(json.dumps(line), flush=True)


def main():
  # keep stdout for the JSON lines only (NCCL prints its version banner there)
  global print
  sys.stdout.flush()
  out = os.fdopen(os.dup(1), 'w')
  os.dup2(2, 1)
  _print = print
  print = lambda *a, **k: _print(*a, **dict(k, file=out, flush=True))
  ap = argparse.ArgumentParser()
  ap.add_argument('--configs', default='c1,c3,c4,c5rbm,c5conv')
  ap.add_argument('--reps', type=int, default=5)
  ap.add_argument('--walkers', type=int, default=None, help='walkers per GPU')
  ap.add_argument('--walker-sweep', default='', help='comma-separated walkers-per-GPU counts (config C5 scaling sweep)')
  ap.add_argument('--sweep-fraction', type=float, default=1.0,
                  help='fraction of a sweep (N steps) per sampler launch')
  args = ap.parse_args()
  for name in args.configs.split(','):
    for w in ([int(x) for x in args.walker_sweep.split(',')] if args.walker_sweep else [args.walkers]):
      run(name.strip(), args.reps, w, args.sweep_fraction)
  if int(os.environ.get('WORLD_SIZE', '1')) > 1:
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
  main()

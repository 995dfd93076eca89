"""Walker sharding across the GPUs of one box.

Walkers are independent Markov chains: rank r owns the contiguous block of
global walker ids [r * B/G, (r + 1) * B/G); parameters are replicated; the only
exchange is one all-reduce(sum) of the packed float32 buffer
  [ sum_b O_b (P) | sum_b E_b O_b (P) | sum E, sum E^2, n, 0 ]
per optimisation step (or a P-vector + scalar for the supervised loss).
Works with any initialised torch.distributed backend (nccl on the GPUs, gloo
in the CPU tests of the host logic).
"""
import torch
import torch.distributed as dist


def is_initialized():
  return dist.is_available() and dist.is_initialized()


def rank():
  return dist.get_rank() if is_initialized() else 0


def world_size():
  return dist.get_world_size() if is_initialized() else 1


def shard(batch_size):
  """(local batch size, global id of the first local walker).  The global
  batch must divide evenly so every rank launches identical kernels."""
  g = world_size()
  if batch_size % g != 0:
    raise ValueError('batch_size %d is not divisible by the number of ranks %d' % (batch_size, g))
  local = batch_size // g
  return local, rank() * local


def pack_sums(sums, stats):
  """[K, P] float32 sums + [4] float64 statistics -> one float32 buffer."""
  return torch.cat([sums.reshape(-1), stats.to(torch.float32)])


def unpack_sums(payload, sums, stats):
  n = sums.numel()
  sums.copy_(payload[:n].view_as(sums))
  stats.copy_(payload[n:n + stats.numel()].to(stats.dtype))


def allreduce_sums(sums, stats):
  """In-place all-reduce(sum) of the estimator accumulators; a no-op on one
  rank."""
  if world_size() == 1:
    return
  payload = pack_sums(sums, stats)
  dist.all_reduce(payload, op=dist.ReduceOp.SUM)
  unpack_sums(payload, sums, stats)


def allreduce_(tensor, op='sum'):
  if world_size() > 1:
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM if op == 'sum' else dist.ReduceOp.MAX)
  return tensor

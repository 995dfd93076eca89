"""Walker sharding across the GPUs of one box.

Walkers are independent Markov chains: rank r owns the contiguous block of
global walker ids [r * B/G, (r + 1) * B/G); parameters are replicated; the only
exchange is one all-reduce(sum) of the packed float32 buffer
  [ sum_b O_b (P) | sum_b E_b O_b (P) | sum E, sum E^2, n, 0 ]
per optimisation step (or a P-vector + scalar for the supervised loss), carried
in float64 so the energy statistics stay exact.
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
  """[K, P] float32 sums + [4] float64 statistics -> one FLOAT64 buffer
  [K * P + 4].  The statistics (sum E, sum E^2, n) stay exact in double
  (a float32 payload loses the walker count above 2^24 and the variance long
  before); the payload is latency-bound either way (86 KB at C2)."""
  return torch.cat([sums.reshape(-1).to(torch.float64), stats.to(torch.float64)])


def unpack_sums(payload, sums, stats):
  n = sums.numel()
  sums.copy_(payload[:n].view_as(sums))
  stats.copy_(payload[n:n + stats.numel()])


def allreduce_sums(sums, stats, out_sums=None, out_stats=None):
  """All-reduce(sum) of the estimator accumulators over the walker shards
  (training.py:550-564 sees the whole batch).  The totals go to `out_sums` /
  `out_stats` when given -- the local accumulators then stay local, so further
  batches can be accumulated and reduced again without double counting -- and
  in place otherwise.  Returns (sums, stats) holding the totals; a no-op on
  one rank."""
  out_sums = sums if out_sums is None else out_sums
  out_stats = stats if out_stats is None else out_stats
  if world_size() == 1:
    if out_sums is not sums:
      out_sums.copy_(sums)
      out_stats.copy_(stats)
    return out_sums, out_stats
  payload = pack_sums(sums, stats)
  dist.all_reduce(payload, op=dist.ReduceOp.SUM)
  unpack_sums(payload, out_sums, out_stats)
  return out_sums, out_stats


_PAYLOADS = {}


def allreduce_payload(sums, stats):
  """The all-reduced FLOAT64 payload [K * P + 4] itself (cgsvmc_epoch_end reads
  it directly: no unpacking pass).  The returned buffer is reused by the next
  call with the same shapes."""
  n = sums.numel()
  key = (n, stats.numel(), sums.device)
  payload = _PAYLOADS.get(key)
  if payload is None:          # one buffer per shape: two converting copies, no allocation per epoch
    payload = _PAYLOADS[key] = torch.empty(n + stats.numel(), dtype=torch.float64, device=sums.device)
  payload[:n].copy_(sums.reshape(-1))
  payload[n:].copy_(stats)
  if world_size() > 1:
    dist.all_reduce(payload, op=dist.ReduceOp.SUM)
  return payload


def broadcast_(tensor, src=0):
  """Rank `src`'s values on every rank (parameter replicas start identical)."""
  if world_size() > 1:
    dist.broadcast(tensor, src=src)
  return tensor


def allreduce_(tensor, op='sum'):
  if world_size() > 1:
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM if op == 'sum' else dist.ReduceOp.MAX)
  return tensor

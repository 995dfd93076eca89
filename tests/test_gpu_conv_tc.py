"""-m gpu parity tests of the tensor-core (tcgen05, 3xTF32) convolutional
forward path against the float64 oracle and against the SIMT float32 path.

Stated tolerance: the 3xTF32 split keeps 22 mantissa bits per operand, so the
error of z is bounded like a float32 evaluation with a 4x larger unit
roundoff: |dz| <= 2e-5 * (sum of |terms| of the forward pass)."""
import os

import numpy as np
import pytest
import torch

from oracle import ansatz as oansatz
from oracle import bits, hamiltonian, lattices, philox

pytestmark = pytest.mark.gpu
F64 = torch.float64

TC_SHAPES = [
    oansatz.AnsatzSpec('conv_2d', 100, num_layers=5, num_filters=16, kernel_size=5, size_x=10, size_y=10),  # C3
    oansatz.AnsatzSpec('conv_2d', 36, num_layers=3, num_filters=16, kernel_size=3, size_x=6, size_y=6),
    oansatz.AnsatzSpec('conv_2d', 64, num_layers=3, num_filters=32, kernel_size=4, size_x=8, size_y=8,
                       nonlinearity='tanh'),
    oansatz.AnsatzSpec('conv_2d', 256, num_layers=4, num_filters=16, kernel_size=5, size_x=16, size_y=16),   # C5
    oansatz.AnsatzSpec('conv_2d', 24, num_layers=4, num_filters=16, kernel_size=3, size_x=4, size_y=6),
    oansatz.AnsatzSpec('conv_1d', 20, num_layers=3, num_filters=16, kernel_size=5),
    oansatz.AnsatzSpec('conv_1d', 70, num_layers=4, num_filters=16, kernel_size=6),
]


def _id(s):
  return '%s_N%d_L%d_C%d_k%d' % (s.kind, s.n_sites, s.num_layers, s.num_filters, s.kernel_size)


@pytest.fixture(scope='module')
def native():
  from cgs_vmc_b200 import _native
  _native.load()
  return _native


@pytest.fixture(autouse=True)
def _tc_on():
  os.environ['CGSVMC_CONV_TC'] = '1'
  yield
  os.environ.pop('CGSVMC_CONV_TC', None)


def _setup(spec, seed, batch, bias=0.1):
  from gpu_util import make_native
  params = oansatz.init_params(spec, seed=seed, bias_scale=bias, dtype=F64)
  cfg = bits.random_sz0_configs(spec.n_sites, batch, np.random.default_rng(seed))
  return make_native(spec, oansatz.flatten(params).numpy()), params, cfg


def _bonds(spec):
  if spec.kind == 'conv_2d':
    if spec.size_x == spec.size_y:
      return lattices.j1j2_couplings(spec.size_x, 0.5)
    return lattices.heisenberg_couplings(lattices.square_nn_bonds(spec.size_x, spec.size_y))
  return lattices.heisenberg_couplings(lattices.chain_bonds(spec.n_sites), -1.0, 1.0)


@pytest.mark.parametrize('spec', TC_SHAPES, ids=_id)
@pytest.mark.parametrize('batch', [1, 5, 301])
def test_tc_log_amp_vs_oracle_and_simt(native, spec, batch):
  from gpu_util import packed_cuda, amp_scale
  if spec.n_sites > 200:
    batch = min(batch, 70)
  a, params, cfg = _setup(spec, seed=spec.n_sites + batch, batch=batch)
  packed = packed_cuda(cfg)
  z_tc = a.log_amp(packed).cpu().numpy()
  os.environ['CGSVMC_CONV_TC'] = '0'
  z_simt = a.log_amp(packed).cpu().numpy()
  os.environ['CGSVMC_CONV_TC'] = '1'
  cfg64 = torch.from_numpy(cfg).to(F64)
  zo = oansatz.log_amp(spec, params, cfg64).numpy()
  scale = amp_scale(spec, params, cfg64) if spec.nonlinearity == 'relu' else np.abs(zo) + spec.n_sites
  assert np.all(np.abs(z_tc - zo) <= 2e-5 * scale), (np.abs(z_tc - zo).max(), scale.min())
  assert np.all(np.abs(z_tc - z_simt) <= 2e-5 * scale), np.abs(z_tc - z_simt).max()


@pytest.mark.parametrize('spec', TC_SHAPES[:3] + TC_SHAPES[5:6], ids=_id)
def test_tc_local_energy_vs_oracle(native, spec):
  from gpu_util import packed_cuda
  a, params, cfg = _setup(spec, seed=7, batch=19)
  ij, jx, jz = _bonds(spec)
  ham = native.Hamiltonian(ij, jx, jz, spec.n_sites)
  e, z, diag, off = a.local_energy(ham, packed_cuda(cfg), want_parts=True)
  cfg64 = torch.from_numpy(cfg).to(F64)
  fn = lambda c: oansatz.log_amp(spec, params, c)
  eo = hamiltonian.local_energy(cfg64, ij, jx, jz, fn).numpy()
  eabs = hamiltonian.local_energy(cfg64, ij, np.abs(jx), np.abs(jz), fn).numpy()
  # ratios exp(z' - z): |dz| <= ~1e-4 for these networks
  assert np.all(np.abs(e.cpu().numpy() - eo) <= 3e-4 * (np.abs(eabs) + 1.0)), np.abs(e.cpu().numpy() - eo).max()
  d_o, _ = hamiltonian.build(cfg64, ij, jx, jz, lambda c: torch.ones(c.shape[0], dtype=F64))
  np.testing.assert_allclose(diag.cpu().numpy(), d_o.numpy(), atol=1e-5)
  np.testing.assert_allclose((diag + off).cpu().numpy(), e.cpu().numpy(), atol=1e-5, rtol=1e-6)


@pytest.mark.parametrize('spec', [TC_SHAPES[1], TC_SHAPES[0]], ids=_id)
def test_tc_sampler_matches_oracle_philox(native, spec):
  """Every proposal identical to the Philox restatement, accept decisions
  identical away from near-ties, Sz conserved, multi-step == single steps."""
  from gpu_util import packed_cuda, unpack_np
  a, params, cfg = _setup(spec, seed=9, batch=21)
  seed, w0 = 0xC65, 500
  fn = lambda c: oansatz.log_amp(spec, params, c)
  cur = cfg.copy()
  walker_ids = np.arange(cfg.shape[0], dtype=np.uint64) + np.uint64(w0)
  for step in range(6):
    packed = packed_cuda(cur)
    count = torch.zeros(1, dtype=torch.int64, device='cuda')
    a.mc_steps(packed, 1, seed, walker_id0=w0, step0=step, accept_count=count)
    got = unpack_np(packed, spec.n_sites)
    down, up, u = philox.fast_proposal(cur, seed, walker_ids, step)
    prop = cur.copy()
    rows = np.arange(cur.shape[0])
    prop[rows, down] += 2
    prop[rows, up] -= 2
    t = torch.from_numpy
    dl = (fn(t(prop).to(F64)) - fn(t(cur).to(F64))).numpy()
    acc = np.exp(2 * dl) > u
    exp = np.where(acc[:, None], prop, cur)
    near = np.abs(np.exp(2 * dl) - u) < 2e-3 * np.exp(2 * dl)
    row_same = np.all(got == exp, axis=1)
    assert np.all(row_same | near)
    other = np.where(acc[:, None], cur, prop)
    assert np.all(row_same | np.all(got == other, axis=1))
    assert int(count.item()) == int(np.all(got == prop, axis=1).sum() - np.all(prop == cur, axis=1).sum())
    cur = got
  assert np.all(cur.sum(axis=1) == 0)
  p_all = packed_cuda(cfg)
  a.mc_steps(p_all, 6, 11, walker_id0=3, step0=0)
  p_steps = packed_cuda(cfg)
  for s in range(0, 6, 2):
    a.mc_steps(p_steps, 2, 11, walker_id0=3, step0=s)
  assert torch.equal(p_all, p_steps)


# ---------------------------------------------------------------------------
# weighted gradient sums on the tensor cores (conv_tc_grad.cu)
# ---------------------------------------------------------------------------
GRAD_SHAPES = [
    oansatz.AnsatzSpec('conv_2d', 100, num_layers=5, num_filters=16, kernel_size=5, size_x=10, size_y=10),  # C3
    oansatz.AnsatzSpec('conv_2d', 36, num_layers=3, num_filters=16, kernel_size=3, size_x=6, size_y=6),
    oansatz.AnsatzSpec('conv_2d', 36, num_layers=5, num_filters=16, kernel_size=5, size_x=6, size_y=6,
                       nonlinearity='tanh'),                                                                # C4 shape
    oansatz.AnsatzSpec('conv_2d', 64, num_layers=4, num_filters=16, kernel_size=4, size_x=8, size_y=8,
                       nonlinearity='sigmoid'),                                # even kernel: pad != k - 1 - pad
    oansatz.AnsatzSpec('conv_2d', 24, num_layers=4, num_filters=16, kernel_size=3, size_x=4, size_y=6),
    oansatz.AnsatzSpec('conv_1d', 20, num_layers=3, num_filters=16, kernel_size=5, nonlinearity='tanh'),
    oansatz.AnsatzSpec('conv_1d', 70, num_layers=4, num_filters=16, kernel_size=6),
]


def _grad_both(a, packed, w, **kw):
  """(tensor-core path, SIMT tile path) of cgsvmc_weighted_grad_sum."""
  os.environ['CGSVMC_CONV_TC_GRAD'] = '1'
  tc_out = a.weighted_grad_sum(packed, w, **kw).cpu().numpy()
  os.environ['CGSVMC_CONV_TC_GRAD'] = '0'
  try:
    simt = a.weighted_grad_sum(packed, w).cpu().numpy()
  finally:
    os.environ.pop('CGSVMC_CONV_TC_GRAD', None)
  return tc_out, simt


@pytest.mark.parametrize('spec', GRAD_SHAPES, ids=_id)
@pytest.mark.parametrize('batch', [1, 5, 301])
def test_tc_weighted_grad_sum_vs_oracle_and_simt(native, spec, batch):
  """S_k = sum_b w_kb d z_b / d params (training.py:545-548) from the tcgen05
  kernel against float64 autograd and against the float32 SIMT kernel.  Stated
  tolerance (the suite's gradient tolerance): 1e-4 of the largest entry per
  entry, 3e-5 of the norm overall; ragged batches (the last batch of a CTA is
  partial), one and two weight columns, a third column through the pair loop."""
  from gpu_util import packed_cuda
  from oracle import estimators
  a, params, cfg = _setup(spec, seed=31 + batch, batch=batch)
  packed = packed_cuda(cfg)
  rng = np.random.default_rng(5)
  w = rng.normal(size=(3, batch)).astype(np.float32)
  w[0] = 1.0
  wt = torch.from_numpy(w).cuda()
  out, simt = _grad_both(a, packed, wt)
  cfg64 = torch.from_numpy(cfg).to(F64)
  ref = estimators.weighted_grad_sum(spec, params, cfg64, torch.from_numpy(w).to(F64)).numpy()
  # relu: a pre-activation within rounding of the kink may be switched the other
  # way by the tcgen05 forward (a whole unit's contribution each): entries such
  # a unit can move get that spread added to the tolerance (gpu_util.relu_kink_band;
  # zero for smooth nonlinearities and when no unit is within 2e-5 of its kink)
  from gpu_util import relu_kink_band
  band = relu_kink_band(spec, params, cfg64, torch.from_numpy(w).to(F64))
  band = band if isinstance(band, np.ndarray) else np.zeros_like(ref)
  for k in range(3):
    scale = np.abs(ref[k]).max() + 1e-3
    err = np.abs(out[k] - ref[k])
    assert np.all(err <= 1e-4 * scale + 1e-4 + 1.5 * band[k]), (k, err.max(), scale, int(err.argmax()))
    assert np.linalg.norm(out[k] - ref[k]) <= 3e-5 * np.linalg.norm(ref[k]) + 1e-4 + 1.5 * np.linalg.norm(band[k])
    assert np.all(np.abs(out[k] - simt[k]) <= 1e-4 * scale + 1e-4 + 1.5 * band[k])
  one = a.weighted_grad_sum(packed, wt[1:2].contiguous())
  np.testing.assert_allclose(one[0].cpu().numpy(), out[1], rtol=1e-5, atol=1e-5)
  twice = a.weighted_grad_sum(packed, wt[1:2].contiguous(), out=one.clone())
  np.testing.assert_allclose(twice[0].cpu().numpy(), 2 * out[1], rtol=1e-5, atol=1e-5)


def test_tc_weighted_grad_sum_large_batches(native):
  """Many configurations per CTA batch (the plan grows G with the walker
  count), deterministic from run to run, linear in the weights."""
  from gpu_util import packed_cuda
  from oracle import estimators
  spec = GRAD_SHAPES[1]
  batch = 2100
  a, params, cfg = _setup(spec, seed=77, batch=batch)
  packed = packed_cuda(cfg)
  rng = np.random.default_rng(8)
  w = rng.normal(size=(2, batch)).astype(np.float32)
  wt = torch.from_numpy(w).cuda()
  out = a.weighted_grad_sum(packed, wt)
  again = a.weighted_grad_sum(packed, wt)
  assert torch.equal(out, again)
  cfg64 = torch.from_numpy(cfg).to(F64)
  ref = estimators.weighted_grad_sum(spec, params, cfg64, torch.from_numpy(w).to(F64)).numpy()
  for k in range(2):
    scale = np.abs(ref[k]).max() + 1e-3
    assert np.abs(out[k].cpu().numpy() - ref[k]).max() <= 1e-4 * scale + 1e-4
  both = a.weighted_grad_sum(packed, (wt[0] + 2 * wt[1]).reshape(1, -1).contiguous())[0]
  scale = float(both.abs().max()) + 1e-3
  assert float((both - (out[0] + 2 * out[1])).abs().max()) <= 1e-4 * scale


def test_tc_accumulate_c3_full_size(native):
  """BASELINE C3 at 8192 walkers: accumulate (K3 + K4 + K5) with the
  tensor-core gradient equals the SIMT gradient path within the gradient
  tolerance (size-independent check; the oracle needs minutes at this size)."""
  from gpu_util import packed_cuda
  spec = GRAD_SHAPES[0]
  batch = 8192
  a, params, cfg = _setup(spec, seed=3, batch=batch)
  packed = packed_cuda(cfg)
  ij, jx, jz = _bonds(spec)
  ham = native.Hamiltonian(ij, jx, jz, spec.n_sites)
  res = []
  for flag in ('1', '0'):
    os.environ['CGSVMC_CONV_TC_GRAD'] = flag
    sums = torch.zeros(2, a.num_params, dtype=torch.float32, device='cuda')
    stats = torch.zeros(4, dtype=torch.float64, device='cuda')
    a.accumulate(ham, packed, sums, stats)
    res.append((sums.cpu().numpy(), stats.cpu().numpy()))
  os.environ.pop('CGSVMC_CONV_TC_GRAD', None)
  np.testing.assert_array_equal(res[0][1], res[1][1])
  for k in range(2):
    scale = np.abs(res[1][0][k]).max()
    assert np.abs(res[0][0][k] - res[1][0][k]).max() <= 1e-4 * scale

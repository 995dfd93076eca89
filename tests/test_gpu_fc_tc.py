"""-m gpu parity tests of the tensor-core (tcgen05) fully-connected forward
path (fc_tc.cu) against the float64 oracle, the golden vectors recorded from
the reference and the SIMT float32 path of net.cu (CGSVMC_FC_TC=0).

Stated tolerance: three-way fp16 split of activations and weights (33 mantissa
bits, exact float32 products) with the tensor core's fp32 accumulation:
|dz| <= 2e-5 * (sum of |terms| of the forward pass), the bound the SIMT path
is held to as well."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ansatz as oansatz
from oracle import bits, hamiltonian, lattices, philox

pytestmark = pytest.mark.gpu
F64 = torch.float64

FC_SHAPES = [
    oansatz.AnsatzSpec('fully_connected', 20, num_layers=3, layer_size=80),          # C1
    oansatz.AnsatzSpec('fully_connected', 36, num_layers=1, layer_size=64),
    oansatz.AnsatzSpec('fully_connected', 70, num_layers=2, layer_size=32, nonlinearity='tanh'),
    oansatz.AnsatzSpec('fully_connected', 100, num_layers=4, layer_size=48),
    oansatz.AnsatzSpec('fully_connected', 12, num_layers=2, layer_size=16, nonlinearity='sigmoid'),
    oansatz.AnsatzSpec('fully_connected', 256, num_layers=2, layer_size=80),
]


def _id(s):
  return 'N%d_L%d_H%d_%s' % (s.n_sites, s.num_layers, s.layer_size, s.nonlinearity)


@pytest.fixture(scope='module')
def native():
  from cgs_vmc_b200 import _native
  _native.load()
  return _native


@pytest.fixture(autouse=True)
def _tc_on():
  os.environ['CGSVMC_FC_TC'] = '1'
  os.environ['CGSVMC_FC_WARP'] = '0'      # small batches too: these tests are about the tensor-core kernels
  yield
  os.environ.pop('CGSVMC_FC_TC', None)
  os.environ.pop('CGSVMC_FC_WARP', None)


def _setup(spec, seed, batch, bias=0.1):
  from gpu_util import make_native
  params = oansatz.init_params(spec, seed=seed, bias_scale=bias, dtype=F64)
  cfg = bits.random_sz0_configs(spec.n_sites, batch, np.random.default_rng(seed))
  return make_native(spec, oansatz.flatten(params).numpy()), params, cfg


@pytest.mark.parametrize('spec', FC_SHAPES, ids=_id)
@pytest.mark.parametrize('batch', [1, 127, 129, 700, 40000])
def test_fc_tc_log_amp_vs_oracle_and_simt(native, spec, batch):
  """Ragged batches around the 128-row tile and the two-tile pipeline
  (40000 > 128 x 148: two tiles in flight per CTA)."""
  from gpu_util import packed_cuda, amp_scale
  a, params, cfg = _setup(spec, seed=spec.n_sites + batch, batch=batch)
  packed = packed_cuda(cfg)
  z_tc = a.log_amp(packed).cpu().numpy()
  os.environ['CGSVMC_FC_TC'] = '0'
  z_simt = a.log_amp(packed).cpu().numpy()
  os.environ['CGSVMC_FC_TC'] = '1'
  sub = slice(0, min(batch, 2000))
  cfg64 = torch.from_numpy(cfg[sub]).to(F64)
  zo = oansatz.log_amp(spec, params, cfg64).numpy()
  scale = amp_scale(spec, params, cfg64) if spec.nonlinearity == 'relu' else np.abs(zo) + spec.n_sites
  assert np.all(np.abs(z_tc[sub] - zo) <= 2e-5 * scale), (np.abs(z_tc[sub] - zo).max(), scale.min())
  assert np.all(np.isfinite(z_tc))
  assert np.abs(z_tc - z_simt).max() <= 2e-5 * (np.abs(z_simt).max() + spec.n_sites)


def test_fc_tc_golden_amplitudes(native):
  """psi of the reference's own FullyConnectedNetwork on the C1 shape."""
  from gpu_util import make_native, packed_cuda
  spec, g = load_golden('fc_chain20')
  a = make_native(spec, g['params_flat'])
  z = a.log_amp(packed_cuda(g['configs'])).double().cpu().numpy()
  np.testing.assert_allclose(np.exp(z - float(g['shift'])), g['psi'], rtol=3e-5)


@pytest.mark.parametrize('spec', FC_SHAPES[:4], ids=_id)
@pytest.mark.parametrize('batch', [19, 1024])
def test_fc_tc_local_energy_vs_oracle(native, spec, batch):
  from gpu_util import packed_cuda
  a, params, cfg = _setup(spec, seed=7, batch=batch)
  n = spec.n_sites
  if n == 36:
    ij, jx, jz = lattices.j1j2_couplings(6, 0.5)
  elif n == 100:
    ij, jx, jz = lattices.j1j2_couplings(10, 0.5)
  else:
    ij, jx, jz = lattices.heisenberg_couplings(lattices.chain_bonds(n), -1.0, 1.0)
  ham = native.Hamiltonian(ij, jx, jz, n)
  e, z, diag, off = a.local_energy(ham, packed_cuda(cfg), want_parts=True)
  os.environ['CGSVMC_FC_TC'] = '0'
  e_simt, z_simt = a.local_energy(ham, packed_cuda(cfg))
  os.environ['CGSVMC_FC_TC'] = '1'
  sub = slice(0, min(batch, 64))
  cfg64 = torch.from_numpy(cfg[sub]).to(F64)
  fn = lambda c: oansatz.log_amp(spec, params, c)
  eo = hamiltonian.local_energy(cfg64, ij, jx, jz, fn).numpy()
  eabs = hamiltonian.local_energy(cfg64, ij, np.abs(jx), np.abs(jz), fn).numpy()
  got = e.cpu().numpy()
  assert np.all(np.abs(got[sub] - eo) <= 3e-4 * (np.abs(eabs) + 1.0)), np.abs(got[sub] - eo).max()
  assert np.abs(got - e_simt.cpu().numpy()).max() <= 3e-4 * (np.abs(e_simt.cpu().numpy()).max() + 1.0)
  np.testing.assert_allclose(z.cpu().numpy(), z_simt.cpu().numpy(), atol=2e-5 * (np.abs(z_simt.cpu().numpy()).max() + n))
  d_o, _ = hamiltonian.build(cfg64, ij, jx, jz, lambda c: torch.ones(c.shape[0], dtype=F64))
  np.testing.assert_allclose(diag.cpu().numpy()[sub], d_o.numpy(), atol=1e-5)
  np.testing.assert_allclose((diag + off).cpu().numpy(), got, atol=1e-5, rtol=1e-6)


def test_fc_tc_golden_local_energy(native):
  from gpu_util import make_native, packed_cuda
  spec, g = load_golden('fc_chain20')
  a = make_native(spec, g['params_flat'])
  ham = native.Hamiltonian(g['bonds_ij'], g['bonds_jx'], g['bonds_jz'], spec.n_sites)
  e, _ = a.local_energy(ham, packed_cuda(g['configs']))
  np.testing.assert_allclose(e.cpu().numpy(), g['local_energy'], rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize('spec', [FC_SHAPES[0], FC_SHAPES[2]], ids=_id)
def test_fc_tc_sampler_matches_oracle_philox(native, spec):
  """Every proposal identical to the Philox restatement, accept decisions
  identical away from near-ties, Sz conserved, multi-step == single steps,
  independent of how the walkers are grouped into CTAs."""
  from gpu_util import packed_cuda, unpack_np
  a, params, cfg = _setup(spec, seed=9, batch=300)
  seed, w0 = 0xC65, 500
  fn = lambda c: oansatz.log_amp(spec, params, c)
  cur = cfg.copy()
  walker_ids = np.arange(cfg.shape[0], dtype=np.uint64) + np.uint64(w0)
  for step in range(5):
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
  z_all = torch.empty(cfg.shape[0], device='cuda')
  a.mc_steps(p_all, 6, 11, walker_id0=3, step0=0, log_amp_out=z_all)
  p_steps = packed_cuda(cfg)
  for s in range(0, 6, 2):
    a.mc_steps(p_steps, 2, 11, walker_id0=3, step0=s)
  assert torch.equal(p_all, p_steps)
  torch.testing.assert_close(z_all, a.log_amp(p_all), rtol=0, atol=1e-4)


def test_fc_tc_sampler_large_batch_two_tiles(native):
  """More than 128 x 148 walkers: two tiles in flight per CTA; trajectories
  equal those of the same walkers sampled in small groups."""
  from gpu_util import packed_cuda
  spec = FC_SHAPES[0]
  a, params, cfg = _setup(spec, seed=3, batch=20000)
  p_big = packed_cuda(cfg)
  a.mc_steps(p_big, 4, 77, walker_id0=0, step0=0)
  p_small = packed_cuda(cfg[:500])
  a.mc_steps(p_small, 4, 77, walker_id0=0, step0=0)
  same = (p_big[:500] == p_small).all(dim=1)
  assert same.float().mean().item() > 0.99      # near-ties may differ between tile placements? no: same arithmetic
  assert torch.equal(p_big[:500], p_small)


# ---------------------------------------------------------------------------
# gradient on the tensor cores (fc_tc_grad.cu)
# ---------------------------------------------------------------------------
# 20000 walkers = two tiles per CTA.  Smooth nonlinearities there: with relu one
# pre-activation out of millions lands within float32 rounding of the kink and
# its derivative (0 or 1) legitimately differs between two float32 evaluations.
GRAD_CASES = [(s, b) for b in (1, 127, 300) for s in FC_SHAPES[:5]] + [
    (oansatz.AnsatzSpec('fully_connected', 20, num_layers=3, layer_size=80, nonlinearity='tanh'), 20000),
    (oansatz.AnsatzSpec('fully_connected', 100, num_layers=4, layer_size=48, nonlinearity='tanh'), 20000),
    (FC_SHAPES[2], 20000), (FC_SHAPES[4], 20000)]


@pytest.mark.parametrize('spec,batch', GRAD_CASES, ids=lambda v: _id(v) if hasattr(v, 'kind') else str(v))
def test_fc_tc_weighted_grad_sum_vs_oracle_and_simt(native, spec, batch):
  """S_k = sum_b w_kb O_b (training.py:545-548, 169-175) from the tcgen05
  forward / backward / weight-gradient GEMMs against float64 autograd and
  against the SIMT kernel of net.cu; gradient tolerance of the suite: 1e-4 of
  the largest entry per entry, 3e-5 in norm."""
  from oracle import estimators
  from gpu_util import packed_cuda
  a, params, cfg = _setup(spec, seed=17 + batch, batch=batch)
  rng = np.random.default_rng(3)
  w = rng.normal(size=(2, batch)).astype(np.float32)
  w[0] = 1.0
  packed = packed_cuda(cfg)
  wt = torch.from_numpy(w).cuda()
  out = a.weighted_grad_sum(packed, wt).cpu().numpy()
  os.environ['CGSVMC_FC_TC_GRAD'] = '0'
  try:
    simt = a.weighted_grad_sum(packed, wt).cpu().numpy()
  finally:
    os.environ.pop('CGSVMC_FC_TC_GRAD', None)
  cfg64 = torch.from_numpy(cfg).to(F64)
  ref = estimators.weighted_grad_sum(spec, params, cfg64, torch.from_numpy(w).to(F64)).numpy()
  for k in range(2):
    scale = np.abs(ref[k]).max() + 1e-3
    err = np.abs(out[k] - ref[k]).max()
    err_simt = np.abs(simt[k] - ref[k]).max()
    # float32 sums over the batch: hold the tensor-core path to the suite's
    # tolerance, or -- for the 20000-walker sums, where float32 summation noise
    # dominates -- to the error of the SIMT kernel
    assert err <= max(1e-4 * scale + 1e-4, 2.0 * err_simt), (k, err, err_simt, scale)
    if batch <= 300:
      assert np.linalg.norm(out[k] - ref[k]) <= 3e-5 * np.linalg.norm(ref[k]) + 1e-4
  # single column and accumulate-into semantics
  one = a.weighted_grad_sum(packed, torch.from_numpy(w[1:2].copy()).cuda())
  np.testing.assert_allclose(one[0].cpu().numpy(), out[1], rtol=1e-5, atol=1e-5 * (np.abs(out[1]).max() + 1))
  twice = a.weighted_grad_sum(packed, torch.from_numpy(w[1:2].copy()).cuda(), out=one.clone())
  np.testing.assert_allclose(twice[0].cpu().numpy(), 2 * out[1], rtol=1e-5, atol=1e-5 * (np.abs(out[1]).max() + 1))
  # deterministic: fixed summation order everywhere
  assert np.array_equal(a.weighted_grad_sum(packed, wt).cpu().numpy(), out)


def test_fc_tc_energy_gradient_golden(native):
  """The energy gradient the reference's EnergyGradientOptimizer graph produced
  for the C1 network (training.py:539-564), through fc_tc local energy + fc_tc
  gradient."""
  from gpu_util import make_native, packed_cuda
  spec, g = load_golden('fc_chain20')
  a = make_native(spec, g['params_flat'])
  ham = native.Hamiltonian(g['bonds_ij'], g['bonds_jx'], g['bonds_jz'], spec.n_sites)
  packed = packed_cuda(g['eg_configs'])
  e, _ = a.local_energy(ham, packed)
  s = a.weighted_grad_sum(packed, torch.stack([torch.ones_like(e), e]))
  stats = native.energy_stats(e).cpu().numpy()
  mean_e = stats[0] / stats[2]
  assert abs(mean_e - float(g['eg_mean_energy'])) < 5e-5 * (1 + abs(mean_e))
  grad = (s[1] - mean_e * s[0]).cpu().numpy()
  ref = g['eg_gradient']
  assert np.linalg.norm(grad - ref) <= 5e-4 * np.linalg.norm(ref) + 1e-4


# ---------------------------------------------------------------------------
# warp-per-walker sampler for small batches (fc_warp.cu)
# ---------------------------------------------------------------------------
FC_WARP_SHAPES = FC_SHAPES + [
    oansatz.AnsatzSpec('fully_connected', 130, num_layers=1, layer_size=16, nonlinearity='tanh'),   # three words
    oansatz.AnsatzSpec('fully_connected', 16, num_layers=8, layer_size=96, nonlinearity='tanh'),    # widest, deepest
]


@pytest.mark.parametrize('spec', FC_WARP_SHAPES, ids=_id)
def test_fc_warp_sampler_matches_oracle_philox(native, spec):
  """cgsvmc_mc_steps on the warp-per-walker kernel (what batches of at most
  2,048 walkers of a fully connected ansatz get): every proposal identical to
  the Philox restatement of graph_builders.py:54-89, accept decisions identical
  away from near-ties (float32 forward), acceptance count, Sz conserved,
  multi-step == single steps, and the same trajectories as the tensor-core /
  tile kernels up to near-ties."""
  from gpu_util import packed_cuda, unpack_np
  os.environ['CGSVMC_FC_WARP'] = '1'
  batch = 77
  a, params, cfg = _setup(spec, seed=9, batch=batch)
  seed, w0 = 0xC65, 500
  fn = lambda c: oansatz.log_amp(spec, params, c)
  cur = cfg.copy()
  walker_ids = np.arange(batch, dtype=np.uint64) + np.uint64(w0)
  mismatches = 0
  for step in range(8):
    packed = packed_cuda(cur)
    count = torch.zeros(1, dtype=torch.int64, device='cuda')
    a.mc_steps(packed, 1, seed, walker_id0=w0, step0=step, accept_count=count)
    got = unpack_np(packed, spec.n_sites)
    down, up, u = philox.fast_proposal(cur, seed, walker_ids, step)
    prop = cur.copy()
    rows = np.arange(batch)
    prop[rows, down] += 2
    prop[rows, up] -= 2
    t = torch.from_numpy
    dl = (fn(t(prop).to(F64)) - fn(t(cur).to(F64))).numpy()
    acc = np.exp(2 * dl) > u
    exp = np.where(acc[:, None], prop, cur)
    near = np.abs(np.exp(2 * dl) - u) < 1e-3 * np.exp(2 * dl)
    row_same = np.all(got == exp, axis=1)
    assert np.all(row_same | near)
    other = np.where(acc[:, None], cur, prop)
    assert np.all(row_same | np.all(got == other, axis=1))
    assert int(count.item()) == int(np.all(got == prop, axis=1).sum())
    mismatches += int((~row_same).sum())
    cur = got
  assert mismatches <= 2
  assert np.all(cur.sum(axis=1) == 0)
  p_all = packed_cuda(cfg)
  z_all = torch.empty(batch, device='cuda')
  a.mc_steps(p_all, 6, 11, walker_id0=3, step0=0, log_amp_out=z_all)
  p_steps = packed_cuda(cfg)
  for s in range(0, 6, 2):
    a.mc_steps(p_steps, 2, 11, walker_id0=3, step0=s)
  assert torch.equal(p_all, p_steps)
  # two shards of the walkers: same trajectories (Philox keyed by the global walker id)
  p_a, p_b = packed_cuda(cfg[:30]), packed_cuda(cfg[30:])
  a.mc_steps(p_a, 6, 11, walker_id0=3, step0=0)
  a.mc_steps(p_b, 6, 11, walker_id0=33, step0=0)
  assert torch.equal(torch.cat([p_a, p_b]), p_all)
  os.environ['CGSVMC_FC_WARP'] = '0'
  torch.testing.assert_close(z_all, a.log_amp(p_all), rtol=0, atol=2e-4)
  p_other = packed_cuda(cfg)
  a.mc_steps(p_other, 6, 11, walker_id0=3, step0=0)
  assert (p_other == p_all).all(dim=1).float().mean().item() >= 0.97

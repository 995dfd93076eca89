"""-m gpu parity tests of the pure-RBM CUDA path (through the C-ABI) against
the oracle and the golden vectors.  Tolerances are stated per test."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ansatz as oansatz
from oracle import bits, estimators, hamiltonian, lattices, philox, sampler

pytestmark = pytest.mark.gpu

F64 = torch.float64


@pytest.fixture(scope='module')
def native():
  from cgs_vmc_b200 import _native
  _native.load()
  return _native


def _c2_spec():
  return oansatz.AnsatzSpec('rbm', 36, num_layers=0, layer_size=144, size_x=6, size_y=6)


def _setup(spec, seed, batch, bias=0.1, scale=1.0):
  from gpu_util import make_native
  params = oansatz.init_params(spec, seed=seed, bias_scale=bias, dtype=F64)
  params = [p * scale for p in params]
  flat = oansatz.flatten(params).numpy()
  rng = np.random.default_rng(seed)
  cfg = bits.random_sz0_configs(spec.n_sites, batch, rng)
  return make_native(spec, flat), params, cfg


RBM_SHAPES = [
    oansatz.AnsatzSpec('rbm', 36, num_layers=0, layer_size=144),    # C2
    oansatz.AnsatzSpec('rbm', 20, num_layers=0, layer_size=40),
    oansatz.AnsatzSpec('rbm', 100, num_layers=0, layer_size=100),   # 2 words
    oansatz.AnsatzSpec('rbm', 256, num_layers=0, layer_size=256),   # C5, W in global
    oansatz.AnsatzSpec('rbm', 16, num_layers=0, layer_size=7),      # ragged H
    oansatz.AnsatzSpec('rbm', 70, num_layers=0, layer_size=150),
    oansatz.AnsatzSpec('rbm', 150, num_layers=0, layer_size=64),    # 3 words per walker
    oansatz.AnsatzSpec('rbm', 40, num_layers=0, layer_size=180),    # 16 lanes per walker
    oansatz.AnsatzSpec('rbm', 12, num_layers=0, layer_size=32),
]


def test_pack_unpack_roundtrip(native):
  rng = np.random.default_rng(0)
  for n in (2, 20, 36, 64, 65, 100, 128, 256):
    cfg = rng.choice([-1.0, 1.0], size=(37, n)).astype(np.float32)
    packed = native.pack_configs(torch.from_numpy(cfg).cuda())
    assert np.array_equal(packed.cpu().numpy().view(np.uint64), bits.pack(cfg))  # bit-exact
    back = native.unpack_configs(packed, n)
    assert np.array_equal(back.cpu().numpy(), cfg)
  empty = native.pack_configs(torch.empty(0, 36, dtype=torch.float32, device='cuda'))
  assert empty.shape == (0, 1)


def test_random_configs_sz0_and_uniform(native):
  for n in (8, 36, 100, 256):
    packed = native.random_configs(4096, n, seed=11)
    cfg = bits.unpack(packed.cpu().numpy().view(np.uint64), n)
    assert np.all(cfg.sum(axis=1) == (n - 2 * (n // 2)))
    # unused high bits are zero
    again = bits.pack(cfg)
    assert np.array_equal(again, packed.cpu().numpy().view(np.uint64))
  packed = native.random_configs(70 * 400, 8, seed=5).cpu().numpy().view(np.uint64)[:, 0]
  _, counts = np.unique(packed, return_counts=True)
  assert len(counts) == 70
  chi2 = ((counts - 400.0) ** 2 / 400.0).sum()
  assert chi2 < 130.0          # 69 dof
  # different walker ids give different configs; same ids reproduce
  a = native.random_configs(64, 36, seed=3, walker_id0=0)
  b = native.random_configs(32, 36, seed=3, walker_id0=32)
  assert torch.equal(a[32:], b)


@pytest.mark.parametrize('name', ['rbm_6x6', 'rbm_4x4_j1j2', 'fc_chain20', 'conv2d_10x10'])
def test_flip_enum_bit_exact(native, name):
  """operators.py:154-167: bit-exact against the reference-recorded flips."""
  from gpu_util import packed_cuda
  spec, g = load_golden(name)
  ham = native.Hamiltonian(g['bonds_ij'], g['bonds_jx'], g['bonds_jz'], spec.n_sites)
  packed = packed_cuda(g['configs'])
  mask, flipped = ham.flip_enum(packed)
  o_mask, o_flipped = bits.flip_enum(bits.pack(g['configs']), g['bonds_ij'], spec.n_sites)
  assert np.array_equal(mask.cpu().numpy().view(np.uint32), o_mask)
  assert np.array_equal(flipped.cpu().numpy().view(np.uint64), o_flipped)
  nb = len(g['bonds_ij'])
  got = flipped.cpu().numpy().view(np.uint64)
  for k in range(0, nb, max(1, nb // 16)):
    assert np.array_equal(bits.unpack(got[:, k], spec.n_sites), g['flipped_configs'][:, k])


def test_flip_enum_large_and_ragged(native):
  rng = np.random.default_rng(2)
  for size, b in ((16, 257), (10, 1000)):
    n = size * size
    ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(size))
    cfg = bits.random_sz0_configs(n, b, rng)
    from gpu_util import packed_cuda
    ham = native.Hamiltonian(ij, jx, jz, n)
    mask, flipped = ham.flip_enum(packed_cuda(cfg))
    o_mask, o_flipped = bits.flip_enum(bits.pack(cfg), ij, n)
    assert np.array_equal(mask.cpu().numpy().view(np.uint32), o_mask)
    assert np.array_equal(flipped.cpu().numpy().view(np.uint64), o_flipped)


@pytest.mark.parametrize('spec', RBM_SHAPES, ids=lambda s: 'N%d_H%d' % (s.n_sites, s.layer_size))
def test_log_amp_vs_oracle(native, spec):
  """z within float32 rounding of the float64 oracle:
  |dz| <= 4e-6 * (sum of |terms| of z)."""
  from gpu_util import packed_cuda, amp_scale
  a, params, cfg = _setup(spec, seed=spec.n_sites + spec.layer_size, batch=203)
  z = a.log_amp(packed_cuda(cfg)).cpu().numpy()
  zo = oansatz.log_amp(spec, params, torch.from_numpy(cfg).to(F64))
  tol = 4e-6 * amp_scale(spec, params, torch.from_numpy(cfg).to(F64))
  assert np.all(np.abs(z - zo.numpy()) <= tol), np.abs(z - zo.numpy()).max()


def test_log_amp_golden(native):
  from gpu_util import make_native, packed_cuda
  for name in ('rbm_6x6', 'rbm_4x4_j1j2'):
    spec, g = load_golden(name)
    a = make_native(spec, g['params_flat'])
    z = a.log_amp(packed_cuda(g['configs'])).cpu().numpy().astype(np.float64)
    psi = np.exp(z - float(g['shift']))
    np.testing.assert_allclose(psi, g['psi'], rtol=3e-5)   # both sides float32


@pytest.mark.parametrize('name', ['rbm_6x6', 'rbm_4x4_j1j2'])
def test_replay_step_golden(native, name):
  """graph_builders.py:59-88 with the reference's recorded uniforms: proposal
  sites exact; accept mask / new configs exact (no near-ties in the fixture)."""
  from gpu_util import make_native, packed_cuda, unpack_np
  spec, g = load_golden(name)
  a = make_native(spec, g['params_flat'])
  params = oansatz.unflatten(spec, torch.from_numpy(g['params_flat']).to(F64))
  for s in range(g['mc_before'].shape[0]):
    packed = packed_cuda(g['mc_before'][s])
    down, up, log_ratio, accept = a.mc_step_replay(
        packed, torch.from_numpy(g['mc_u_sites'][s]).cuda(),
        torch.from_numpy(g['mc_u_acc'][s]).cuda())
    before = torch.from_numpy(g['mc_before'][s]).to(F64)
    _, o_acc, o_lr, o_down, o_up = sampler.mc_step(
        before, torch.from_numpy(g['mc_u_sites'][s]).to(F64),
        torch.from_numpy(g['mc_u_acc'][s]).to(F64),
        lambda c: oansatz.log_amp(spec, params, c))
    assert np.array_equal(down.cpu().numpy(), o_down.numpy())
    assert np.array_equal(up.cpu().numpy(), o_up.numpy())
    np.testing.assert_allclose(log_ratio.cpu().numpy(), o_lr.numpy(), atol=2e-5, rtol=1e-5)
    assert np.array_equal(accept.cpu().numpy().astype(bool), o_acc.numpy())
    assert np.array_equal(unpack_np(packed, spec.n_sites), g['mc_after'][s])
    assert int(accept.sum()) == int(g['mc_accept_count'][s])


@pytest.mark.parametrize('spec', RBM_SHAPES[:4], ids=lambda s: 'N%d_H%d' % (s.n_sites, s.layer_size))
def test_replay_step_vs_oracle(native, spec):
  from gpu_util import packed_cuda, unpack_np
  a, params, cfg = _setup(spec, seed=7, batch=301)
  rng = np.random.default_rng(1)
  u_sites = rng.random((cfg.shape[0], spec.n_sites)).astype(np.float32)
  u_acc = rng.random(cfg.shape[0]).astype(np.float32)
  packed = packed_cuda(cfg)
  down, up, log_ratio, accept = a.mc_step_replay(
      packed, torch.from_numpy(u_sites).cuda(), torch.from_numpy(u_acc).cuda())
  new, o_acc, o_lr, o_down, o_up = sampler.mc_step(
      torch.from_numpy(cfg).to(F64), torch.from_numpy(u_sites).to(F64),
      torch.from_numpy(u_acc).to(F64), lambda c: oansatz.log_amp(spec, params, c))
  assert np.array_equal(down.cpu().numpy(), o_down.numpy())    # exact
  assert np.array_equal(up.cpu().numpy(), o_up.numpy())        # exact
  lr = log_ratio.cpu().numpy()
  assert np.all(np.abs(lr - o_lr.numpy()) <= 2e-5 + 1e-5 * np.abs(o_lr.numpy()))
  # accept mask exact except where |ratio - sqrt(u)| is within rounding
  near = np.abs(np.exp(o_lr.numpy()) - np.sqrt(u_acc)) < 1e-4 * np.exp(o_lr.numpy())
  same = accept.cpu().numpy().astype(bool) == o_acc.numpy()
  assert np.all(same | near)
  got = unpack_np(packed, spec.n_sites)
  exp = new.numpy().astype(np.float32)
  assert np.array_equal(got[same], exp[same])


@pytest.mark.parametrize('spec', RBM_SHAPES[:4], ids=lambda s: 'N%d_H%d' % (s.n_sites, s.layer_size))
def test_fast_sampler_matches_oracle_philox(native, spec):
  """cgsvmc_mc_steps one step at a time against the numpy restatement of the
  Philox proposal rule: every proposal and (away from near-ties) every accept
  decision identical; also exercises walker_id0 / step0 offsets."""
  from gpu_util import packed_cuda, unpack_np
  a, params, cfg = _setup(spec, seed=9, batch=64)
  seed, w0 = 0xC65, 1000
  fn = lambda c: oansatz.log_amp(spec, params, c)
  cur = cfg.copy()
  walker_ids = np.arange(cfg.shape[0], dtype=np.uint64) + np.uint64(w0)
  mismatches = 0
  for step in range(40):
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
    near = np.abs(np.exp(2 * dl) - u) < 2e-4 * np.exp(2 * dl)
    row_same = np.all(got == exp, axis=1)
    assert np.all(row_same | near)
    # a row that differs must be the un-/accepted version of the same proposal
    other = np.where(acc[:, None], cur, prop)
    assert np.all(row_same | np.all(got == other, axis=1))
    mismatches += int((~row_same).sum())
    assert int(count.item()) == int(np.all(got == prop, axis=1).sum())
    cur = got
  assert mismatches <= 2
  assert np.all(cur.sum(axis=1) == 0)


def test_multi_step_launch_equals_single_steps(native):
  """n_steps in one launch == the same steps launched one by one (Philox is
  keyed by (walker, step)), independent of how walkers are split into shards:
  bit-identical trajectories."""
  from gpu_util import packed_cuda
  spec = _c2_spec()
  a, params, cfg = _setup(spec, seed=21, batch=512)
  p_all = packed_cuda(cfg)
  a.mc_steps(p_all, 72, 5, walker_id0=0, step0=0)
  p_steps = packed_cuda(cfg)
  for s in range(0, 72, 9):
    a.mc_steps(p_steps, 9, 5, walker_id0=0, step0=s)
  assert torch.equal(p_all, p_steps)
  shards = []
  for lo in range(0, 512, 128):
    p = packed_cuda(cfg[lo:lo + 128])
    a.mc_steps(p, 72, 5, walker_id0=lo, step0=0)
    shards.append(p)
  assert torch.equal(p_all, torch.cat(shards))


def test_sampler_cached_log_amp_drift(native):
  """After 3600 incremental theta updates the cached log-amplitude equals a
  fresh forward pass: |dz| <= 2e-4 (float32 accumulation of 3600 rank-2
  updates), C2 shape at full batch."""
  spec = _c2_spec()
  a, params, _ = _setup(spec, seed=4, batch=1)
  packed = native.random_configs(8192, 36, seed=1)
  z_cached = torch.empty(8192, dtype=torch.float32, device='cuda')
  count = torch.zeros(1, dtype=torch.int64, device='cuda')
  a.mc_steps(packed, 3600, 77, accept_count=count, log_amp_out=z_cached)
  z_fresh = a.log_amp(packed)
  assert float((z_cached - z_fresh).abs().max()) <= 2e-4
  cfg = bits.unpack(packed.cpu().numpy().view(np.uint64), 36)
  assert np.all(cfg.sum(axis=1) == 0)                 # Sz conserved
  assert 0 < int(count.item()) < 8192 * 3600
  # n_steps = 0 is the identity
  before = packed.clone()
  a.mc_steps(packed, 0, 77)
  assert torch.equal(before, packed)


def test_constant_amplitude_uniform_sampling(native):
  """All-zero parameters => psi constant => every move accepted and the chain
  is uniform on the 70 Sz=0 states of N=8 (chi^2, 69 dof)."""
  from gpu_util import make_native
  spec = oansatz.AnsatzSpec('rbm', 8, num_layers=0, layer_size=4)
  a = make_native(spec, np.zeros(oansatz.num_params(spec), dtype=np.float32))
  packed = native.random_configs(28000, 8, seed=2)
  count = torch.zeros(1, dtype=torch.int64, device='cuda')
  a.mc_steps(packed, 64, 123, accept_count=count)
  assert int(count.item()) == 28000 * 64
  _, counts = np.unique(packed.cpu().numpy().view(np.uint64)[:, 0], return_counts=True)
  assert len(counts) == 70
  chi2 = ((counts - 400.0) ** 2 / 400.0).sum()
  assert chi2 < 130.0


def test_sampled_energy_matches_exact_expectation(native):
  """End to end on a 12-site chain: sampler + local energy reproduce the
  exact <E> = sum |psi|^2 E_loc / sum |psi|^2 (float64 enumeration) within
  5 standard errors."""
  from gpu_util import make_native
  spec = oansatz.AnsatzSpec('rbm', 12, num_layers=0, layer_size=16)
  params = oansatz.init_params(spec, seed=5, bias_scale=0.1, dtype=F64)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.chain_bonds(12))
  from oracle import ed
  basis = ed.sz0_basis(12)
  all_cfg = torch.from_numpy(bits.unpack(basis.astype(np.uint64)[:, None], 12, np.float64))
  fn = lambda c: oansatz.log_amp(spec, params, c)
  e_all = hamiltonian.local_energy(all_cfg, ij, jx, jz, fn).numpy()
  w = np.exp(2 * fn(all_cfg).numpy())
  exact = float((w * e_all).sum() / w.sum())
  a = make_native(spec, oansatz.flatten(params).numpy())
  ham = native.Hamiltonian(ij, jx, jz, 12)
  packed = native.random_configs(16384, 12, seed=8)
  a.mc_steps(packed, 12 * 40, 99)
  means = []
  for r in range(8):
    e, _ = a.local_energy(ham, packed)
    means.append(e.double().mean().item())
    a.mc_steps(packed, 12 * 4, 99, step0=12 * 40 + r * 48)
  e_np = e.cpu().numpy()
  stderr = e_np.std() / np.sqrt(e_np.size * 8) * 2.0     # generous: correlated samples
  assert abs(np.mean(means) - exact) < 5 * stderr + 1e-4, (np.mean(means), exact, stderr)


@pytest.mark.parametrize('spec', RBM_SHAPES, ids=lambda s: 'N%d_H%d' % (s.n_sites, s.layer_size))
def test_local_energy_vs_oracle(native, spec):
  """|dE| <= 2e-5 * (|diag| + sum_k |jx_k/2| ratio_k): float32 ratios."""
  from gpu_util import packed_cuda
  a, params, cfg = _setup(spec, seed=13, batch=67, scale=0.7)
  n = spec.n_sites
  if n in (36, 100, 256, 16):
    size = int(round(n ** 0.5))
    ij, jx, jz = lattices.j1j2_couplings(size, 0.5)
  else:
    ij, jx, jz = lattices.heisenberg_couplings(lattices.chain_bonds(n), -1.0, 1.0)
  ham = native.Hamiltonian(ij, jx, jz, n)
  e, z, diag, off = a.local_energy(ham, packed_cuda(cfg), want_parts=True)
  cfg64 = torch.from_numpy(cfg).to(F64)
  fn = lambda c: oansatz.log_amp(spec, params, c)
  eo = hamiltonian.local_energy(cfg64, ij, jx, jz, fn).numpy()
  eabs = hamiltonian.local_energy(cfg64, ij, np.abs(jx), np.abs(jz), fn).numpy()
  d_o, _ = hamiltonian.build(cfg64, ij, jx, jz, lambda c: torch.ones(c.shape[0], dtype=F64))
  scale = np.abs(eabs) + np.abs(d_o.numpy()) * 2 + 1.0
  assert np.all(np.abs(e.cpu().numpy() - eo) <= 2e-5 * scale), np.abs(e.cpu().numpy() - eo).max()
  np.testing.assert_allclose(diag.cpu().numpy(), d_o.numpy(), atol=1e-5)
  np.testing.assert_allclose((diag + off).cpu().numpy(), e.cpu().numpy(), atol=1e-5, rtol=1e-6)
  zo = fn(cfg64).numpy()
  assert np.all(np.abs(z.cpu().numpy() - zo) <= 1e-4 + 1e-5 * np.abs(zo))


@pytest.mark.parametrize('name', ['rbm_6x6', 'rbm_4x4_j1j2'])
def test_local_energy_golden(native, name):
  from gpu_util import make_native, packed_cuda
  spec, g = load_golden(name)
  a = make_native(spec, g['params_flat'])
  ham = native.Hamiltonian(g['bonds_ij'], g['bonds_jx'], g['bonds_jz'], spec.n_sites)
  e, z, diag, off = a.local_energy(ham, packed_cuda(g['configs']), want_parts=True)
  scale = np.abs(g['bond_offdiag']).sum(axis=1) / g['psi'] + np.abs(g['ham_diag']) + 1
  assert np.all(np.abs(e.cpu().numpy() - g['local_energy']) <= 5e-5 * scale)
  np.testing.assert_allclose(diag.cpu().numpy(), g['ham_diag'], atol=1e-5)
  # apply_in_place = (diag + off/psi) * psi, operators.py:270-271
  psi = np.exp(z.cpu().numpy().astype(np.float64) - float(g['shift']))
  aip = e.cpu().numpy() * psi
  assert np.all(np.abs(aip - g['apply_in_place']) <= 1e-4 * scale * g['psi'])


def test_local_energy_linear_in_couplings_full_size(native):
  """Size-independent property at the C2 bench size (B = 8192):
  E_loc[2 jx, 2 jz] == 2 E_loc[jx, jz] exactly (power-of-two scaling), and the
  diagonal part equals the bit-count formula."""
  spec = _c2_spec()
  a, params, _ = _setup(spec, seed=3, batch=1)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(6))
  packed = native.random_configs(8192, 36, seed=6)
  h1 = native.Hamiltonian(ij, jx, jz, 36)
  h2 = native.Hamiltonian(ij, 2 * jx, 2 * jz, 36)
  e1, _, d1, _ = a.local_energy(h1, packed, want_parts=True)
  e2, _ = a.local_energy(h2, packed)
  assert torch.equal(2 * e1, e2)
  cfg = bits.unpack(packed.cpu().numpy().view(np.uint64), 36)
  n_act = hamiltonian.n_active(cfg, ij)
  np.testing.assert_allclose(d1.cpu().numpy(), 0.25 * (len(ij) - 2 * n_act), atol=1e-5)


@pytest.mark.parametrize('spec', RBM_SHAPES, ids=lambda s: 'N%d_H%d' % (s.n_sites, s.layer_size))
def test_weighted_grad_sum_vs_oracle(native, spec):
  """S_k = sum_b w_kb O_b against float64 autograd:
  ||dS|| <= 1e-5 * sum_b |w_kb| |O_b| per entry (float32 accumulation)."""
  from gpu_util import packed_cuda
  batch = 333 if spec.n_sites < 200 else 97
  a, params, cfg = _setup(spec, seed=17, batch=batch)
  rng = np.random.default_rng(3)
  w = rng.normal(size=(2, batch)).astype(np.float32)
  w[0] = 1.0
  out = a.weighted_grad_sum(packed_cuda(cfg), torch.from_numpy(w).cuda())
  cfg64 = torch.from_numpy(cfg).to(F64)
  ref = estimators.weighted_grad_sum(spec, params, cfg64, torch.from_numpy(w).to(F64)).numpy()
  bound = estimators.weighted_grad_sum(
      spec, [p.abs() for p in params], cfg64.abs(), torch.from_numpy(np.abs(w)).to(F64)).numpy()
  err = np.abs(out.cpu().numpy() - ref)
  assert np.all(err <= 1e-5 * np.maximum(bound, batch * 1.0) + 1e-5), err.max()
  # accumulate semantics + single weight column + K = 3 path
  out2 = a.weighted_grad_sum(packed_cuda(cfg), torch.from_numpy(w).cuda(), out=out.clone())
  np.testing.assert_allclose(out2.cpu().numpy(), 2 * out.cpu().numpy(), rtol=1e-6, atol=1e-6)
  w3 = np.concatenate([w, w[:1] * 0.5]).astype(np.float32)
  out3 = a.weighted_grad_sum(packed_cuda(cfg), torch.from_numpy(w3).cuda())
  np.testing.assert_allclose(out3[:2].cpu().numpy(), out.cpu().numpy(), rtol=1e-6, atol=1e-6)
  np.testing.assert_allclose(out3[2].cpu().numpy(), 0.5 * out[0].cpu().numpy(), rtol=1e-5, atol=1e-5)


def test_energy_gradient_golden(native):
  """training.py:539-564 for one batch through K3 + K4 + K5."""
  from gpu_util import make_native, packed_cuda
  spec, g = load_golden('rbm_6x6')
  a = make_native(spec, g['params_flat'])
  ham = native.Hamiltonian(g['bonds_ij'], g['bonds_jx'], g['bonds_jz'], spec.n_sites)
  packed = packed_cuda(g['eg_configs'])
  e, _ = a.local_energy(ham, packed)
  w = torch.stack([torch.ones_like(e), e])
  s = a.weighted_grad_sum(packed, w)
  stats = native.energy_stats(e).cpu().numpy()
  mean_e = stats[0] / stats[2]
  assert stats[2] == e.numel()
  assert abs(mean_e - float(g['eg_mean_energy'])) < 2e-5 * (1 + abs(mean_e))
  grad = (s[1] - mean_e * s[0]).cpu().numpy()
  ref = g['eg_gradient']
  assert np.linalg.norm(grad - ref) <= 2e-4 * np.linalg.norm(ref) + 1e-5


def test_swo_gradient_golden(native):
  """training.py:166-175 through K1 + K4 with w = 2 (1 - t/psi) / B."""
  from gpu_util import make_native, packed_cuda
  spec, g = load_golden('rbm_6x6')
  a = make_native(spec, g['params_flat'])
  target = make_native(spec, g['swo_target_params_flat'])
  packed = packed_cuda(g['swo_configs'])
  psi = torch.exp(a.log_amp(packed).double() - float(g['shift']))
  t = torch.exp(target.log_amp(packed).double() - float(g['swo_target_shift'])) * (2.0 ** 18)
  loss = torch.mean((psi - t) ** 2 / psi ** 2)
  assert abs(loss.item() - float(g['swo_loss'])) <= 2e-4 * abs(loss.item()) + 1e-6
  w = (2.0 * (1.0 - t / psi) / psi.numel()).float().reshape(1, -1).contiguous()
  grad = a.weighted_grad_sum(packed, w)[0].cpu().numpy()
  ref = g['swo_gradient']
  assert np.linalg.norm(grad - ref) <= 5e-4 * np.linalg.norm(ref) + 1e-6


@pytest.mark.parametrize('spec', [RBM_SHAPES[0], RBM_SHAPES[2], RBM_SHAPES[3], RBM_SHAPES[6]],
                         ids=lambda s: 'N%d_H%d' % (s.n_sites, s.layer_size))
def test_accumulate_equals_separate_calls(native, spec):
  """cgsvmc_accumulate == local_energy + weighted_grad_sum(1, E) + energy_stats
  (training.py:539-558), including the += semantics and ragged batch sizes."""
  from gpu_util import packed_cuda
  n = spec.n_sites
  ij, jx, jz = lattices.heisenberg_couplings(lattices.chain_bonds(n), -1.0, 1.0)
  for batch in (1, 5, 333, 2500):
    a, params, cfg = _setup(spec, seed=31 + batch, batch=batch, scale=0.7)
    ham = native.Hamiltonian(ij, jx, jz, n)
    packed = packed_cuda(cfg)
    e, z = a.local_energy(ham, packed)
    w = torch.stack([torch.ones_like(e), e]).contiguous()
    ref = a.weighted_grad_sum(packed, w)
    st_ref = native.energy_stats(e)
    sums = torch.zeros(2, a.num_params, device='cuda')
    stats = torch.zeros(4, dtype=torch.float64, device='cuda')
    e_out = torch.empty(batch, device='cuda')
    z_out = torch.empty(batch, device='cuda')
    a.accumulate(ham, packed, sums, stats, e_loc_out=e_out, log_amp_out=z_out)
    # (the two calls may place the tables differently -- shared vs global memory --
    # and then sum the bond terms in rounds of different size)
    assert torch.equal(z_out, z)
    np.testing.assert_allclose(e_out.cpu().numpy(), e.cpu().numpy(), rtol=3e-6, atol=3e-5)
    scale = ref.abs().max().item() + 1.0
    assert float((sums - ref).abs().max()) <= 2e-6 * scale * max(1.0, batch ** 0.5)
    np.testing.assert_allclose(stats.cpu().numpy(), st_ref.cpu().numpy(), rtol=2e-7)   # from e_out vs e
    a.accumulate(ham, packed, sums, stats)
    assert float((sums - 2 * ref).abs().max()) <= 4e-6 * scale * max(1.0, batch ** 0.5)
    assert stats[2].item() == 2 * batch


def test_accumulate_full_size_sharding_invariance(native):
  """Size-independent property at the C2 bench size: accumulating 8192 walkers
  in one call equals accumulating four shards of 2048 (the all-reduce sum of
  the multi-GPU run) up to float32 reassociation."""
  spec = _c2_spec()
  a, params, _ = _setup(spec, seed=3, batch=1)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(6))
  ham = native.Hamiltonian(ij, jx, jz, 36)
  packed = native.random_configs(8192, 36, seed=6)
  sums = torch.zeros(2, a.num_params, device='cuda')
  stats = torch.zeros(4, dtype=torch.float64, device='cuda')
  a.accumulate(ham, packed, sums, stats)
  sums4 = torch.zeros_like(sums)
  stats4 = torch.zeros_like(stats)
  for lo in range(0, 8192, 2048):
    a.accumulate(ham, packed[lo:lo + 2048].contiguous(), sums4, stats4)
  scale = sums.abs().max().item()
  assert float((sums - sums4).abs().max()) <= 2e-5 * scale
  np.testing.assert_allclose(stats.cpu().numpy(), stats4.cpu().numpy(), rtol=1e-10)


def test_in_place_parameter_updates_are_seen(native):
  """The reference's variables are updated in place by apply_gradients
  (training.py:565-567); the derived exp(+-4W) tables must follow."""
  from gpu_util import packed_cuda
  spec = RBM_SHAPES[1]
  a, params, cfg = _setup(spec, seed=5, batch=40)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.chain_bonds(spec.n_sites), -1.0, 1.0)
  ham = native.Hamiltonian(ij, jx, jz, spec.n_sites)
  packed = packed_cuda(cfg)
  e0, _ = a.local_energy(ham, packed)
  a.params.mul_(0.5)                                   # torch in-place update
  e1, _ = a.local_energy(ham, packed)
  params_half = [p * 0.5 for p in params]
  fn = lambda c: oansatz.log_amp(spec, params_half, c)
  eo = hamiltonian.local_energy(torch.from_numpy(cfg).to(F64), ij, jx, jz, fn).numpy()
  np.testing.assert_allclose(e1.cpu().numpy(), eo, rtol=2e-5, atol=2e-5)
  assert float((e1 - e0).abs().max()) > 1e-3
  p2 = packed.clone()
  a.params.mul_(2.0)
  a.mc_steps(p2, 8, 3)
  p3 = packed.clone()
  b, _, _ = _setup(spec, seed=5, batch=40)
  b.mc_steps(p3, 8, 3)
  assert torch.equal(p2, p3)


def test_graphed_batch_step_equals_separate_launches(native):
  """engine.GraphedBatchStep (accumulate + sweep captured as one CUDA graph,
  training.py:614-617) replays bit-identically to the separate launches,
  including after an in-place parameter update between replays."""
  from cgs_vmc_b200 import engine
  spec = _c2_spec()
  a, params, _ = _setup(spec, seed=3, batch=1)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(6))
  ham = native.Hamiltonian(ij, jx, jz, 36)
  s1 = engine.WalkerState(1000, 36, seed=5, walker_id0=7)
  s2 = engine.WalkerState(1000, 36, seed=5, walker_id0=7)
  s1.mc_steps(a, 11)
  s2.mc_steps(a, 11)
  sums1 = engine.EnergyGradientSums(a, 1000)
  sums2 = engine.EnergyGradientSums(a, 1000)
  g = engine.GraphedBatchStep(s1, a, ham, sums1, 36)
  assert torch.equal(s1.packed, s2.packed)            # construction leaves the state untouched
  for rep in range(4):
    if rep == 2:
      a.params.mul_(0.9)                              # optimizer-style in-place update
    g.replay()
    sums2.accumulate(ham, s2.packed)
    s2.mc_steps(a, 36)
    assert torch.equal(s1.packed, s2.packed)
  assert torch.equal(sums1.sums, sums2.sums)
  assert torch.equal(sums1.stats, sums2.stats)
  assert torch.equal(s1.accept_count, s2.accept_count)
  assert s1.step == s2.step and int(s1.step_dev.item()) == s1.step


def _close_sums(got, ref):
  """Same float32 sums up to the summation order: 2e-6 of the largest entry."""
  ref = ref.cpu().numpy()
  np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=0, atol=2e-6 * np.abs(ref).max() + 1e-6)


@pytest.mark.parametrize('n_side,hidden,batch,j1j2', [(6, 144, 777, False), (6, 144, 8192, False), (4, 24, 130, True),
                                                      (16, 256, 300, False), (10, 64, 257, True)])
def test_batch_step_equals_accumulate_then_sweep(native, n_side, hidden, batch, j1j2):
  """cgsvmc_batch_step (estimators + sweep fused in one kernel for the pure
  RBM) against cgsvmc_accumulate followed by cgsvmc_mc_steps: identical
  configurations, acceptance counts, local energies and sums."""
  from cgs_vmc_b200 import engine
  n = n_side * n_side
  spec = oansatz.AnsatzSpec('rbm', n, num_layers=0, layer_size=hidden, size_x=n_side, size_y=n_side)
  a, _, _ = _setup(spec, seed=11, batch=1)
  if j1j2:
    ij, jx, jz = lattices.j1j2_couplings(n_side, 0.5)
  else:
    ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(n_side))
  ham = native.Hamiltonian(ij, jx, jz, n)
  s1 = engine.WalkerState(batch, n, seed=9, walker_id0=3)
  s2 = engine.WalkerState(batch, n, seed=9, walker_id0=3)
  sums1 = engine.EnergyGradientSums(a, batch, want_log_amp=True)
  sums2 = engine.EnergyGradientSums(a, batch, want_log_amp=True)
  for n_steps in (n, 5, 0):
    e1 = sums1.batch_step(ham, s1, n_steps).clone()
    e2 = sums2.accumulate(ham, s2.packed).clone()
    s2.mc_steps(a, n_steps)
    assert torch.equal(s1.packed, s2.packed)
    assert torch.equal(e1, e2)
  assert torch.equal(s1.accept_count, s2.accept_count) and s1.step == s2.step
  np.testing.assert_allclose(sums1.sums.cpu().numpy(), sums2.sums.cpu().numpy(), rtol=1e-6, atol=1e-6)
  assert torch.equal(sums1.stats, sums2.stats)
  assert torch.equal(sums1.log_amp, sums2.log_amp)


@pytest.mark.parametrize('n_side,hidden,batch,n_batches', [(6, 144, 8192, 5), (6, 144, 777, 3), (4, 24, 130, 4),
                                                           (16, 256, 6000, 3), (10, 64, 30000, 2)])
def test_batch_steps_equals_repeated_batch_step(native, n_side, hidden, batch, n_batches):
  """cgsvmc_batch_steps (the inner loop of run_optimization_epoch,
  training.py:614-617, as one persistent kernel) against n_batches calls of
  cgsvmc_batch_step: configurations, acceptance counts and the local energies
  of every iteration bit for bit; gradient sums to float32 summation order
  (the partial sums are reduced once instead of n_batches times), energy
  statistics to float64 rounding.  Covers one walker batch per CTA (walkers
  stay in registers) and several (read back from the packed array)."""
  from cgs_vmc_b200 import engine
  n = n_side * n_side
  spec = oansatz.AnsatzSpec('rbm', n, num_layers=0, layer_size=hidden, size_x=n_side, size_y=n_side)
  a, _, _ = _setup(spec, seed=11, batch=1)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(n_side))
  ham = native.Hamiltonian(ij, jx, jz, n)
  s1 = engine.WalkerState(batch, n, seed=9, walker_id0=3)
  s2 = engine.WalkerState(batch, n, seed=9, walker_id0=3)
  sums1 = engine.EnergyGradientSums(a, batch)
  sums2 = engine.EnergyGradientSums(a, batch)
  e_all = torch.empty(n_batches, batch, dtype=torch.float32, device='cuda')
  for rep in range(2):          # the second call continues from the first (step offsets, accumulation)
    sums1.batch_steps(ham, s1, n, n_batches, e_loc_out=e_all)
    for i in range(n_batches):
      e2 = sums2.batch_step(ham, s2, n).clone()
      assert torch.equal(e_all[i], e2), (rep, i)
    assert torch.equal(s1.packed, s2.packed)
  assert torch.equal(s1.accept_count, s2.accept_count) and s1.step == s2.step
  assert sums1.n_batches == sums2.n_batches
  _close_sums(sums1.sums, sums2.sums)
  np.testing.assert_allclose(sums1.stats.cpu().numpy(), sums2.stats.cpu().numpy(), rtol=1e-12)
  assert float(sums1.stats[2]) == 2 * n_batches * batch


def test_graphed_epoch_equals_graphed_batch_steps(native):
  """engine.GraphedEpoch (one captured persistent kernel per epoch) against
  engine.GraphedBatchStep replayed per batch, across a parameter update."""
  from cgs_vmc_b200 import engine
  spec = _c2_spec()
  a, _, _ = _setup(spec, seed=3, batch=1)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(6))
  ham = native.Hamiltonian(ij, jx, jz, 36)
  B, nb = 8192, 6
  s1 = engine.WalkerState(B, 36, seed=5)
  s2 = engine.WalkerState(B, 36, seed=5)
  sums1, sums2 = engine.EnergyGradientSums(a, B), engine.EnergyGradientSums(a, B)
  g1 = engine.GraphedEpoch(s1, a, ham, sums1, 36, nb)
  g2 = engine.GraphedBatchStep(s2, a, ham, sums2, 36)
  for epoch in range(3):
    if epoch == 2:
      a.params.mul_(0.97)
    g1.replay()
    for _ in range(nb):
      g2.replay()
    assert torch.equal(s1.packed, s2.packed)
    assert s1.step == s2.step and int(s1.step_dev.item()) == s1.step
    _close_sums(sums1.sums, sums2.sums)
    np.testing.assert_allclose(sums1.stats.cpu().numpy(), sums2.stats.cpu().numpy(), rtol=1e-12)
  assert sums1.n_batches == sums2.n_batches == 3 * nb


@pytest.mark.parametrize('n_side,hidden,batch,scale', [(6, 144, 8192, 1.0), (6, 144, 777, 1.0), (6, 100, 300, 1.0),
                                                       (6, 60, 500, 1.0), (6, 60, 16000, 1.0),
                                                       (4, 24, 130, 1.0), (6, 144, 96, 2.0), (6, 144, 96, 3.0)])
def test_tensor_core_gradient_equals_register_tiles(native, n_side, hidden, batch, scale, monkeypatch):
  """The pair-table walker kernel forms S_k = sum_b w_kb sigma_bi tanh theta_bj
  on the tensor cores (bf16 x fp16 split planes, TMEM accumulators) when the
  shape fits; CGSVMC_RBM2_TC_GRAD=0 keeps the FP32 register tiles.  Same
  local energies bit for bit; sums within 2e-6 x sqrt(B) of the largest entry
  (22-bit tanh planes, 33-bit weights; fp32 accumulation both ways).  The
  scaled cases have local energies beyond the plain fp16 range (the operand
  planes hold E_loc 2^-20: |E_loc| < 6.8e10 is representable; beyond that the
  tensor-core sums overflow and CGSVMC_RBM2_TC_GRAD=0 is the way out)."""
  from cgs_vmc_b200 import engine
  n = n_side * n_side
  spec = oansatz.AnsatzSpec('rbm', n, num_layers=0, layer_size=hidden, size_x=n_side, size_y=n_side)
  a, _, _ = _setup(spec, seed=11, batch=1, scale=scale)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(n_side))
  ham = native.Hamiltonian(ij, jx, jz, n)
  res = []
  for flag in ('1', '0'):
    monkeypatch.setenv('CGSVMC_RBM2_TC_GRAD', flag)
    st = engine.WalkerState(batch, n, seed=9, walker_id0=3)
    sums = engine.EnergyGradientSums(a, batch)
    e = []
    for _ in range(3):
      e.append(sums.batch_step(ham, st, n).clone())
    sums.batch_steps(ham, st, n, 2)
    res.append((torch.stack(e), sums.sums.clone(), sums.stats.clone(), st.packed.clone()))
  assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][3], res[1][3])
  assert torch.equal(res[0][2], res[1][2])
  e_max = float(res[0][0].abs().max())
  if not e_max < 6.0e10:
    pytest.skip('local energies up to %.3g: outside the range of the tensor-core operand planes' % e_max)
  ref = res[1][1]
  finite = torch.isfinite(ref)
  assert finite.float().mean() > 0.99 or scale > 1.0
  scale_s = float(ref[finite].abs().max()) + 1.0
  err = float((res[0][1] - ref)[finite].abs().max())
  assert err <= 2e-6 * scale_s * max(1.0, (5 * batch) ** 0.5), (err, scale_s)


def test_pair_tables_of_two_hamiltonians_do_not_evict_each_other(native):
  """The walker kernel keeps one bond-pair table per (ansatz, Hamiltonian);
  a captured graph for one Hamiltonian must stay correct when the same
  ansatz is evaluated with another one in between (and after a parameter
  update seen first by the other Hamiltonian)."""
  from cgs_vmc_b200 import engine
  from gpu_util import packed_cuda
  spec = _c2_spec()
  a, params, cfg = _setup(spec, seed=3, batch=300)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(6))
  ham_x = native.Hamiltonian(ij, jx, jz, 36)
  ij2, jx2, jz2 = lattices.j1j2_couplings(6, 0.5)
  ham_y = native.Hamiltonian(ij2[72:110], jx2[72:110], jz2[72:110], 36)     # some diagonal bonds
  s1 = engine.WalkerState(300, 36, seed=5, packed=packed_cuda(cfg))
  s2 = engine.WalkerState(300, 36, seed=5, packed=packed_cuda(cfg))
  sums1, sums2 = engine.EnergyGradientSums(a, 300), engine.EnergyGradientSums(a, 300)
  g = engine.GraphedBatchStep(s1, a, ham_x, sums1, 36)
  for rep in range(4):
    if rep == 2:
      a.params.mul_(0.9)
    e_y, _ = a.local_energy(ham_y, s2.packed)                 # other Hamiltonian in between
    g.replay()
    e2 = sums2.accumulate(ham_x, s2.packed).clone()
    s2.mc_steps(a, 36)
    assert torch.equal(s1.packed, s2.packed)
    assert torch.equal(sums1.weights[1], e2)
    assert torch.isfinite(e_y).all()
  fn = lambda c: oansatz.log_amp(spec, [p * 0.9 for p in params], c)
  cfg_now = torch.from_numpy(bits.unpack(s2.packed.cpu().numpy().view(np.uint64), 36)).to(F64)
  e_y, _ = a.local_energy(ham_y, s2.packed)
  eo = hamiltonian.local_energy(cfg_now, ij2[72:110], jx2[72:110], jz2[72:110], fn).numpy()
  np.testing.assert_allclose(e_y.cpu().numpy(), eo, rtol=2e-4, atol=2e-4)
  assert torch.equal(sums1.sums, sums2.sums)


@pytest.mark.parametrize('host_pack', [False, True, 'auto'])
def test_host_fed_batch_step(native, host_pack):
  """engine.HostFedBatchStep (pinned host configurations in, energy statistics
  out every batch, gradient sums on request; one graph per buffer slot) gives
  the statistics and sums of accumulate() on the same configurations, batch
  after batch, including after a parameter update -- with the float32 batch
  uploaded as it is, and bit-packed on the host cores first
  (cgsvmc_pack_configs_host)."""
  from cgs_vmc_b200 import engine
  spec = _c2_spec()
  a, _, _ = _setup(spec, seed=3, batch=1)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(6))
  ham = native.Hamiltonian(ij, jx, jz, 36)
  B = 600
  s1 = engine.WalkerState(B, 36, seed=5)
  sums1 = engine.EnergyGradientSums(a, B)
  fed = engine.HostFedBatchStep(s1, a, ham, sums1, 36, host_pack=host_pack)
  assert fed.h2d_bytes == (B * 8 if fed.host_pack else B * 36 * 4)
  ref_sums = engine.EnergyGradientSums(a, B)
  gen = np.random.default_rng(0)
  for k in range(5):
    if k == 3:
      a.params.mul_(0.95)
    cfg = bits.random_sz0_configs(36, B, gen).astype(np.float32)
    host = torch.from_numpy(cfg).pin_memory()
    if k == 2:      # the library's own bit-packed layout from the host
      host = native.pack_configs(torch.from_numpy(cfg).cuda()).cpu().pin_memory()
    fed.submit(host)
    got_stats = fed.result()
    ref_sums.accumulate(ham, native.pack_configs(torch.from_numpy(cfg).cuda()))
    torch.cuda.synchronize()
    np.testing.assert_allclose(got_stats.numpy(), ref_sums.stats.cpu().numpy(), rtol=1e-12)
    if k in (1, 4):
      _close_sums(fed.fetch_sums(), ref_sums.sums)
  assert fed.outstanding() == 0 and s1.step == 5 * 36
  assert int(s1.step_dev.item()) == s1.step


@pytest.mark.parametrize('scale', [6.0, 20.0])
def test_large_weights_take_the_safe_path(native, scale):
  """Weights far beyond the initialisation scale (|W| up to ~2 / ~7): the
  four-term products of the fast ratio loop and the unnormalised sampler
  state overflow float32, and the kernels must fall back to term-by-term
  evaluation.  Local energies stay finite and match the float64 oracle where
  it is finite itself; the sampler keeps Sz and matches replayed oracle ratios."""
  from gpu_util import packed_cuda, unpack_np
  spec = _c2_spec()
  a, params, cfg = _setup(spec, seed=17, batch=96, scale=scale)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(6))
  ham = native.Hamiltonian(ij, jx, jz, 36)
  packed = packed_cuda(cfg)
  e, z = a.local_energy(ham, packed)
  fn = lambda c: oansatz.log_amp(spec, params, c)
  cfg64 = torch.from_numpy(cfg).to(F64)
  eo = hamiltonian.local_energy(cfg64, ij, jx, jz, fn).numpy()
  eabs = hamiltonian.local_energy(cfg64, ij, np.abs(jx), np.abs(jz), fn).numpy()
  ok = np.isfinite(eo) & (np.abs(eabs) < 1e30)
  assert ok.sum() > 0
  got = e.cpu().numpy()
  assert np.all(np.isfinite(got[ok]))
  assert np.all(np.abs(got[ok] - eo[ok]) <= 1e-4 * (np.abs(eabs[ok]) + 20.0)), np.abs(got[ok] - eo[ok]).max()
  # sampler: run, then check conservation and that it moved at all
  before = packed.clone()
  acc = torch.zeros(1, dtype=torch.int64, device='cuda')
  a.mc_steps(packed, 200, seed=3, accept_count=acc)
  after = unpack_np(packed, 36)
  assert np.all(after.sum(axis=1) == 0)
  assert int(acc.item()) > 0 and not torch.equal(before, packed)
  # fused path agrees with the split one under the same conditions
  from cgs_vmc_b200 import engine
  s1 = engine.WalkerState(96, 36, seed=5, packed=before.clone())
  s2 = engine.WalkerState(96, 36, seed=5, packed=before.clone())
  sums1, sums2 = engine.EnergyGradientSums(a, 96), engine.EnergyGradientSums(a, 96)
  e1 = sums1.batch_step(ham, s1, 36).clone()
  e2 = sums2.accumulate(ham, s2.packed).clone()
  s2.mc_steps(a, 36)
  assert torch.equal(s1.packed, s2.packed) and torch.equal(e1, e2)


def test_energy_stats(native):
  e = torch.randn(100003, device='cuda')
  stats = native.energy_stats(e)
  stats = native.energy_stats(e, stats)
  ed = e.double()
  np.testing.assert_allclose(stats.cpu().numpy()[:3],
                             [2 * ed.sum().item(), 2 * (ed * ed).sum().item(), 2 * e.numel()],
                             rtol=1e-12)


def test_error_behaviour(native):
  """Same exception types as the reference for the same conditions."""
  with pytest.raises(ValueError, match='not registered'):     # wavefunctions.py:1196
    native.Ansatz('no_such_type', 8)
  spec = oansatz.AnsatzSpec('rbm', 8, num_layers=0, layer_size=4)
  from gpu_util import make_native
  a = make_native(spec, np.zeros(oansatz.num_params(spec), dtype=np.float32))
  with pytest.raises(ValueError):                              # shape mismatch
    a.log_amp(torch.zeros(4, 2, dtype=torch.int64, device='cuda'))
  with pytest.raises(ValueError):
    native.Hamiltonian([(0, 9)], -1.0, 1.0, 8)
  ham = native.Hamiltonian([(0, 1)], -1.0, 1.0, 12)
  with pytest.raises(ValueError):
    a.local_energy(ham, native.random_configs(4, 8, seed=1))
  # empty batch is a no-op
  assert a.log_amp(torch.zeros(0, 1, dtype=torch.int64, device='cuda')).shape == (0,)


@pytest.mark.parametrize('n_side,hidden,batch', [(6, 144, 8192), (6, 144, 300), (16, 256, 2000), (4, 24, 130)])
def test_in_kernel_reduction_equals_reduction_kernel(native, n_side, hidden, batch, monkeypatch):
  """The cross-CTA reduction at the end of the (cooperatively launched) walker
  kernel against the separate reduction kernel (CGSVMC_RBM2_FUSED_REDUCE=0):
  same trajectories, local energies, statistics and step counter; gradient sums
  equal up to the summation order; and no state is left behind between
  launches (the arrive / depart counters re-arm themselves)."""
  from cgs_vmc_b200 import engine
  n = n_side * n_side
  spec = oansatz.AnsatzSpec('rbm', n, num_layers=0, layer_size=hidden, size_x=n_side, size_y=n_side)
  a, _, _ = _setup(spec, seed=11, batch=1)
  ij, jx, jz = lattices.heisenberg_couplings(lattices.square_nn_bonds(n_side))
  ham = native.Hamiltonian(ij, jx, jz, n)
  out = []
  for fused in ('1', '0'):
    monkeypatch.setenv('CGSVMC_RBM2_FUSED_REDUCE', fused)
    st = engine.WalkerState(batch, n, seed=9, walker_id0=3)
    sums = engine.EnergyGradientSums(a, batch)
    energies = []
    for rep in range(3):
      energies.append(sums.batch_step(ham, st, n).clone())
      sums.accumulate(ham, st.packed)              # the gradient-only launch reduces the same way
    w = torch.ones(1, batch, device='cuda')
    extra = a.weighted_grad_sum(st.packed, w)      # K = 1: odd number of output columns
    torch.cuda.synchronize()
    out.append((st.packed.clone(), torch.stack(energies), sums.sums.clone(), sums.stats.clone(),
                st.accept_count.clone(), extra.clone()))
  (p1, e1, s1, t1, c1, x1), (p0, e0, s0, t0, c0, x0) = out
  assert torch.equal(p1, p0) and torch.equal(e1, e0) and torch.equal(c1, c0)
  assert t1[2].item() == 6 * batch
  np.testing.assert_allclose(t1.cpu().numpy(), t0.cpu().numpy(), rtol=1e-13)
  _close_sums(s1, s0)
  _close_sums(x1, x0)

"""The oracle's restatement of signed output activations and the sum / diff /
prod composites (wavefunctions.py:61-165, 350-353, 1178-1194) against vectors
recorded from the reference (tests/golden/make_golden_composites.py) -- CPU only."""
import numpy as np
import pytest
import torch

from conftest import cmp_golden_names, load_cmp_golden
from oracle import composite

F64 = torch.float64


def test_composite_golden_files_present():
  assert len(cmp_golden_names()) >= 5


@pytest.mark.parametrize('name', cmp_golden_names())
def test_amplitudes_and_local_energy(name):
  kind, leaves, _, g = load_cmp_golden(name)
  cfg = torch.from_numpy(g['configs']).to(F64)
  psi = composite.psi(kind, leaves, cfg).numpy()
  np.testing.assert_allclose(psi, g['psi'], rtol=3e-5, atol=3e-6)
  assert (np.sign(psi) == np.sign(g['psi'])).all()
  e = composite.local_energy(kind, leaves, cfg, g['bonds_ij'], g['bonds_jx'], g['bonds_jz']).numpy()
  scale = 1.0 + np.abs(g['local_energy'])
  assert np.all(np.abs(e - g['local_energy']) <= 2e-4 * scale * (1.0 + 1e-2 / np.abs(g['psi'])))
  np.testing.assert_allclose(e * psi, g['apply_in_place'], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize('name', cmp_golden_names())
def test_energy_gradient(name):
  kind, leaves, _, g = load_cmp_golden(name)
  cfg = torch.from_numpy(g['eg_configs']).to(F64)
  grad, e_mean = composite.energy_gradient(kind, leaves, cfg, g['bonds_ij'], g['bonds_jx'], g['bonds_jz'])
  assert abs(e_mean - float(g['eg_mean_energy'])) <= 3e-4 * (1 + abs(e_mean))
  ref = g['eg_gradient']
  assert grad.numel() == ref.size
  np.testing.assert_allclose(grad.numpy(), ref, rtol=0, atol=1e-3 * np.abs(ref).max())

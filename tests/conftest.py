"""pytest configuration: markers and shared fixtures."""
import glob
import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
  sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _names():
  return sorted(os.path.basename(p)[:-4]
                for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def golden_names():
  """Cases of make_golden.py (amplitudes, sampler, Hamiltonian, EnergyGradient, SWO)."""
  return [n for n in _names() if not n.startswith(('opt_', 'cmp_', 'layers_'))]


def opt_golden_names():
  """Cases of make_golden_optimizers.py (LogOverlapSWO, DualSamplingSWO,
  LogOverlapImaginaryTimeSWO)."""
  return [n for n in _names() if n.startswith('opt_')]


def load_golden(name):
  """Returns (spec, dict of arrays) for tests/golden/<name>.npz."""
  from oracle import ansatz
  data = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz')))
  spec = ansatz.AnsatzSpec(**json.loads(str(data.pop('spec_json'))))
  return spec, data


@pytest.fixture(params=golden_names())
def golden(request):
  spec, data = load_golden(request.param)
  return request.param, spec, data


def cmp_golden_names():
  """Cases of make_golden_composites.py (signed outputs, sum / diff / prod)."""
  return [n for n in _names() if n.startswith('cmp_')]


def load_cmp_golden(name):
  """Returns (kind, oracle leaves (float64), hparams overrides, arrays)."""
  import torch
  from oracle import ansatz, composite
  data = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz')))
  case = json.loads(str(data.pop('case_json')))
  hp = case['hparams']
  kind = hp['wavefunction_type'] if hp['wavefunction_type'] in ('sum', 'diff', 'prod') else 'single'
  acts = hp.get('composite_output_activations') or [hp.get('output_activation', 'exp')]
  leaves, off = [], 0
  for k, spec_d in enumerate(case['leaf_specs']):
    spec = ansatz.AnsatzSpec(**spec_d)
    n = int(data['leaf_sizes'][k])
    flat = torch.from_numpy(data['leaf_params_flat'][off:off + n]).to(torch.float64)
    off += n
    shift = float(data['leaf_shifts'][k])
    leaves.append(composite.Leaf(spec, ansatz.unflatten(spec, flat), acts[k],
                                 shift if acts[k] == 'exp' else None))
  return kind, leaves, hp, data


@pytest.fixture(params=opt_golden_names())
def opt_golden(request):
  spec, data = load_golden(request.param)
  return request.param, spec, data

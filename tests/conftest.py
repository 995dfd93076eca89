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
  return [n for n in _names() if not n.startswith('opt_')]


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


@pytest.fixture(params=opt_golden_names())
def opt_golden(request):
  spec, data = load_golden(request.param)
  return request.param, spec, data

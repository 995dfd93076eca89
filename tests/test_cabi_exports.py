"""CPU-side checks of the boundary: the shared library loads and exports every
symbol include/cgsvmc.h declares; no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  text = open(os.path.join(REPO, 'include', 'cgsvmc.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(cgsvmc_[a-z_0-9]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
  import __graft_entry__
  if not os.path.exists(__graft_entry__.LIB):
    __graft_entry__.build()
  return ctypes.CDLL(__graft_entry__.LIB)


def test_header_declares_expected_entry_points():
  syms = _declared_symbols()
  for must in ('cgsvmc_mc_steps', 'cgsvmc_local_energy', 'cgsvmc_log_amp',
               'cgsvmc_weighted_grad_sum', 'cgsvmc_flip_enum', 'cgsvmc_ansatz_create'):
    assert must in syms
  assert len(syms) >= 18


def test_library_exports_every_declared_symbol(lib):
  for name in _declared_symbols():
    assert hasattr(lib, name), name


def test_binding_lists_every_declared_symbol():
  from cgs_vmc_b200 import _native
  assert sorted(_native.EXPORTS) == _declared_symbols()


def test_version_and_error_string(lib):
  lib.cgsvmc_version.restype = ctypes.c_int
  assert lib.cgsvmc_version() == 100
  lib.cgsvmc_last_error.restype = ctypes.c_char_p
  assert isinstance(lib.cgsvmc_last_error(), bytes)


def test_argument_errors_without_gpu(lib):
  """Argument validation happens before any CUDA call."""
  lib.cgsvmc_last_error.restype = ctypes.c_char_p
  assert lib.cgsvmc_ansatz_create(None, None) == -1
  out = ctypes.c_void_p()
  assert lib.cgsvmc_ham_create(None, None, None, 3, 8, ctypes.byref(out)) == -1
  assert b'NULL' in lib.cgsvmc_last_error()
  assert lib.cgsvmc_log_amp(None, None, ctypes.c_int64(4), None, None) == -1


def test_product_fails_loudly_without_cuda():
  import torch
  if torch.cuda.is_available():
    pytest.skip('CUDA present')
  from cgs_vmc_b200 import _native
  with pytest.raises(_native.NativeError, match='no CPU fallback'):
    _native.Ansatz('rbm', 8, layer_size=4)


@pytest.mark.parametrize('shape', [(1, 1), (7, 20), (300, 36), (64, 64), (33, 65), (20, 100), (9, 256), (5000, 36)])
@pytest.mark.parametrize('threads', [0, 1, 3])
def test_pack_configs_host_matches_oracle_layout(shape, threads):
  """cgsvmc_pack_configs_host is host code (layout conversion of the
  reference's float32 [B, N] tensor before the upload): bit-exact against the
  oracle's packed layout, for ragged word counts and any thread count."""
  import numpy as np
  import torch
  from cgs_vmc_b200 import _native
  from oracle import bits
  b, n = shape
  rng = np.random.default_rng(b * 1000 + n)
  cfg = rng.choice([-1.0, 1.0], size=(b, n)).astype(np.float32)
  out = torch.full((b, (n + 63) // 64), -1, dtype=torch.int64)
  _native.pack_configs_host(torch.from_numpy(cfg), out, threads)
  assert np.array_equal(out.numpy().view(np.uint64), bits.pack(cfg).reshape(b, -1))
  with pytest.raises(ValueError):
    _native.pack_configs_host(torch.from_numpy(cfg), out[:, :0].contiguous(), threads)

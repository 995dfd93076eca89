"""The oracle's restatement of LogOverlapSWO, DualSamplingSWO and
LogOverlapImaginaryTimeSWO (training.py:298-503, 626-778) against gradients
recorded from the reference's own build_opt_ops
(tests/golden/make_golden_optimizers.py) -- CPU only."""
import numpy as np
import torch

from oracle import ansatz, estimators, hamiltonian

F64 = torch.float64
GRAD_RTOL = 2e-4    # of the largest entry; the reference ran in float32


def _params(spec, flat):
  return ansatz.unflatten(spec, torch.from_numpy(flat).to(F64))


def _close(got, ref, rtol=GRAD_RTOL):
  scale = float(np.abs(ref).max())
  assert scale > 0
  np.testing.assert_allclose(np.asarray(got), ref, rtol=0, atol=rtol * scale)


def test_optimizer_golden_files_present():
  from conftest import opt_golden_names
  assert len(opt_golden_names()) >= 5


def test_log_overlap_swo_gradient(opt_golden):
  _, spec, g = opt_golden
  params, tparams = _params(spec, g['params_flat']), _params(spec, g['target_params_flat'])
  cfg = torch.from_numpy(g['lo_configs']).to(F64)
  psi = ansatz.psi(spec, params, cfg, shift=float(g['shift']))
  psi_t = ansatz.psi(spec, tparams, cfg, shift=float(g['target_shift']))
  grad = estimators.log_overlap_grad(spec, params, cfg, psi_t / psi)
  _close(grad.numpy(), g['lo_gradient'])


def test_dual_sampling_swo_loss_and_gradient(opt_golden):
  _, spec, g = opt_golden
  params, tparams = _params(spec, g['params_flat']), _params(spec, g['target_params_flat'])
  cfg = torch.from_numpy(np.concatenate([g['ds_psi_configs'], g['ds_target_configs']])).to(F64)
  assert cfg.shape[0] == 12 and g['ds_psi_configs'].shape[0] == 6    # batch_size // 2 each
  psi_t = ansatz.psi(spec, tparams, cfg, shift=float(g['ds_target_shift']))
  loss, grad = estimators.dual_sampling_loss_and_grad(spec, params, cfg, psi_t,
                                                      shift=float(g['shift']))
  assert abs(float(loss) - float(g['ds_loss'])) <= 1e-4 * abs(float(g['ds_loss']))
  _close(grad.numpy(), g['ds_gradient'])


def test_log_overlap_imaginary_time_swo(opt_golden):
  _, spec, g = opt_golden
  params, oparams = _params(spec, g['params_flat']), _params(spec, g['it_omega_params_flat'])
  cfg = torch.from_numpy(g['it_configs']).to(F64)
  psi = ansatz.psi(spec, params, cfg, shift=float(g['shift']))
  psi_o = ansatz.psi(spec, oparams, cfg, shift=float(g['it_omega_shift']))
  e_o = hamiltonian.local_energy(cfg, g['bonds_ij'], g['bonds_jx'], g['bonds_jz'],
                                 lambda c: ansatz.log_amp(spec, oparams, c))
  assert abs(float(e_o.mean()) - float(g['it_energy'])) <= 2e-5 * (1 + abs(float(g['it_energy'])))
  ratio = estimators.imaginary_time_ratio(psi, psi_o, e_o, float(g['it_beta']))
  grad = estimators.log_overlap_grad(spec, params, cfg, ratio)
  _close(grad.numpy(), g['it_gradient'])

import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from oracle import ansatz as oansatz, bits, estimators
from gpu_util import make_native, packed_cuda
F64 = torch.float64
spec = oansatz.AnsatzSpec('fully_connected', 100, num_layers=4, layer_size=48)
for batch in (300, 2000, 20000):
  params = oansatz.init_params(spec, seed=17 + batch, bias_scale=0.1, dtype=F64)
  cfg = bits.random_sz0_configs(spec.n_sites, batch, np.random.default_rng(17 + batch))
  a = make_native(spec, oansatz.flatten(params).numpy())
  w = np.random.default_rng(3).normal(size=(2, batch)).astype(np.float32); w[0] = 1.0
  packed = packed_cuda(cfg); wt = torch.from_numpy(w).cuda()
  out = a.weighted_grad_sum(packed, wt).cpu().numpy()
  out2 = a.weighted_grad_sum(packed, wt).cpu().numpy()
  os.environ['CGSVMC_FC_TC_GRAD'] = '0'
  simt = a.weighted_grad_sum(packed, wt).cpu().numpy()
  os.environ.pop('CGSVMC_FC_TC_GRAD')
  ref = estimators.weighted_grad_sum(spec, params, torch.from_numpy(cfg).to(F64), torch.from_numpy(w).to(F64)).numpy()
  shapes = [s for _, s in oansatz.param_shapes(spec)]
  off = 0
  print('batch', batch, 'repeatable', np.array_equal(out, out2))
  for name_shape in oansatz.param_shapes(spec):
    n = int(np.prod(name_shape[1]))
    for k in range(2):
      sl = slice(off, off + n)
      e_tc = np.abs(out[k][sl] - ref[k][sl]).max(); e_s = np.abs(simt[k][sl] - ref[k][sl]).max()
      print('  %-8s k=%d  max|ref| %9.3f  err tc %.3e  err simt %.3e  argmax %d' % (
          name_shape[0], k, np.abs(ref[k][sl]).max(), e_tc, e_s, int(np.abs(out[k][sl] - ref[k][sl]).argmax())))
    off += n

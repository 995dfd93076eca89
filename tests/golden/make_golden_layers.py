"""Generates tests/golden/layers_periodic.npz by calling the UNMODIFIED
reference modules layers.Conv1dPeriodic / layers.Conv2dPeriodic
(/root/reference/cgs_vmc/layers.py:24-160) on random inputs, on top of the eager
TensorFlow / Sonnet stand-in of tests/golden/tf_shim (see make_golden.py).

    python tests/golden/make_golden_layers.py

Recorded per case: input, the module's variables (w, b -- b overwritten with
random values so the bias path is exercised) and the module's output.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
sys.path.insert(0, '/root/reference/cgs_vmc')

import tensorflow as tf              # noqa: E402  (the shim)
import layers                        # noqa: E402  (reference)

CASES = [   # rank, spatial shape, c_in, c_out, kernel
    (1, (7,), 1, 5, 3), (1, (10,), 3, 16, 4), (1, (12,), 2, 3, 5), (1, (9,), 4, 6, 6),
    (2, (4, 6), 1, 5, 2), (2, (10, 10), 1, 16, 5), (2, (6, 6), 16, 16, 3), (2, (5, 7), 3, 4, 4),
]


def main():
  out = {}
  gen = torch.Generator().manual_seed(2024)
  for n, (rank, shape, c_in, c_out, k) in enumerate(CASES):
    x = torch.randn((6,) + shape + (c_in,), generator=gen)
    cls = layers.Conv1dPeriodic if rank == 1 else layers.Conv2dPeriodic
    module = cls(c_out, k, name='layer_%d' % n)
    module(x)                                           # creates the variables
    variables = [t for name, t in tf._VARIABLES.items() if name.startswith('layer_%d/' % n)]
    assert len(variables) == 2, [name for name in tf._VARIABLES]
    w, b = sorted(variables, key=lambda v: -v.dim())
    tf.assign(b, torch.randn(c_out, generator=gen))
    y = module(x)
    out['case%d_meta' % n] = np.array([rank, c_in, c_out, k], dtype=np.int64)
    out['case%d_x' % n] = x.numpy()
    out['case%d_w' % n] = w.detach().numpy().copy()
    out['case%d_b' % n] = b.detach().numpy().copy()
    out['case%d_y' % n] = y.detach().numpy().copy()
  np.savez_compressed(os.path.join(HERE, 'layers_periodic.npz'), **out)
  print('wrote layers_periodic.npz with', len(CASES), 'cases')


if __name__ == '__main__':
  main()

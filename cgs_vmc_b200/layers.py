"""Mirror of the reference's layers.py for the in-scope pieces.

Inside the conv ansaetze the periodic convolutions are fused into whole-network
kernels (csrc/conv_tc.cu, csrc/net.cu: wrap-padded images / wrap tables instead
of the concat padding of layers.py:51-74 and 117-148); there the classes below
carry the hyper-parameters of a layer.  Called by themselves they are the
reference's modules: `Conv2dPeriodic(channels, k)(inputs)` pads periodically
and convolves on the device (cgsvmc_conv_periodic), with Sonnet's default
initialisation on first use.
"""
import math

import torch

from . import _native

NONLINEARITIES = {   # layers.py:13-21; values are the kernel-side names
    'relu': 'relu', 'exp': 'exp', 'cos': 'cos', 'tan': 'tan', 'tanh': 'tanh',
    'sigmoid': 'sigmoid', 'identity': 'identity',
}


class _ConvPeriodic:
  rank = None

  def __init__(self, output_channels, kernel_shape, stride=1, name=None):
    if stride != 1:
      raise NotImplementedError('periodic convolutions are built for stride 1 '
                                '(from_hparams never passes another value)')
    self._output_channels = output_channels
    self._kernel_shape = kernel_shape
    self._stride = stride
    self.name = name or 'conv_%dd_periodic' % self.rank
    self.w = None   # views into the owning ansatz' flat parameter buffer
    self.b = None

  def initialize(self, input_channels, device='cuda', generator=None):
    """Sonnet v1 defaults (snt.Conv1D / Conv2D): weights truncated normal within
    two sigma, sigma = 1 / sqrt(fan_in); zero bias."""
    k = self._kernel_shape
    shape = (k,) * self.rank + (int(input_channels), self._output_channels)
    sigma = 1.0 / math.sqrt(k ** self.rank * int(input_channels))
    w = torch.empty(shape)
    torch.nn.init.trunc_normal_(w, 0.0, sigma, -2 * sigma, 2 * sigma, generator=generator)
    self.w = w.to(device)
    self.b = torch.zeros(self._output_channels, device=device)
    return self

  def __call__(self, inputs):
    """inputs: float32 CUDA tensor [B, L, C] (rank 1) / [B, X, Y, C] (rank 2)
    -> [B, ..., output_channels]; layers.py:76-80 / 150-160."""
    if inputs.dim() != self.rank + 2:
      raise ValueError('Input tensor has wrong shape.')
    if self.w is None:
      self.initialize(inputs.shape[-1], device=inputs.device)
    return _native.conv_periodic(inputs.contiguous(), self.w.contiguous(), self.b)

  def pad_sizes(self):
    """(before, after) wrap padding per axis, layers.py:64-73 / 132-141."""
    k = self._kernel_shape
    if k % 2 == 1:
      return (k - 1) // 2, (k - 1) // 2
    return (k // 2, k // 2 - 1) if self.rank == 1 else (k // 2 - 1, k // 2)


class Conv1dPeriodic(_ConvPeriodic):
  rank = 1


class Conv2dPeriodic(_ConvPeriodic):
  rank = 2

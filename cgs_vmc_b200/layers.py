"""Mirror of the reference's layers.py for the in-scope pieces.

The periodic convolutions are fused into the conv ansatz kernels
(csrc/net.cu, wrap tables instead of the concat padding of layers.py:51-74 and
117-148), so the classes here only carry the hyper-parameters and own their
parameter views; they are not separately callable on the device.
"""

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

"""Checkpoints with the reference's naming and retention
(tf.train.Saver(wavefunction.get_trainable_variables(), max_to_keep=5),
run_training.py:134-146).  Like the reference only the trainable variables are
saved (not Adam slots, num_epochs, walkers); the
exp_norm_shift of every leaf is stored additionally (the reference loses it
on restore, which changes a sum / difference of wavefunctions)."""
import os

import torch


class Saver:
  def __init__(self, wavefunction, max_to_keep=5):
    self._wf = wavefunction
    self._max_to_keep = max_to_keep
    self._kept = []

  def save(self, session, save_path):
    del session
    path = save_path + '.pt'
    variables = [v.detach().cpu().clone() for v in self._wf.get_trainable_variables()]
    # one shift per leaf: in a sum / difference the relative scale of the parts matters
    shifts = [leaf._exp_norm_shift for leaf in self._wf.leaves()]
    torch.save({'variables': variables, 'exp_norm_shift': self._wf._exp_norm_shift,
                'leaf_exp_norm_shifts': shifts}, path)
    self._kept.append(path)
    while self._max_to_keep and len(self._kept) > self._max_to_keep:
      old = self._kept.pop(0)
      if os.path.exists(old):
        os.remove(old)
    with open(os.path.join(os.path.dirname(path), 'checkpoint'), 'w') as f:
      f.write(os.path.basename(path) + '\n')
    return path

  def restore(self, session, path):
    del session
    if path is None or not os.path.exists(path):
      raise ValueError('checkpoint not found: %r' % (path,))
    data = torch.load(path, map_location='cpu')
    variables = self._wf.get_trainable_variables()
    if len(variables) != len(data['variables']):
      raise ValueError('checkpoint does not match the wavefunction structure')
    for dst, src in zip(variables, data['variables']):
      if tuple(dst.shape) != tuple(src.shape):
        raise ValueError('checkpoint variable shape %s != %s' % (tuple(src.shape), tuple(dst.shape)))
      dst.copy_(src.to(dst.device))
    shifts = data.get('leaf_exp_norm_shifts')
    leaves = self._wf.leaves()
    if shifts is not None and len(shifts) == len(leaves):
      for leaf, shift in zip(leaves, shifts):
        if shift is not None:
          leaf._exp_norm_shift = float(shift)


def latest_checkpoint(checkpoint_dir):
  marker = os.path.join(checkpoint_dir, 'checkpoint')
  if not os.path.exists(marker):
    return None
  with open(marker) as f:
    name = f.read().strip()
  return os.path.join(checkpoint_dir, name)

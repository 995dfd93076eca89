"""Stand-in for tf.Session so the reference's driver loops read the same.

The reference wires a graph once (`build_opt_ops`) and then executes named
ops with `session.run(op)` (run_training.py:129-148, training.py:608-623).
Here `build_opt_ops` returns `Op` objects bound to device state; `run`
executes them (each is one or a few CUDA launches) and returns their value.
"""
import torch


class Op:
  """A deferred computation: `session.run(op)` calls it."""

  def __init__(self, fn, name=''):
    self._fn = fn
    self.name = name

  def __call__(self, *args, **kwargs):
    return self._fn(*args, **kwargs)

  def __repr__(self):
    return '<Op %s>' % self.name


class Session:
  def run(self, fetches, **kwargs):
    if fetches is None:
      return None
    if isinstance(fetches, (list, tuple)):
      return [self.run(f, **kwargs) for f in fetches]
    if isinstance(fetches, Op):
      value = fetches(**kwargs)
    elif callable(fetches) and kwargs:
      value = fetches(**kwargs)
    else:
      value = fetches
    if isinstance(value, torch.Tensor) and value.dim() == 0:
      return value.item()
    return value

  def close(self):
    pass

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    return False

"""Placeholder for the HParamDef proto imported by the reference's utils.py:13
(only utils.load_hparams uses it; the golden generator never calls that)."""


class HParamDef:
  pass

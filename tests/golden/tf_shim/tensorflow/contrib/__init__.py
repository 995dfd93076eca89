"""tf.contrib namespace of the eager test shim (see tensorflow/__init__.py)."""
from . import training   # noqa: F401
import types as _types
data = _types.SimpleNamespace(CsvDataset=object)

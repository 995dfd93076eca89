"""CPU oracle for the cgs-vmc variational-Monte-Carlo hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product path (``cgs_vmc_b200``) never imports this package and raises when its
CUDA library is missing.

What it is: a plain restatement (numpy for the bit work, torch-on-CPU for the
floating point so that float32 and float64 share one code path and autograd
provides the gradient oracle) of the reference algorithm, each function citing
the reference ``file:line`` it follows (paths relative to
``/root/reference/cgs_vmc``).

How parity is pinned: the reference ships no tests, golden vectors or
fixtures, and TensorFlow 1.x / Sonnet v1 cannot be
installed here.  Instead ``tests/golden/make_golden.py`` executes the
UNMODIFIED reference Python modules (``wavefunctions.py``, ``layers.py``,
``operators.py``, ``graph_builders.py``, ``training.py``) from
``/root/reference`` on top of an eager stand-in for the TensorFlow / Sonnet op
set (``tests/golden/tf_shim``; documented op semantics implemented with
torch-CPU float32) and commits the resulting input/output vectors under
``tests/golden/*.npz``.  The oracle is checked against those vectors and
against exact-diagonalisation known answers (``oracle/ed.py``) by the
``-m "not gpu"`` tests.  The third-party arithmetic itself (TF's Eigen
kernels) is restated, not run: that residual gap is stated in DESIGN.md.
"""

from . import bits, lattices, ansatz, sampler, hamiltonian, estimators, ed  # noqa: F401

"""cgs-vmc hot path on B200: Python host mirror of the reference's
Wavefunction / Operator / graph_builders API over the hand-written sm_100a
CUDA library `libcgsvmc.so` (C-ABI: include/cgsvmc.h).

There is no CPU fallback: every compute entry point raises if the CUDA
library is missing or no CUDA device is present.
"""
__version__ = '0.1.0'

"""lbzip2_b200 -- B200-native bzip2 block-compression engine.

The product is the C-ABI shared library ``libbz2b200.so`` (hand-written
sm_100a CUDA kernels, see ``csrc/`` and ``include/lbzip2_b200.h``).  This
package is only a thin ctypes mirror of that ABI for tests and bench.py; it
never falls back to a CPU implementation: importing :mod:`lbzip2_b200.api`
raises if the library is missing, and creating an engine raises if there is
no usable GPU.
"""
from .api import Decoder, Engine, LbzError, load_library, lib_path  # noqa: F401

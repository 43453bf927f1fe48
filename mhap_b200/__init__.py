"""mhap_b200 -- B200-native MinHash sketch + overlap search behind marbl/MHAP's seams.

Only what the hot path needs: csrc/ (hand-written sm_100a kernels + the C ABI of
include/mhap_b200.h), host/ (the C++ command-line driver mirroring MhapMain's -s / -q / -p modes),
native.py (ctypes binding of that ABI), distributed.py (one-rank-per-GPU callers of the ABI's
multi-GPU entry points), synth.py (synthetic PacBio-shape reads for tests and bench.py).  The oracle under oracle/ is never imported from here.
"""
from .native import Engine, MhapError, SearchParams, SketchParams, pack_reads  # noqa: F401

__version__ = "0.1"

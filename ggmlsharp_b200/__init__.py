"""ggmlsharp_b200 -- B200 (sm_100a) backend for GGMLSharp's matrix-multiply hot path.

The product is ``lib/libggb200.so`` (hand-written CUDA kernels behind the C ABI of
``include/ggb200.h``) plus ``lib/libggml_host.so`` (a C++ mirror of the reference's host API for
this path).  This package is a thin ctypes front-end used by tests and bench.py; it never falls
back to a CPU implementation: a missing library raises ImportError-like errors loudly.
"""
from . import native  # noqa: F401
from .native import GgbError, lib, host  # noqa: F401

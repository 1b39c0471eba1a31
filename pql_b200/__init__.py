"""pql_b200 - B200 (sm_100a) native learner hot path for Parallel Q-Learning.

Mirrors the reference's module layout for the path it replaces
(`pql.replay`, `pql.models`, `pql.algo.pql_{v,p}_learner`, `pql.utils`), with every
operator executed by hand-written CUDA kernels reached through the C ABI in
`include/pqlb200.h` (`pql_b200/libpqlb200.so`).  There is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"

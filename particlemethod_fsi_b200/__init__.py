"""particlemethod_fsi_b200 -- B200-native explicit MPH / total-Lagrangian FSI step (host side).

The first call into `lib` loads libmphx.so (hand-written sm_100a CUDA behind the extern-"C" layer of
include/mphx.h).  There is no CPU fallback: a missing library is an ImportError at that call.
"""
from . import abi, cases  # noqa: F401
from .solver import Solver, MphxError, lib  # noqa: F401
from . import solver  # noqa: F401

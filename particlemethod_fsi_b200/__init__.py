"""particlemethod_fsi_b200 -- B200-native explicit MPH / total-Lagrangian FSI step (host side)."""
from . import abi, cases  # noqa: F401

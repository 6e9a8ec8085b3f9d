"""opm_simulators_b200 -- B200-native drop-in for OPM Flow's per-Newton-step linear solve
(BiCGSTAB + ILU0/DILU on block-CSR Jacobians).  The compute path is libopmb200.so (hand-written
CUDA for sm_100a behind the C ABI of include/opmb200.h); this package is the thin host side:
ctypes binding, the Dune-shaped FlexibleSolver / PreconditionerFactory / PropertyTree mirrors,
file formats and synthetic system generators.  There is no CPU fallback.
"""
from .bcsr import BCSR  # noqa: F401

__all__ = ["BCSR"]

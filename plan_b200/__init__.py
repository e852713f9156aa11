"""plan_b200 -- B200 (sm_100a) execution hot path for daviszhen/plan.

The product is libplangpu.so (hand-written CUDA behind the C ABI of include/plangpu.h).
This package is the thin host side: ctypes bindings plus a mirror of the reference's
operator interface (OperatorExec / PhysicalOperator / Expr / Chunk) used by tests and
benchmarks where the Go toolchain is unavailable.  There is NO CPU fallback: importing
the bindings fails loudly when the CUDA library is missing.
"""
from . import _lib  # noqa: F401
from ._lib import PlanGpuError, lib  # noqa: F401

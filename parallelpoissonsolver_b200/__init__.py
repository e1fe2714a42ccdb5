"""parallelpoissonsolver_b200 -- B200-native hot path of lucak17/ParallelPoissonSolver.

The product is the C-ABI library ``csrc/libpps_b200.so`` (include/pps_b200.h) plus the C++ host driver in
``driver/``; this Python package is only the thin ctypes binding the tests and ``bench.py`` drive it with.
There is no CPU fallback: importing works anywhere, but creating a solver needs the built library and a GPU.
"""
from .api import (  # noqa: F401
    ARITH_FAST, ARITH_PARITY, CHEB_EIG_GLOBAL, CHEB_EIG_LOCAL, CHEB_FP32, CHEB_FP64, FLAG_NO_DOT_VECTOR, FLAG_OPERATOR_ONLY, FUSE_AUTO, FUSE_FULL, FUSE_SPLIT, PRECOND_BICGSTAB_LOCAL, PRECOND_CG_CHEB_LOCAL,
    PRECOND_CHEBYSHEV, PRECOND_NONE, SOLVER_BICGSTAB, SOLVER_CG, SOLVER_CHEBYSHEV, BlockInfo, Config, PoissonSolver, PpsError, default_config, get_unique_id, library_path, load_library,
    make_config,
)

__all__ = [
    "ARITH_FAST", "ARITH_PARITY", "CHEB_EIG_GLOBAL", "CHEB_EIG_LOCAL", "CHEB_FP32", "CHEB_FP64", "FLAG_NO_DOT_VECTOR", "FLAG_OPERATOR_ONLY", "FUSE_AUTO", "FUSE_FULL", "FUSE_SPLIT", "PRECOND_BICGSTAB_LOCAL", "PRECOND_CG_CHEB_LOCAL",
    "PRECOND_CHEBYSHEV", "PRECOND_NONE", "SOLVER_BICGSTAB", "SOLVER_CG", "SOLVER_CHEBYSHEV", "BlockInfo", "Config", "PoissonSolver", "PpsError", "default_config",
    "get_unique_id", "library_path", "load_library", "make_config",
]

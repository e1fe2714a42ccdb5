"""ctypes binding of the CPU oracle (oracle/pps_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
module; the product path (parallelpoissonsolver_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpps_oracle.so")

SOLVER_BICGSTAB, SOLVER_CG, SOLVER_CHEBYSHEV = 0, 1, 2
PRECOND_NONE, PRECOND_CHEBYSHEV, PRECOND_BICGSTAB_LOCAL, PRECOND_CG_CHEB_LOCAL = 0, 1, 2, 3


class OrcConfig(C.Structure):
    _fields_ = [
        ("np", C.c_int * 3), ("nranks", C.c_int * 3), ("ds", C.c_double * 3), ("origin", C.c_double * 3),
        ("bcs", C.c_int * 6), ("solver", C.c_int), ("precond", C.c_int), ("tolerance", C.c_double),
        ("max_iter", C.c_int), ("cheb_max", C.c_int), ("cheb_epsilon", C.c_double),
        ("cheb_rescale_min", C.c_double), ("cheb_rescale_max", C.c_double),
        ("precond_tolerance", C.c_double), ("precond_max_iter", C.c_int), ("order_neumann", C.c_int), ("dim", C.c_int),
        ("cheb_eig_local", C.c_int), ("cheb_f32", C.c_int), ("precond_comm", C.c_int),
    ]


class OrcBlockInfo(C.Structure):
    _fields_ = [
        ("rank", C.c_int), ("loc", C.c_int * 3), ("nlocal", C.c_int * 3), ("nguards", C.c_int * 3),
        ("limits_data", C.c_int * 6), ("limits_solver", C.c_int * 6), ("has_boundary", C.c_int * 6),
        ("has_comm", C.c_int * 6), ("ntot", C.c_long),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "pps_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libpps_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        P = C.c_void_p
        D = C.POINTER(C.c_double)
        L.orc_default_config.argtypes = [C.POINTER(OrcConfig)]
        L.orc_create.restype = P
        L.orc_create.argtypes = [C.POINTER(OrcConfig)]
        L.orc_destroy.argtypes = [P]
        L.orc_world.argtypes = [P]
        L.orc_block.argtypes = [P, C.c_int, C.POINTER(OrcBlockInfo)]
        L.orc_eigenvalues.argtypes = [P, C.c_int, D, D]
        for f in ("orc_exact_u", "orc_exact_f"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double] * 3
        L.orc_exact_dudn.restype = C.c_double
        L.orc_exact_dudn.argtypes = [C.c_double] * 3 + [C.c_int]
        L.orc_x.restype = D
        L.orc_x.argtypes = [P, C.c_int]
        L.orc_b.restype = D
        L.orc_b.argtypes = [P, C.c_int]
        L.orc_zero_fields.argtypes = [P]
        L.orc_set_problem.argtypes = [P]
        L.orc_neumann_face.restype = C.c_long
        L.orc_neumann_face.argtypes = [P, C.c_int, C.c_int, D]
        L.orc_apply.argtypes = [P, C.c_int, D, D]
        L.orc_halo_exchange.argtypes = [P, C.POINTER(D)]
        L.orc_reset_neumann.argtypes = [P, C.c_int, D, C.c_int, C.c_double]
        L.orc_adjust_b.argtypes = [P, C.c_int, D, D]
        L.orc_precondition.argtypes = [P, C.POINTER(D), C.POINTER(D)]
        L.orc_solve.argtypes = [P]
        L.orc_iters.argtypes = [P]
        for f in ("orc_error_iteration", "orc_error_operator", "orc_norm_b", "orc_loop_seconds"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [P]
        for f in ("orc_history", "orc_alpha_history", "orc_omega_history", "orc_rho_history"):
            getattr(L, f).restype = D
            getattr(L, f).argtypes = [P]
        L.orc_check_solution.argtypes = [P, D, D]
        _lib = L
    return _lib


def _dptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_config(np_=(128, 128, 256), nranks=(1, 1, 1), ds=(0.1, 0.1, 0.1), origin=(0.0, 0.0, 0.0),
                bcs=(0, 1, 0, 1, 0, 1), solver=SOLVER_BICGSTAB, precond=PRECOND_NONE, tolerance=1e-8,
                max_iter=1700, cheb_max=11, cheb_epsilon=1e-4, cheb_rescale_min=500.0,
                cheb_rescale_max=1 - 1e-4, precond_tolerance=1e4 * 1e-10, precond_max_iter=150, order_neumann=2, dim=3,
                cheb_eig_local=0, cheb_f32=0, precond_comm=0) -> OrcConfig:
    c = OrcConfig()
    c.np[:] = list(np_)
    c.nranks[:] = list(nranks)
    c.ds[:] = list(ds)
    c.origin[:] = list(origin)
    c.bcs[:] = list(bcs)
    c.solver, c.precond, c.tolerance, c.max_iter = solver, precond, tolerance, max_iter
    c.cheb_max, c.cheb_epsilon = cheb_max, cheb_epsilon
    c.cheb_rescale_min, c.cheb_rescale_max = cheb_rescale_min, cheb_rescale_max
    c.precond_tolerance, c.precond_max_iter = precond_tolerance, precond_max_iter
    c.order_neumann = order_neumann
    c.dim = dim
    c.cheb_eig_local, c.cheb_f32, c.precond_comm = int(cheb_eig_local), int(cheb_f32), int(precond_comm)
    return c


class Oracle:
    """All px*py*pz ranks of one decomposition, advanced in lock-step on the CPU."""

    def __init__(self, cfg: OrcConfig):
        self.L = lib()
        self.cfg = cfg
        self.h = self.L.orc_create(C.byref(cfg))
        self.world = self.L.orc_world(self.h)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def block(self, rank: int) -> OrcBlockInfo:
        bi = OrcBlockInfo()
        self.L.orc_block(self.h, rank, C.byref(bi))
        return bi

    def shape(self, rank: int):
        bi = self.block(rank)
        return (bi.nguards[2], bi.nguards[1], bi.nguards[0])  # (k, j, i), x fastest

    def _view(self, ptr, rank):
        shp = self.shape(rank)
        return np.ctypeslib.as_array(ptr, shape=shp)

    def x(self, rank: int) -> np.ndarray:
        return self._view(self.L.orc_x(self.h, rank), rank)

    def b(self, rank: int) -> np.ndarray:
        return self._view(self.L.orc_b(self.h, rank), rank)

    def eigenvalues(self, rank=0):
        g = (C.c_double * 2)()
        l = (C.c_double * 2)()
        self.L.orc_eigenvalues(self.h, rank, g, l)
        return tuple(g), tuple(l)

    def zero_fields(self):
        self.L.orc_zero_fields(self.h)

    def set_problem(self):
        self.L.orc_set_problem(self.h)

    def neumann_face(self, rank: int, face: int) -> np.ndarray:
        bi = self.block(rank)
        n = [bi.nlocal[0], bi.nlocal[1], bi.nlocal[2]]
        d = face // 2
        t = [a for a in range(3) if a != d]
        out = np.zeros((n[t[1]], n[t[0]]), dtype=np.float64)
        cnt = self.L.orc_neumann_face(self.h, rank, face, _dptr(out))
        assert cnt == out.size
        return out

    def apply(self, rank: int, field: np.ndarray) -> np.ndarray:
        out = np.zeros_like(field)
        self.L.orc_apply(self.h, rank, _dptr(field), _dptr(out))
        return out

    def halo_exchange(self, fields):
        arr = (C.POINTER(C.c_double) * self.world)(*[_dptr(f) for f in fields])
        self.L.orc_halo_exchange(self.h, arr)

    def reset_neumann(self, rank, field, with_bc_value=False, norm_b=1.0):
        self.L.orc_reset_neumann(self.h, rank, _dptr(field), int(with_bc_value), float(norm_b))

    def adjust_b(self, rank, x, b):
        self.L.orc_adjust_b(self.h, rank, _dptr(x), _dptr(b))

    def precondition(self, X, B):
        ax = (C.POINTER(C.c_double) * self.world)(*[_dptr(f) for f in X])
        ab = (C.POINTER(C.c_double) * self.world)(*[_dptr(f) for f in B])
        self.L.orc_precondition(self.h, ax, ab)

    def solve(self) -> int:
        return self.L.orc_solve(self.h)

    @property
    def iters(self):
        return self.L.orc_iters(self.h)

    @property
    def error_iteration(self):
        return self.L.orc_error_iteration(self.h)

    @property
    def error_operator(self):
        return self.L.orc_error_operator(self.h)

    @property
    def norm_b(self):
        return self.L.orc_norm_b(self.h)

    @property
    def loop_seconds(self):
        return self.L.orc_loop_seconds(self.h)

    def history(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.orc_history(self.h), shape=(self.iters + 1,)).copy()

    def scalar_histories(self):
        n = self.iters
        f = lambda p: np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].copy()
        return f(self.L.orc_alpha_history(self.h)), f(self.L.orc_omega_history(self.h)), f(self.L.orc_rho_history(self.h))

    def check_solution(self):
        s = np.zeros(self.world)
        m = np.zeros(self.world)
        self.L.orc_check_solution(self.h, _dptr(s), _dptr(m))
        return s, m

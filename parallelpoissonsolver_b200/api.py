"""ctypes binding of include/pps_b200.h.  Mirrors the reference's solver-class concept
(T_Solver ctor / setProblem / operator() / getters, solverPoissonMPI_CPU/src/main.cpp:83-126)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

SOLVER_BICGSTAB, SOLVER_CG, SOLVER_CHEBYSHEV = 0, 1, 2
PRECOND_NONE, PRECOND_CHEBYSHEV, PRECOND_BICGSTAB_LOCAL, PRECOND_CG_CHEB_LOCAL = 0, 1, 2, 3
ARITH_FAST, ARITH_PARITY = 0, 1
FUSE_AUTO, FUSE_SPLIT, FUSE_FULL = 0, 1, 2
CHEB_EIG_GLOBAL, CHEB_EIG_LOCAL = 0, 1      # alpaka tree inputParam.hpp:21-22 (`global` / `local`)
CHEB_FP64, CHEB_FP32 = 0, 1                 # alpaka tree solverSetup.hpp:14 (T_data_chebyshev)
FLAG_OPERATOR_ONLY = 1
FLAG_NO_DOT_VECTOR = 2
ABI_VERSION = 1
UNIQUE_ID_BYTES = 128

_HERE = os.path.dirname(os.path.abspath(__file__))


class PpsError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int), ("dim", C.c_int), ("npglobal", C.c_int * 3), ("nranks", C.c_int * 3),
        ("ds", C.c_double * 3), ("origin", C.c_double * 3), ("guards", C.c_int * 3), ("bcs_type", C.c_int * 6),
        ("solver", C.c_int), ("precond", C.c_int), ("tolerance", C.c_double), ("max_iter", C.c_int),
        ("cheb_max_iter", C.c_int), ("cheb_epsilon", C.c_double), ("cheb_rescale_min", C.c_double),
        ("cheb_rescale_max", C.c_double), ("order_neumann", C.c_int), ("arithmetic", C.c_int), ("fusion", C.c_int),
        ("device", C.c_int), ("flags", C.c_int), ("precond_max_iter", C.c_int), ("precond_tolerance", C.c_double),
        ("cheb_eigenvalues", C.c_int),
        ("cheb_precision", C.c_int),
        ("cheb_block", C.c_int),
        ("precond_communication", C.c_int),
    ]


class BlockInfo(C.Structure):
    _fields_ = [
        ("rank", C.c_int), ("global_location", C.c_int * 3), ("nlocal_noguards", C.c_int * 3),
        ("nlocal_guards", C.c_int * 3), ("limits_data", C.c_int * 6), ("limits_solver", C.c_int * 6),
        ("has_boundary", C.c_int * 6), ("has_communication", C.c_int * 6), ("ntot_guards", C.c_longlong),
    ]


def library_path() -> str:
    return os.environ.get("PPS_B200_LIBRARY", os.path.join(_HERE, "csrc", "libpps_b200.so"))


_lib = None


def load_library():
    """Load libpps_b200.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise PpsError(f"{path} is missing: run `python -m parallelpoissonsolver_b200.build` (nvcc, sm_100a). "
                       "There is no CPU fallback.")
    L = C.CDLL(path)
    P = C.c_void_p
    D = C.POINTER(C.c_double)
    L.pps_last_error.restype = C.c_char_p
    L.pps_version.restype = C.c_int
    L.pps_default_config.argtypes = [C.POINTER(Config)]
    L.pps_default_config.restype = None
    L.pps_get_unique_id.argtypes = [C.c_char_p]
    L.pps_create.argtypes = [C.POINTER(Config), C.c_int, C.c_int, C.c_char_p, C.POINTER(P)]
    L.pps_destroy.argtypes = [P]
    L.pps_num_local_blocks.argtypes = [P]
    L.pps_block_info_get.argtypes = [P, C.c_int, C.POINTER(BlockInfo)]
    L.pps_eigenvalues.argtypes = [P, C.c_int, D, D]
    L.pps_set_fields.argtypes = [P, C.c_int, C.c_void_p, C.c_void_p]
    L.pps_set_neumann_face.argtypes = [P, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    L.pps_solve.argtypes = [P]
    L.pps_save_fields.argtypes = [P]
    L.pps_restore_fields.argtypes = [P]
    L.pps_get_solution.argtypes = [P, C.c_int, C.c_void_p]
    L.pps_get_rhs.argtypes = [P, C.c_int, C.c_void_p]
    L.pps_get_iterations.argtypes = [P]
    L.pps_get_preconditioner_iterations.argtypes = [P]
    L.pps_get_preconditioner_iterations.restype = C.c_longlong
    for f in ("pps_get_error_iteration", "pps_get_error_operator", "pps_get_norm_b", "pps_get_solver_seconds",
              "pps_get_loop_seconds"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [P]
    L.pps_get_history.argtypes = [P, C.c_int, D, C.c_int]
    L.pps_check_solution.argtypes = [P, C.c_int, C.c_void_p, D, D]
    L.pps_apply_operator.argtypes = [P, C.c_int, C.c_void_p, C.c_void_p]
    L.pps_apply_preconditioner.argtypes = [P, C.c_int, C.c_void_p, C.c_void_p]
    L.pps_bench_operator.argtypes = [P, C.c_int, C.c_int, D]
    L.pps_set_profiling.argtypes = [P, C.c_int]
    L.pps_get_kernel_stats.argtypes = [P, C.c_int, D, C.POINTER(C.c_longlong), C.POINTER(C.c_char_p)]
    L.pps_get_launch_count.argtypes = [P]
    L.pps_get_launch_count.restype = C.c_longlong
    L.pps_synchronize.argtypes = [P]
    L.pps_set_max_iterations.argtypes = [P, C.c_int]
    L.pps_allgather.argtypes = [P, D, C.c_int, D]
    L.pps_debug_fused.argtypes = [P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.c_int]
    L.pps_debug_peek.argtypes = [P, C.c_int, C.c_longlong, C.c_int, D]
    L.pps_device_count.restype = C.c_int
    _lib = L
    return L


def default_config() -> Config:
    c = Config()
    load_library().pps_default_config(C.byref(c))
    return c


def make_config(npglobal, nranks=(1, 1, 1), ds=(0.1, 0.1, 0.1), origin=(0.0, 0.0, 0.0), bcs=(0, 0, 0, 0, 0, 0),
                solver=SOLVER_BICGSTAB, precond=PRECOND_NONE, tolerance=1e-8, max_iter=1700, cheb_max_iter=11,
                cheb_epsilon=1e-4, cheb_rescale_min=500.0, cheb_rescale_max=1 - 1e-4, arithmetic=ARITH_FAST,
                fusion=FUSE_AUTO, device=-1, flags=0, order_neumann=2, precond_tolerance=1e4 * 1e-10, precond_max_iter=150, dim=3,
                cheb_eigenvalues=CHEB_EIG_GLOBAL, cheb_precision=CHEB_FP64, cheb_block=0, precond_communication=0) -> Config:
    c = Config()
    c.abi_version = ABI_VERSION
    c.dim = int(dim)
    c.npglobal[:] = list(npglobal)
    c.nranks[:] = list(nranks)
    c.ds[:] = [float(v) for v in ds]
    c.origin[:] = [float(v) for v in origin]
    c.guards[:] = [1, 1, 1]
    c.bcs_type[:] = list(bcs)
    c.solver, c.precond = solver, precond
    c.tolerance, c.max_iter = float(tolerance), int(max_iter)
    c.cheb_max_iter, c.cheb_epsilon = int(cheb_max_iter), float(cheb_epsilon)
    c.cheb_rescale_min, c.cheb_rescale_max = float(cheb_rescale_min), float(cheb_rescale_max)
    c.order_neumann = int(order_neumann)
    c.precond_tolerance, c.precond_max_iter = float(precond_tolerance), int(precond_max_iter)
    c.arithmetic, c.fusion, c.device = arithmetic, fusion, device
    c.flags = flags
    c.cheb_eigenvalues, c.cheb_precision, c.cheb_block = int(cheb_eigenvalues), int(cheb_precision), int(cheb_block)
    c.precond_communication = int(precond_communication)
    return c


def get_unique_id() -> bytes:
    L = load_library()
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    if L.pps_get_unique_id(buf):
        raise PpsError(L.pps_last_error().decode())
    return buf.raw


def _ptr(a):
    """host pointer of a numpy array or a (pinned) CPU torch tensor"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need contiguous float64"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.device.type == "cpu" and a.is_contiguous() and a.element_size() == 8
        return a.data_ptr()
    raise TypeError(type(a))


class PoissonSolver:
    """One handle of libpps_b200.so: the T_Solver object of the reference (main.cpp:83), on one GPU."""

    def __init__(self, cfg: Config, rank: int = 0, world_size: int = 1, unique_id: bytes | None = None):
        self.L = load_library()
        self.cfg = cfg
        self.rank, self.world_size = rank, world_size
        self.h = C.c_void_p()
        rc = self.L.pps_create(C.byref(cfg), rank, world_size, unique_id, C.byref(self.h))
        if rc:
            self.h = None
            raise PpsError(self.L.pps_last_error().decode())
        nr = cfg.nranks[0] * cfg.nranks[1] * cfg.nranks[2]
        self.local_ranks = list(range(nr)) if world_size == 1 else [rank]

    def _ck(self, rc):
        if rc:
            raise PpsError(self.L.pps_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.pps_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # --- BlockGrid getters (blockGrid.hpp:40-145)
    def block(self, rank: int) -> BlockInfo:
        bi = BlockInfo()
        self._ck(self.L.pps_block_info_get(self.h, rank, C.byref(bi)))
        return bi

    def shape(self, rank: int):
        bi = self.block(rank)
        return (bi.nlocal_guards[2], bi.nlocal_guards[1], bi.nlocal_guards[0])

    def eigenvalues(self, rank: int = 0):
        g = (C.c_double * 2)()
        l = (C.c_double * 2)()
        self._ck(self.L.pps_eigenvalues(self.h, rank, g, l))
        return tuple(g), tuple(l)

    # --- problem hand-over (main.cpp:94-99)
    def set_fields(self, rank: int, x, b):
        self._ck(self.L.pps_set_fields(self.h, rank, _ptr(x), _ptr(b)))

    def set_neumann_face(self, rank: int, face: int, dudn: np.ndarray):
        dudn = np.ascontiguousarray(dudn, dtype=np.float64)
        self._ck(self.L.pps_set_neumann_face(self.h, rank, face, dudn.ctypes.data, dudn.size))

    def solve(self):
        self._ck(self.L.pps_solve(self.h))
        return self.iterations

    def save_fields(self):
        self._ck(self.L.pps_save_fields(self.h))

    def restore_fields(self):
        self._ck(self.L.pps_restore_fields(self.h))

    def get_solution(self, rank: int, out=None):
        if out is None:
            out = np.empty(self.shape(rank), dtype=np.float64)
        self._ck(self.L.pps_get_solution(self.h, rank, _ptr(out)))
        return out

    def get_rhs(self, rank: int, out=None):
        if out is None:
            out = np.empty(self.shape(rank), dtype=np.float64)
        self._ck(self.L.pps_get_rhs(self.h, rank, _ptr(out)))
        return out

    # --- getters (iterativeSolverBase.hpp:410-425)
    @property
    def iterations(self) -> int:
        return self.L.pps_get_iterations(self.h)

    @property
    def preconditioner_iterations(self) -> int:
        """iterations of a nested Krylov preconditioner during the last solve (summed over calls and local blocks)"""
        return self.L.pps_get_preconditioner_iterations(self.h)

    @property
    def error_iteration(self) -> float:
        return self.L.pps_get_error_iteration(self.h)

    @property
    def error_operator(self) -> float:
        return self.L.pps_get_error_operator(self.h)

    @property
    def norm_b(self) -> float:
        return self.L.pps_get_norm_b(self.h)

    @property
    def solver_seconds(self) -> float:
        return self.L.pps_get_solver_seconds(self.h)

    @property
    def loop_seconds(self) -> float:
        return self.L.pps_get_loop_seconds(self.h)

    def history(self, which: int = 0) -> np.ndarray:
        n = self.iterations + (1 if which == 0 else 0)
        out = np.zeros(max(n, 1), dtype=np.float64)
        self._ck(self.L.pps_get_history(self.h, which, out.ctypes.data_as(C.POINTER(C.c_double)), out.size))
        return out[:n]

    def check_solution(self, rank: int, u_exact: np.ndarray):
        s, m = C.c_double(), C.c_double()
        self._ck(self.L.pps_check_solution(self.h, rank, _ptr(u_exact), C.byref(s), C.byref(m)))
        return s.value, m.value

    # --- building blocks
    def apply_operator(self, rank: int, field: np.ndarray) -> np.ndarray:
        out = np.empty_like(field)
        self._ck(self.L.pps_apply_operator(self.h, rank, _ptr(field), _ptr(out)))
        return out

    def apply_preconditioner(self, rank: int, field: np.ndarray) -> np.ndarray:
        out = np.empty_like(field)
        self._ck(self.L.pps_apply_preconditioner(self.h, rank, _ptr(field), _ptr(out)))
        return out

    def bench_operator(self, reps: int = 20, with_dot: bool = False, with_halo: bool = False) -> float:
        ms = C.c_double()
        self._ck(self.L.pps_bench_operator(self.h, reps, int(with_dot) | (2 if with_halo else 0), C.byref(ms)))
        return ms.value

    def set_profiling(self, on: bool):
        self.L.pps_set_profiling(self.h, int(on))

    def kernel_stats(self):
        out = []
        k = 0
        while True:
            ms, n, name = C.c_double(), C.c_longlong(), C.c_char_p()
            if self.L.pps_get_kernel_stats(self.h, k, C.byref(ms), C.byref(n), C.byref(name)):
                break
            if n.value:
                out.append(dict(id=k, name=name.value.decode(), avg_ms=ms.value, launches=n.value))
            k += 1
        return out

    @property
    def launch_count(self) -> int:
        return self.L.pps_get_launch_count(self.h)

    def debug_fused(self, which: int, variant: int, reps: int):
        """see pps_debug_fused (include/pps_b200.h)"""
        out = (C.c_longlong * 32)()
        self._ck(self.L.pps_debug_fused(self.h, which, variant, reps, out, 32))
        return list(out)

    def debug_peek(self, array: int, offset: int, n: int) -> np.ndarray:
        out = np.zeros(n)
        self._ck(self.L.pps_debug_peek(self.h, array, offset, n, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def synchronize(self):
        self._ck(self.L.pps_synchronize(self.h))

    def set_max_iterations(self, n: int):
        self._ck(self.L.pps_set_max_iterations(self.h, int(n)))

    def allgather(self, values) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = np.zeros(v.size * self.world_size)
        D = C.POINTER(C.c_double)
        self._ck(self.L.pps_allgather(self.h, v.ctypes.data_as(D), v.size, out.ctypes.data_as(D)))
        return out.reshape(self.world_size, v.size)

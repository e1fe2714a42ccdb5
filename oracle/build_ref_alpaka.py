#!/usr/bin/env python3
"""Build the UNMODIFIED alpaka tree of the reference (solverPoissonMPI_alpaka) with alpaka's OpenMP-blocks CPU
accelerator, as the pin of the alpaka-only configuration surface (mixed-precision and
local-eigenvalue Chebyshev).  TEST INFRASTRUCTURE ONLY -- nothing in the product path uses this.

What makes it buildable here without the reference's build system (cmake + Boost + MPI + OpenMP):
  * alpaka 1.2.0 and mdspan are vendored by the reference itself (thirdParty/alpaka, thirdParty/alpaka/_deps/mdspan-src)
    and are header-only; they are used where they lie;
  * alpaka needs Boost only for compiler detection macros and type-name demangling: oracle/boost_shim/ supplies those
    three headers for g++ / Linux; `-DALPAKA_DISABLE_ATOMIC_ATOMICREF` selects alpaka's own lock-based atomics -- the
    configuration alpaka's cmake falls back to when neither std::atomic_ref (C++20) nor Boost.Atomic is available
    (thirdParty/alpaka/cmake/alpakaCommon.cmake:251-253);
  * MPI is the threads-as-ranks oracle/mpi_shim/ (plus its MPI-IO subset for src/main.cpp:137-145);
  * `-DALPAKA_ACC_CPU_B_OMP2_T_SEQ_ENABLED -fopenmp` is what env/lumi/CMakeListsCPU.txt:34 asks alpaka's cmake for and
    the accelerator solverSetup.hpp:21 names; `-DALPAKA_USE_MDSPAN` is env/lumi/CMakeListsCPU.txt:41.
Per configuration the 15 headers of solverPoissonMPI_alpaka/include are copied into oracle/_ref/cfg_alpaka/<name>/
(git-ignored, never committed) and ONLY constants / typedefs of inputParam.hpp and solverSetup.hpp are edited there.
Outputs: oracle/_ref/bin/alp_solver_<name> (the tree's own main.cpp) and oracle/_ref/bin/alp_dump_<name>
(oracle/ref_dump_alpaka.cpp).  Run the binaries with OMP_NUM_THREADS=1: the kernels' atomicAdd reductions are then
summed in block order and the outputs are reproducible.
"""
from __future__ import annotations

import os
import shutil
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.build_ref import OUT, REF_ROOT, _fmt, _run, _sub  # noqa: E402

# -fopenmp needs the compiler driver that knows libgomp.spec: this image's $CXX wrapper (/opt/gcc/bin/g++) does not, /usr/bin/g++ does
CXX = os.environ.get("PPS_ALPAKA_CXX") or ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else os.environ.get("CXX", "g++"))

REF_ALP = os.path.join(REF_ROOT, "solverPoissonMPI_alpaka")

_MAIN = "BiCGstabAlpaka<DIM, T_data, T_data, tollMainSolver, iterMaxMainSolver, isbiCGMainLoop1, communicationON, %s>"
SOLVER_TYPEDEFS = {
    "bicgstab_none": _MAIN % "T_NoneSolverAlpaka",
    "bicgstab_chebglobal": _MAIN % "T_PreconditionerChebGlobal",   # the tree as shipped, inputParam.hpp:36
    "bicgstab_cheblocal": _MAIN % "T_PreconditionerChebLocal",     # `local`, inputParam.hpp:27
    "bicgstab_bicgloc": _MAIN % "T_PreconditionerBiCGStabLocal",   # nested block-local BiCGSTAB, inputParam.hpp:31
    "bicgstab_bicgglob": _MAIN % "T_PreconditionerBiCGStabGlobal",  # nested GLOBAL BiCGSTAB (communicationON), inputParam.hpp:33
}
PRECOND_TYPEDEF = {"bicgstab_none": "T_NoneSolverAlpaka", "bicgstab_chebglobal": "T_PreconditionerChebGlobal",
                   "bicgstab_cheblocal": "T_PreconditionerChebLocal", "bicgstab_bicgloc": "T_PreconditionerBiCGStabLocal",
                   "bicgstab_bicgglob": "T_PreconditionerBiCGStabGlobal"}

DIRICHLET = (0, 0, 0, 0, 0, 0)
MIXED = (0, 1, 0, 1, 0, 1)
M24 = dict(np=(24, 20, 28), bcs=MIXED, ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1))


def cfg(np, bcs=DIRICHLET, solver="bicgstab_chebglobal", cheb_type="double", ds=(0.1, 0.1, 0.1), origin=(0, 0, 0),
        toll_scaling=1e-8, toll_main=1, iter_max=1700, cheb_max=11, rescale_min=500.0, rescale_max=1 - 1e-4, write_files=False,
        toll_precond=None, precond_iter_max=None):
    # (toll_main stays 1 unless toll_precond is given: tollPreconditionerSolver = tollMainSolver * 1e8 must fit an int, solverSetup.hpp:63)
    return dict(np=tuple(np), bcs=tuple(bcs), solver=solver, cheb_type=cheb_type, ds=tuple(ds), origin=tuple(origin),
                toll_scaling=toll_scaling, toll_main=toll_main, iter_max=iter_max, cheb_max=cheb_max,
                rescale_min=rescale_min, rescale_max=rescale_max, write_files=write_files, toll_precond=toll_precond,
                precond_iter_max=precond_iter_max)


CONFIGS = {
    # the alpaka tree exactly as shipped except for the grid size (64^3, mixed BCs {0,1,1,0,1,0}, Chebyshev(24), rescaleEigMin 100)
    "alp_shipped32": cfg((32, 32, 32), (0, 1, 1, 0, 1, 0), toll_scaling=1e-10, iter_max=1000, cheb_max=24, rescale_min=100.0),
    # T_data_chebyshev = float (solverSetup.hpp:14)
    "alp_f32_d24": cfg((24, 24, 24), cheb_type="float"),
    "alp_f32_m24": cfg(cheb_type="float", **M24),
    "alp_f32_m24_c24": cfg(cheb_type="float", cheb_max=24, rescale_min=100.0, **M24),
    # `local` eigenvalue bounds (inputParam.hpp:21-22,27), fp64 and fp32 iterates
    "alp_loc_m24": cfg(solver="bicgstab_cheblocal", **M24),
    "alp_f32loc_m24": cfg(solver="bicgstab_cheblocal", cheb_type="float", **M24),
    "alp_f32loc_d32": cfg((32, 32, 32), solver="bicgstab_cheblocal", cheb_type="float"),
    # fp64 global: the alpaka kernels' folded 7-point form against the CPU tree's expression order
    "alp_f64_m24": cfg(**M24),
    "alp_none_m24": cfg(solver="bicgstab_none", **M24),
    # nested Krylov preconditioners of the alpaka tree (inputParam.hpp:31,33) with the CPU tree's settings: main tolerance 100 * 1e-10,
    # nested tolerance 1e4 * 1e-10, nested iteration cap 150 (or 8: the nested solves then stop on the cap)
    "alp_nbl_m24": cfg(solver="bicgstab_bicgloc", toll_scaling=1e-10, toll_main=100, toll_precond=10000, precond_iter_max=150, **M24),
    "alp_nbg_m24": cfg(solver="bicgstab_bicgglob", toll_scaling=1e-10, toll_main=100, toll_precond=10000, precond_iter_max=150, **M24),
    "alp_nbg_m24_i8": cfg(solver="bicgstab_bicgglob", toll_scaling=1e-10, toll_main=100, toll_precond=10000, precond_iter_max=8, **M24),
    # writeResidual / writeSolution = true (inputParam.hpp:46-47): residualHistory.txt and solution.dat from the tree's own main.cpp
    "alp_files_m24": cfg(write_files=True, **M24),
}

CXXFLAGS = ["-std=c++17", "-O3", "-DNDEBUG", "-pthread", "-w", "-fopenmp",
            "-DALPAKA_ACC_CPU_B_OMP2_T_SEQ_ENABLED", "-DALPAKA_USE_MDSPAN", "-DALPAKA_DISABLE_ATOMIC_ATOMICREF"]


def make_cfg_dir(name, c):
    src = os.path.join(REF_ALP, "include")
    dst = os.path.join(OUT, "cfg_alpaka", name)
    os.makedirs(dst, exist_ok=True)
    for f in os.listdir(src):
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    p = os.path.join(dst, "inputParam.hpp")
    t = open(p).read()
    t = _sub(t, r"^using T_Solver = .*;$", "using T_Solver = " + SOLVER_TYPEDEFS[c["solver"]] + ";", p)
    t = _sub(t, r"npglobal=\{[^}]*\}", "npglobal={%s}" % ",".join(map(str, c["np"])), p)
    t = _sub(t, r"> ds=\{[^}]*\}", "> ds={%s}" % ",".join(_fmt(float(v)) for v in c["ds"]), p)
    t = _sub(t, r"origin=\{[^}]*\}", "origin={%s}" % ",".join(_fmt(float(v)) for v in c["origin"]), p)
    t = _sub(t, r"bcsType=\{[^}]*\}", "bcsType={%s}" % ",".join(map(str, c["bcs"])), p)
    if c.get("write_files"):
        t = _sub(t, r"writeResidual = false;", "writeResidual = true;", p)
        t = _sub(t, r"writeSolution = false;", "writeSolution = true;", p)
    open(p, "w").write(t)
    p = os.path.join(dst, "solverSetup.hpp")
    t = open(p).read()
    t = _sub(t, r"using T_data_chebyshev=\w+;", "using T_data_chebyshev=%s;" % c["cheb_type"], p)
    t = _sub(t, r"tollScalingFactor = [^;]*;", "tollScalingFactor = %s;" % _fmt(float(c["toll_scaling"])), p)
    t = _sub(t, r"tollMainSolver=[^;]*;", "tollMainSolver=%d;" % c["toll_main"], p)
    t = _sub(t, r"iterMaxMainSolver=[^;]*;", "iterMaxMainSolver=%d;" % c["iter_max"], p)
    t = _sub(t, r"chebyshevMax=[^;]*;", "chebyshevMax=%d;" % c["cheb_max"], p)
    t = _sub(t, r"rescaleEigMin= [^;]*;", "rescaleEigMin= %s;" % _fmt(float(c["rescale_min"])), p)
    t = _sub(t, r"rescaleEigMax= [^;]*;", "rescaleEigMax= %s;" % _fmt(float(c["rescale_max"])), p)
    if c.get("toll_precond") is not None:
        t = _sub(t, r"tollPreconditionerSolver=[^;]*;", "tollPreconditionerSolver=%d;" % c["toll_precond"], p)
    if c.get("precond_iter_max") is not None:
        t = _sub(t, r"iterMaxPreconditioner=[^;]*;", "iterMaxPreconditioner=%d;" % c["precond_iter_max"], p)
    open(p, "w").write(t)
    return dst


def build_one(name, shim_obj, force=False):
    c = CONFIGS[name]
    bindir = os.path.join(OUT, "bin")
    os.makedirs(bindir, exist_ok=True)
    solver_bin = os.path.join(bindir, "alp_solver_" + name)
    dump_bin = os.path.join(bindir, "alp_dump_" + name)
    dump_src = os.path.join(HERE, "ref_dump_alpaka.cpp")
    fresh = all(os.path.exists(b) and os.path.getmtime(b) >= max(os.path.getmtime(dump_src), os.path.getmtime(shim_obj),
                                                                  os.path.getmtime(__file__)) for b in (solver_bin, dump_bin))
    if not force and fresh:
        return solver_bin, dump_bin
    cfgdir = make_cfg_dir(name, c)
    alp = os.path.join(REF_ALP, "thirdParty", "alpaka")
    inc = ["-I" + os.path.join(HERE, "boost_shim"), "-I" + os.path.join(HERE, "mpi_shim"), "-I" + cfgdir,
           "-I" + os.path.join(alp, "include"), "-I" + os.path.join(alp, "_deps", "mdspan-src", "include")]
    obj = os.path.join(cfgdir, "alp_main.o")
    # main.cpp includes its headers by name: the per-config copies are found through -I<cfgdir>
    _run([CXX] + CXXFLAGS + inc + ["-Dmain=ref_main", "-c", os.path.join(REF_ALP, "src", "main.cpp"), "-o", obj])
    _run([CXX] + CXXFLAGS + inc + [obj, os.path.join(HERE, "ref_launcher.cpp"), shim_obj, "-o", solver_bin])
    _run([CXX] + CXXFLAGS + inc + ["-DPPS_ALP_PRECOND=" + PRECOND_TYPEDEF[c["solver"]], dump_src, shim_obj, "-o", dump_bin])
    return solver_bin, dump_bin


def build(names=None, force=False, jobs=None):
    """Build the listed configurations (default: all).  Returns {name: (solver_bin, dump_bin)}."""
    if not available():
        raise FileNotFoundError(REF_ALP + " not present (the reference only exists in the build container)")
    os.makedirs(OUT, exist_ok=True)
    shim_obj = os.path.join(OUT, "mpi_shim.o")
    shim_src = os.path.join(HERE, "mpi_shim", "mpi_shim.cpp")
    if force or not os.path.exists(shim_obj) or os.path.getmtime(shim_obj) < os.path.getmtime(shim_src):
        _run([CXX, "-std=c++17", "-O3", "-DNDEBUG", "-pthread", "-w", "-I" + os.path.join(HERE, "mpi_shim"), "-c", shim_src, "-o", shim_obj])
    names = list(names or CONFIGS)
    with ThreadPoolExecutor(max_workers=jobs or min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(lambda n: build_one(n, shim_obj, force), names))
    return dict(zip(names, res))


def available():
    return os.path.isdir(os.path.join(REF_ALP, "thirdParty", "alpaka", "include", "alpaka"))


if __name__ == "__main__":
    force = "--force" in sys.argv
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or None
    for n, (s, d) in build(names, force=force).items():
        print(n, s, d)

#!/usr/bin/env python3
"""Build the UNMODIFIED reference CPU solver (solverPoissonMPI_CPU) as the parity oracle
and CPU baseline.  TEST INFRASTRUCTURE ONLY -- nothing in the product path uses this.

The reference is configured at compile time (README.md:36), so one binary is built per
configuration.  For each entry of CONFIGS this script

  1. copies the 11 headers of /root/reference/solverPoissonMPI_CPU/include into
     oracle/_ref/cfg/<name>/ (git-ignored build output, never committed) and edits ONLY
     constants / the T_Solver typedef there (inputParam.hpp:33,41-45, solverSetup.hpp:22-40)
     -- this is the reference's own configuration mechanism;
  2. compiles solverPoissonMPI_CPU/src/main.cpp where it lies, with -Dmain=ref_main, against
     oracle/mpi_shim/mpi.h (threads-as-ranks; this image has no MPI)  -> _ref/bin/ref_solver_<name>
  3. compiles oracle/ref_dump.cpp (golden-data extractor around the same headers)
                                                                    -> _ref/bin/ref_dump_<name>

Flags follow the reference's CMakeLists.txt:15-31 (-O3 -DNDEBUG, C++17); no -march so the
binaries also run on the GPU box's host CPU.  The reference's own build system is not used.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("PPS_REFERENCE_ROOT", "/root/reference")
REF_CPU = os.path.join(REF_ROOT, "solverPoissonMPI_CPU")
OUT = os.path.join(HERE, "_ref")

SOLVER_TYPEDEFS = {
    # the commented alternative at inputParam.hpp:32
    "bicgstab_none": "BiCGSTAB<DIM, T_data, tollMainSolver, iterMaxMainSolver, isbiCGMainLoop1, communicationON, T_NoneSolver>",
    # the shipped default, inputParam.hpp:33
    "bicgstab_cheb": "BiCGSTAB<DIM, T_data, tollMainSolver, iterMaxMainSolver, isbiCGMainLoop1, communicationON, T_Preconditioner2>",
    "cg_none": "BaseCG<DIM, T_data, tollMainSolver, iterMaxMainSolver, isbiCGMainLoop1, communicationON, T_NoneSolver>",
    "cg_cheb": "BaseCG<DIM, T_data, tollMainSolver, iterMaxMainSolver, isbiCGMainLoop1, communicationON, T_Preconditioner2>",
    # nested Krylov preconditioners: the local BiCGSTAB of inputParam.hpp:31 and the local CG with
    # Chebyshev inside of inputParam.hpp:29 in the preconditioner slot
    "bicgstab_bicgloc": "BiCGSTAB<DIM, T_data, tollMainSolver, iterMaxMainSolver, isbiCGMainLoop1, communicationON, T_Preconditioner>",
    "bicgstab_cgcheb": "BiCGSTAB<DIM, T_data, tollMainSolver, iterMaxMainSolver, isbiCGMainLoop1, communicationON, T_Preconditioner3>",
    # GLOBAL Chebyshev preconditioner: the same class with communicationON in the preconditioner slot (chebyshevIteration.hpp:69-73,97-101)
    "bicgstab_chebglobal": "BiCGSTAB<DIM, T_data, tollMainSolver, iterMaxMainSolver, isbiCGMainLoop1, communicationON, "
                           "ChebyshevIteration<DIM, T_data, tollPreconditionerSolver, chebyshevMax, ischebyshevMainLoop, communicationON, T_NoneSolver>>",
    # GLOBAL nested BiCGSTAB preconditioner: the class of inputParam.hpp:31 with communicationON (face exchanges and allreduces inside the
    # preconditioner; the alpaka tree names it T_PreconditionerBiCGStabGlobal, solverPoissonMPI_alpaka/include/inputParam.hpp:33)
    "bicgstab_bicgglob": "BiCGSTAB<DIM, T_data, tollMainSolver, iterMaxMainSolver, isbiCGMainLoop1, communicationON, "
                         "BiCGSTAB<DIM, T_data, tollPreconditionerSolver, iterMaxPreconditioner, isbiCGMainLoop2, communicationON, T_NoneSolver>>",
    # Chebyshev iteration as the MAIN solver (chebyshevIteration.hpp:61-67,132-139): isMainLoop = true, communicationON
    "cheb_main": "ChebyshevIteration<DIM, T_data, tollMainSolver, chebyshevMax, true, communicationON, T_NoneSolver>",
}

DIRICHLET = (0, 0, 0, 0, 0, 0)
MIXED = (0, 1, 0, 1, 0, 1)  # shipped default


def cfg(np, bcs=DIRICHLET, solver="bicgstab_none", ds=(0.1, 0.1, 0.1), origin=(0, 0, 0), toll_scaling=1e-10,
        toll_main=100, iter_max=1700, cheb_max=11, order_neumann=2, rescale_min=None, rescale_max=None, dim=3, precond_iter_max=None):
    return dict(np=tuple(np), bcs=tuple(bcs), solver=solver, ds=tuple(ds), origin=tuple(origin),
                toll_scaling=toll_scaling, toll_main=toll_main, iter_max=iter_max, cheb_max=cheb_max,
                order_neumann=order_neumann, rescale_min=rescale_min, rescale_max=rescale_max, dim=dim, precond_iter_max=precond_iter_max)


# name -> configuration.  "default" is the reference exactly as shipped.
CONFIGS = {
    "default": cfg((128, 128, 256), MIXED, "bicgstab_cheb"),
    # small golden cases (tests/golden/): every BC family, both preconditioner settings, CG
    "d16": cfg((16, 16, 16)),
    "d32": cfg((32, 32, 32)),
    "d64": cfg((64, 64, 64)),
    "d64_t12": cfg((64, 64, 64), toll_scaling=1e-14),
    "d32_cheb": cfg((32, 32, 32), solver="bicgstab_cheb"),
    "d64_cheb": cfg((64, 64, 64), solver="bicgstab_cheb"),
    "m24": cfg((24, 20, 28), MIXED, "bicgstab_none", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1)),
    "m24_cheb": cfg((24, 20, 28), MIXED, "bicgstab_cheb", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1)),
    "m32_cheb": cfg((32, 32, 64), MIXED, "bicgstab_cheb"),
    "n24": cfg((24, 24, 24), (1, 0, 0, 1, 1, 1), "bicgstab_none"),
    "cg32": cfg((32, 32, 32), solver="cg_none"),
    "cg32_cheb": cfg((32, 32, 32), solver="cg_cheb"),
    "cgm24": cfg((24, 20, 28), MIXED, "cg_none", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1)),
    "nb24": cfg((24, 20, 28), MIXED, "bicgstab_bicgloc", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1)),
    "nc24": cfg((24, 20, 28), MIXED, "bicgstab_cgcheb", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1)),
    "nbg24": cfg((24, 20, 28), MIXED, "bicgstab_bicgglob", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1)),
    "nbgd32": cfg((32, 32, 32), DIRICHLET, "bicgstab_bicgglob"),
    "nbg24_i8": cfg((24, 20, 28), MIXED, "bicgstab_bicgglob", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1), precond_iter_max=8),
    # first-order Neumann closure (solverSetup.hpp:25 orderNeumanBcs = 1) and Chebyshev as main solver
    "o1m24": cfg((24, 20, 28), MIXED, "bicgstab_none", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1), order_neumann=1),
    "o1m24_cheb": cfg((24, 20, 28), MIXED, "bicgstab_cheb", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1), order_neumann=1),
    "o1cgm24": cfg((24, 20, 28), MIXED, "cg_none", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1), order_neumann=1),
    "chm24": cfg((24, 20, 28), MIXED, "cheb_main", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1), cheb_max=60,
                 rescale_min=1.0, rescale_max=1.0),
    "chd32": cfg((32, 32, 32), DIRICHLET, "cheb_main", cheb_max=25),
    # DIM = 2 and DIM = 1 (inputParam.hpp:16; argv is then `px py` / `px`, main.cpp:40-48)
    "q24": cfg((24, 20, 1), MIXED, "bicgstab_none", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1), dim=2),
    "q24_cheb": cfg((24, 20, 1), MIXED, "bicgstab_cheb", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1), dim=2),
    "qd40": cfg((40, 36, 1), DIRICHLET, "bicgstab_none", dim=2),
    "qcg40": cfg((40, 36, 1), DIRICHLET, "cg_cheb", dim=2),
    "l48": cfg((48, 1, 1), MIXED, "bicgstab_none", ds=(0.05, 0.1, 0.1), origin=(0.3, -0.2, 0.1), dim=1),
    "l48_cheb": cfg((48, 1, 1), (1, 0, 0, 0, 0, 0), "bicgstab_cheb", ds=(0.05, 0.1, 0.1), origin=(0.3, -0.2, 0.1), dim=1),
    # edge cases: a grid the rank grid does not divide (blockGrid.hpp:165 truncates nlocal = npglobal / nranks: the run silently covers
    # 24 x 20 x 28 of the declared 25 x 21 x 29 points while the global eigenvalue bounds keep using the declared sizes, :322-339) and the
    # smallest blocks the solver accepts (3 points per axis, 2 of them unknowns next to a Dirichlet face)
    "e25": cfg((25, 21, 29), MIXED, "bicgstab_none", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1)),
    "e25_cheb": cfg((25, 21, 29), MIXED, "bicgstab_cheb", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1)),
    "tiny6": cfg((6, 6, 6)),
    "tiny6_cheb": cfg((6, 6, 6), solver="bicgstab_cheb"),
    "d128": cfg((128, 128, 128)),
    "d32_chebg": cfg((32, 32, 32), solver="bicgstab_chebglobal"),
    "m24_chebg": cfg((24, 20, 28), MIXED, "bicgstab_chebglobal", ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1)),
    # CPU-baseline samples for bench.py --impl reference (bounded: fixed iteration count)
    "bench256": cfg((256, 256, 256), iter_max=10000),
    "bench512_it8": cfg((512, 512, 512), iter_max=8),
    "bench1024_it2": cfg((1024, 1024, 1024), iter_max=2),
    # lock-step goldens of the benchmarked configurations (BASELINE.json configs[1], [2]): the first 20 iterations
    "bench512_it20": cfg((512, 512, 512), iter_max=20),
    "bench1024_it20": cfg((1024, 1024, 1024), iter_max=20),
}

CXX = os.environ.get("CXX", "g++")
CXXFLAGS = ["-std=c++17", "-O3", "-DNDEBUG", "-pthread", "-w"]


def _fmt(v):
    return repr(float(v)) if isinstance(v, float) else str(v)


def _sub(text, pattern, repl, path, count=1):
    new, n = re.subn(pattern, repl, text, count=count, flags=re.M)
    if n != count:
        raise RuntimeError(f"pattern {pattern!r} matched {n} times in {path}")
    return new


def make_cfg_dir(name, c):
    src = os.path.join(REF_CPU, "include")
    dst = os.path.join(OUT, "cfg", name)
    os.makedirs(dst, exist_ok=True)
    for f in os.listdir(src):
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    p = os.path.join(dst, "inputParam.hpp")
    t = open(p).read()
    t = _sub(t, r"^using T_Solver = .*;$", "using T_Solver = " + SOLVER_TYPEDEFS[c["solver"]] + ";", p)
    t = _sub(t, r"^constexpr int DIM=\d;", "constexpr int DIM=%d;" % c.get("dim", 3), p)
    t = _sub(t, r"npglobal=\{[^}]*\}", "npglobal={%s}" % ",".join(map(str, c["np"])), p)
    t = _sub(t, r"> ds=\{[^}]*\}", "> ds={%s}" % ",".join(_fmt(float(v)) for v in c["ds"]), p)
    t = _sub(t, r"origin=\{[^}]*\}", "origin={%s}" % ",".join(_fmt(float(v)) for v in c["origin"]), p)
    t = _sub(t, r"bcsType=\{[^}]*\}", "bcsType={%s}" % ",".join(map(str, c["bcs"])), p)
    open(p, "w").write(t)
    p = os.path.join(dst, "solverSetup.hpp")
    t = open(p).read()
    t = _sub(t, r"tollScalingFactor = [^;]*;", "tollScalingFactor = %s;" % _fmt(float(c["toll_scaling"])), p)
    t = _sub(t, r"tollMainSolver=[^;]*;", "tollMainSolver=%d;" % c["toll_main"], p)
    t = _sub(t, r"iterMaxMainSolver=[^;]*;", "iterMaxMainSolver=%d;" % c["iter_max"], p)
    t = _sub(t, r"chebyshevMax=[^;]*;", "chebyshevMax=%d;" % c["cheb_max"], p)
    t = _sub(t, r"orderNeumanBcs=[^;]*;", "orderNeumanBcs=%d;" % c.get("order_neumann", 2), p)
    if c.get("precond_iter_max") is not None:
        t = _sub(t, r"iterMaxPreconditioner=[^;]*;", "iterMaxPreconditioner=%d;" % c["precond_iter_max"], p)
    if c.get("rescale_min") is not None:
        t = _sub(t, r"rescaleEigMin= [^;]*;", "rescaleEigMin= %s;" % _fmt(float(c["rescale_min"])), p)
    if c.get("rescale_max") is not None:
        t = _sub(t, r"rescaleEigMax= [^;]*;", "rescaleEigMax= %s;" % _fmt(float(c["rescale_max"])), p)
    open(p, "w").write(t)
    return dst


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))


def build_dropin(name, force=False):
    """The drop-in demonstration: the reference's own main.cpp + its own inputParam.hpp / solverSetup.hpp (config
    <name>), compiled UNMODIFIED against include/reference_compat (GPU-backed solver classes) and linked with
    libpps_b200.so  -> _ref/bin/ref_main_on_b200_<name>.  Returns "" when the library has not been built."""
    root = os.path.dirname(HERE)
    lib_dir = os.path.join(root, "parallelpoissonsolver_b200", "csrc")
    if not os.path.exists(os.path.join(lib_dir, "libpps_b200.so")):
        return ""
    exe = os.path.join(OUT, "bin", "ref_main_on_b200_" + name)
    compat = os.path.join(root, "include", "reference_compat")
    deps = [os.path.join(compat, f) for f in os.listdir(compat)] + [os.path.join(root, "include", "pps_b200.h"),
            os.path.join(root, "parallelpoissonsolver_b200", "driver", "ref_main_launcher.cpp")]
    if not force and os.path.exists(exe) and all(os.path.getmtime(d) <= os.path.getmtime(exe) for d in deps):
        return exe
    cfgdir = make_cfg_dir(name, CONFIGS[name])
    only = os.path.join(cfgdir, "config_only")
    os.makedirs(only, exist_ok=True)
    for f in ("inputParam.hpp", "solverSetup.hpp"):
        shutil.copyfile(os.path.join(cfgdir, f), os.path.join(only, f))
    inc = ["-I" + os.path.join(root, "include", "reference_compat"), "-I" + os.path.join(root, "include"), "-I" + only]
    obj = os.path.join(only, "ref_main_b200.o")
    _run([CXX] + CXXFLAGS + inc + ["-Dmain=ref_main", "-c", os.path.join(REF_CPU, "src", "main.cpp"), "-o", obj])
    _run([CXX] + CXXFLAGS + inc + [obj, os.path.join(root, "parallelpoissonsolver_b200", "driver", "ref_main_launcher.cpp"), "-o", exe,
                                   "-L" + lib_dir, "-lpps_b200", "-Wl,-rpath," + lib_dir, "-Wl,-rpath,$ORIGIN/../../../parallelpoissonsolver_b200/csrc"])
    return exe


DROPIN_CONFIGS = ["default", "d64", "m24_cheb", "cg32", "nb24", "nbg24_i8", "nc24", "chm24", "o1cgm24", "q24_cheb", "l48", "m24_chebg"]


def build_one(name, shim_obj, force=False):
    c = CONFIGS[name]
    bindir = os.path.join(OUT, "bin")
    os.makedirs(bindir, exist_ok=True)
    solver_bin = os.path.join(bindir, "ref_solver_" + name)
    dump_bin = os.path.join(bindir, "ref_dump_" + name)
    if not force and os.path.exists(solver_bin) and os.path.exists(dump_bin):
        return solver_bin, dump_bin
    cfgdir = make_cfg_dir(name, c)
    inc = ["-I" + os.path.join(HERE, "mpi_shim"), "-I" + cfgdir]
    obj = os.path.join(cfgdir, "ref_main.o")
    _run([CXX] + CXXFLAGS + inc + ["-Dmain=ref_main", "-c", os.path.join(REF_CPU, "src", "main.cpp"), "-o", obj])
    _run([CXX] + CXXFLAGS + inc + [obj, os.path.join(HERE, "ref_launcher.cpp"), shim_obj, "-o", solver_bin])
    dump_defs = ["-DPPS_DUMP_SKIP_NORM"] if c["solver"] == "cheb_main" else []
    _run([CXX] + CXXFLAGS + dump_defs + inc + [os.path.join(HERE, "ref_dump.cpp"), shim_obj, "-o", dump_bin])
    return solver_bin, dump_bin


def build(names=None, force=False, jobs=None):
    """Build the listed configurations (default: all).  Returns {name: (solver_bin, dump_bin)}."""
    if not os.path.isdir(REF_CPU):
        raise FileNotFoundError(REF_CPU + " not present (the reference only exists in the build container)")
    os.makedirs(OUT, exist_ok=True)
    shim_obj = os.path.join(OUT, "mpi_shim.o")
    shim_src = os.path.join(HERE, "mpi_shim", "mpi_shim.cpp")
    if force or not os.path.exists(shim_obj) or os.path.getmtime(shim_obj) < os.path.getmtime(shim_src):
        _run([CXX] + CXXFLAGS + ["-I" + os.path.join(HERE, "mpi_shim"), "-c", shim_src, "-o", shim_obj])
        force = True
    names = list(names or CONFIGS)
    with ThreadPoolExecutor(max_workers=jobs or min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(lambda n: build_one(n, shim_obj, force), names))
        list(ex.map(lambda n: build_dropin(n, force), [n for n in DROPIN_CONFIGS if n in names]))
    return dict(zip(names, res))


def available():
    return os.path.isdir(REF_CPU)


if __name__ == "__main__":
    force = "--force" in sys.argv
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or None
    for n, (s, d) in build(names, force=force).items():
        print(n, s, d)

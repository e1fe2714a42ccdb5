#!/usr/bin/env python3
"""Generate tests/golden/alpaka/*.npz from the UNMODIFIED alpaka tree of the reference (solverPoissonMPI_alpaka built with
alpaka's OpenMP-blocks CPU accelerator by oracle/build_ref_alpaka.py; run with OMP_NUM_THREADS=1 so that its atomicAdd
reductions are ordered).  These fixtures pin the alpaka-only configuration surface: fp32 Chebyshev
iterates (solverSetup.hpp:14) and block-local eigenvalue bounds (inputParam.hpp:21-22,27).  Per (config, px py pz) case:

    precond_x     the preconditioner class -- ChebyshevIterationAlpaka::operator()(bufX, bufB) (chebyshevIterationAlpaka.hpp:119-310) or
                  the nested BiCGstabAlpaka::operator()(bufX, bufB) (BiCGstabAlpaka.hpp:480-861) -- applied ONCE to the test field of
                  oracle/ref_dump_alpaka.cpp (`alpaka_test_field` below builds the same numbers); X on the global data range
    history, iters, precond_iters, error_iteration, error_operator, max_point_error       one full solve (src/main.cpp:83-101)
    np, nranks, ds, origin, bcs, tolerance, max_iter, cheb_max, cheb_rescale_min/max, cheb_f32, cheb_eig_local, precond  (the configuration)
Only runnable where /root/reference exists (the build container); the fixtures travel.
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import build_ref_alpaka as ba  # noqa: E402
from tests.golden.make_golden import assemble, read_summary  # noqa: E402

OUT = os.path.join(HERE, "alpaka")

CASES = [
    ("alp_f32_d24", (1, 1, 1)), ("alp_f32_d24", (2, 2, 2)),
    ("alp_f32_m24", (1, 1, 1)), ("alp_f32_m24", (2, 1, 2)), ("alp_f32_m24", (1, 1, 4)),
    ("alp_f32_m24_c24", (1, 2, 1)),
    ("alp_f32loc_m24", (1, 1, 2)), ("alp_f32loc_m24", (3, 2, 1)),
    ("alp_f32loc_d32", (2, 2, 2)),
    ("alp_loc_m24", (1, 1, 2)), ("alp_loc_m24", (2, 2, 1)),
    ("alp_f64_m24", (1, 1, 1)), ("alp_none_m24", (1, 1, 1)),
    ("alp_shipped32", (1, 1, 1)), ("alp_shipped32", (1, 1, 2)),
    # nested BiCGSTAB preconditioners of the alpaka tree (inputParam.hpp:31,33).  (The block-local one on more than one rank does not
    # finish within minutes in the alpaka tree itself -- not used.)
    ("alp_nbl_m24", (1, 1, 1)),
    ("alp_nbg_m24", (1, 1, 1)), ("alp_nbg_m24", (1, 1, 2)), ("alp_nbg_m24", (2, 2, 1)),
    ("alp_nbg_m24_i8", (1, 1, 1)), ("alp_nbg_m24_i8", (1, 1, 2)), ("alp_nbg_m24_i8", (3, 1, 2)),
]
TIMEOUT = 600


def alpaka_test_field(gk, gj, gi):
    """testField() of oracle/ref_dump_alpaka.cpp on integer index arrays (global, 0-based, data range)"""
    gi, gj, gk = (np.asarray(a, dtype=np.int64) for a in (gi, gj, gk))
    h1 = (gi * 73856093 + gj * 19349663 + gk * 83492791) % 4001
    h2 = (gi * 2654435761 + gj * 40503 + gk * 9973) % 1021
    return (h1 - 2000) / 4096.0 + h2 / 1099511627776.0


def run_files_case(name="alp_files_m24", ranks=(1, 1, 2)):
    """the tree's own main.cpp with writeResidual / writeSolution = true: the two result files as it writes them, next to the
    full-precision history and the per-rank blocks of the same (deterministic) solve from the dump harness"""
    c = ba.CONFIGS[name]
    world = ranks[0] * ranks[1] * ranks[2]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    with tempfile.TemporaryDirectory() as td:
        r = subprocess.run([os.path.join(ba.OUT, "bin", "alp_solver_" + name), *map(str, ranks)], check=True, capture_output=True, text=True,
                           env=env, cwd=td)
        hist_txt = np.frombuffer(open(td + "/residualHistory.txt", "rb").read(), dtype=np.uint8)
        sol = np.fromfile(td + "/solution.dat")
        stdout = r.stdout
    with tempfile.TemporaryDirectory() as td:
        subprocess.run([os.path.join(ba.OUT, "bin", "alp_dump_" + name), *map(str, ranks), td, "solve"], check=True, capture_output=True, text=True, env=env)
        s = read_summary(td + "/summary.txt")
        blocks = np.stack([np.fromfile(f"{td}/rank{q}.x") for q in range(world)])
        out = dict(residual_history_txt=hist_txt, solution_dat=sol, blocks=blocks, history=np.fromfile(td + "/history.bin"),
                   iters=int(s["iters"][0]), precond_iters=int(s["precond_iters"][0]), nranks=np.array(ranks), np=np.array(c["np"]),
                   max_iter=int(c["iter_max"]),
                   report_labels=np.array([l.split()[0] for l in stdout.splitlines() if l.startswith("timeTot")]))
    os.makedirs(OUT, exist_ok=True)
    fn = os.path.join(OUT, "files_%s_%d%d%d.npz" % ((name,) + tuple(ranks)))
    np.savez_compressed(fn, **out)
    return fn


def run_case(name, ranks):
    c = ba.CONFIGS[name]
    exe = os.path.join(ba.OUT, "bin", "alp_dump_" + name)
    world = ranks[0] * ranks[1] * ranks[2]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = dict(np=np.array(c["np"]), nranks=np.array(ranks), ds=np.array(c["ds"], dtype=float), origin=np.array(c["origin"], dtype=float),
               bcs=np.array(c["bcs"]), tolerance=float(c["toll_main"] * c["toll_scaling"]), max_iter=int(c["iter_max"]),
               cheb_max=int(c["cheb_max"]), cheb_rescale_min=float(c["rescale_min"]), cheb_rescale_max=float(c["rescale_max"]),
               cheb_f32=int(c["cheb_type"] == "float"), cheb_eig_local=int(c["solver"] == "bicgstab_cheblocal"),
               precond={"bicgstab_none": "none", "bicgstab_bicgloc": "bicgloc", "bicgstab_bicgglob": "bicgglob"}.get(c["solver"], "cheb"),
               precond_tolerance=float((c.get("toll_precond") or c["toll_main"] * 1e8) * c["toll_scaling"]),
               precond_max_iter=int(c.get("precond_iter_max") or 500))
    if out["precond"] != "none":
        with tempfile.TemporaryDirectory() as td:
            subprocess.run([exe, *map(str, ranks), td, "precond"], check=True, capture_output=True, text=True, env=env, timeout=TIMEOUT)
            for r in range(world):   # assemble() reads rank<r>.x
                os.rename(f"{td}/rank{r}.px", f"{td}/rank{r}.x")
            out["precond_x"] = assemble(td, world, c["np"])
    with tempfile.TemporaryDirectory() as td:
        r = subprocess.run([exe, *map(str, ranks), td, "solve"], check=True, capture_output=True, text=True, env=env, timeout=TIMEOUT)
        s = read_summary(td + "/summary.txt")
        maxerr = [l for l in r.stdout.splitlines() if l.startswith("Max error local point")]
        out.update(history=np.fromfile(td + "/history.bin"), iters=int(s["iters"][0]), precond_iters=int(s["precond_iters"][0]),
                   error_iteration=float(s["error_iteration"][0]), error_operator=float(s["error_operator"][0]),
                   max_point_error=float(maxerr[0].split()[4]) if maxerr else np.nan)
    os.makedirs(OUT, exist_ok=True)
    fn = os.path.join(OUT, f"{name}_{ranks[0]}{ranks[1]}{ranks[2]}.npz")
    np.savez_compressed(fn, **out)
    return fn, out["iters"]


if __name__ == "__main__":
    ba.build(sorted({c[0] for c in CASES} | {"alp_files_m24"}))
    only = set(sys.argv[1:])
    if not only or "alp_files_m24" in only:
        fn = run_files_case()
        print(os.path.basename(fn), os.path.getsize(fn) // 1024, "KiB")
    for name, ranks in CASES:
        if only and name not in only:
            continue
        fn, iters = run_case(name, ranks)
        print(os.path.basename(fn), "iters", iters, os.path.getsize(fn) // 1024, "KiB")

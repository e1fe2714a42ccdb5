#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Runs oracle/_ref/bin/ref_dump_<config> (the reference's own classes compiled against the
threads-as-ranks MPI shim, see oracle/build_ref.py) for a list of (config, px py pz) cases
and stores, per case, one compressed .npz with
    history       errorFromIterationHistory_[0..iters]      (BiCGSTAB.hpp:114-117,278-281)
    iters, norm_b, error_iteration, error_operator, tolerance
    x             final solution, data range only, assembled to the global grid (small cases)
    np, nranks, ds, origin, bcs, solver, precond, max_iter, cheb_max, order_neumann, cheb_rescale_min/max    (the configuration)
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
from oracle import build_ref as br  # noqa: E402

# (config name, ranks, store_solution)
CASES = [
    ("d16", (1, 1, 1), True), ("d16", (2, 2, 2), True), ("d16", (1, 1, 2), True),
    ("d32", (1, 1, 1), True), ("d32", (1, 1, 2), False), ("d32", (2, 1, 1), False), ("d32", (1, 2, 1), False), ("d32", (2, 2, 2), False),
    ("d32_cheb", (1, 1, 1), True), ("d32_cheb", (2, 2, 2), False),
    ("m24", (1, 1, 1), True), ("m24", (2, 2, 2), False), ("m24", (1, 1, 2), False),
    ("m24_cheb", (1, 1, 1), True), ("m24_cheb", (1, 1, 2), True), ("m24_cheb", (2, 2, 1), False),
    ("n24", (1, 1, 1), True), ("n24", (2, 1, 2), False),
    ("cg32", (1, 1, 1), True), ("cg32", (1, 2, 2), False),
    ("cg32_cheb", (1, 1, 1), True), ("cg32_cheb", (2, 1, 1), False),
    ("cgm24", (1, 1, 1), True),
    ("m32_cheb", (1, 1, 1), False), ("m32_cheb", (1, 1, 2), False),
    ("d64", (1, 1, 1), False), ("d64", (1, 1, 2), False), ("d64", (2, 2, 2), False),
    ("d64_t12", (1, 1, 1), False), ("d64_t12", (1, 1, 2), False), ("d64_t12", (2, 2, 2), False),
    ("d64_cheb", (1, 1, 1), False),
    ("default", (1, 1, 1), False), ("default", (1, 1, 2), False), ("default", (2, 1, 1), False), ("default", (2, 2, 2), False),
    # layouts with MIDDLE ranks (neighbours on both sides of an axis)
    ("d32", (1, 1, 4), False), ("d32", (4, 1, 1), False), ("d32_cheb", (1, 4, 1), False), ("d32_cheb", (1, 1, 4), True),
    ("m24", (3, 1, 1), False), ("m24", (1, 1, 4), False), ("m24_cheb", (3, 2, 1), False), ("m24_cheb", (1, 1, 4), False),
    ("cg32", (1, 1, 4), False), ("cg32_cheb", (4, 2, 1), False), ("d64", (1, 1, 8), False), ("d64_cheb", (1, 1, 8), False),
    # nested Krylov preconditioners
    ("nb24", (1, 1, 1), True), ("nb24", (1, 1, 2), False), ("nb24", (3, 1, 2), False),
    ("nc24", (1, 1, 1), True), ("nc24", (2, 2, 1), False), ("nc24", (3, 1, 2), False),
    # GLOBAL nested BiCGSTAB preconditioner (communicationON in the preconditioner slot; alpaka: T_PreconditionerBiCGStabGlobal)
    ("nbg24", (1, 1, 2), True), ("nbg24_i8", (1, 1, 1), True), ("nbg24_i8", (1, 1, 2), True), ("nbg24_i8", (2, 2, 1), False), ("nbg24_i8", (3, 1, 2), False),
    ("nbgd32", (2, 2, 2), False),
    # first-order Neumann closure (orderNeumanBcs = 1): BiCGSTAB, BiCGSTAB + Chebyshev, CG (which then resets ghosts)
    ("o1m24", (1, 1, 1), True), ("o1m24", (1, 1, 2), False), ("o1m24", (3, 1, 2), False),
    ("o1m24_cheb", (1, 1, 1), True), ("o1m24_cheb", (2, 2, 1), False),
    ("o1cgm24", (1, 1, 1), True), ("o1cgm24", (1, 2, 2), False),
    # Chebyshev iteration as main solver (fixed number of sweeps, no history)
    ("chm24", (1, 1, 1), True), ("chm24", (1, 1, 2), True), ("chm24", (3, 2, 1), False),
    ("chd32", (1, 1, 1), True), ("chd32", (2, 2, 2), False),
    # DIM = 2 and DIM = 1
    ("q24", (1, 1, 1), True), ("q24", (2, 1, 1), False), ("q24", (3, 2, 1), True),
    ("q24_cheb", (1, 1, 1), True), ("q24_cheb", (2, 2, 1), False),
    ("qd40", (1, 1, 1), True), ("qd40", (2, 3, 1), False),
    ("qcg40", (1, 1, 1), True), ("qcg40", (2, 2, 1), False),
    ("l48", (1, 1, 1), True), ("l48", (4, 1, 1), True),
    ("l48_cheb", (1, 1, 1), True), ("l48_cheb", (2, 1, 1), False),
    # GLOBAL Chebyshev preconditioner (communicationON in the preconditioner slot)
    ("d32_chebg", (1, 1, 2), True), ("d32_chebg", (2, 2, 2), False), ("m24_chebg", (1, 1, 2), True), ("m24_chebg", (3, 2, 1), False),
    # edge cases: non-divisible grid (truncated blocks), smallest blocks
    ("e25", (2, 2, 2), True), ("e25", (1, 1, 2), True), ("e25_cheb", (2, 2, 2), True), ("e25_cheb", (1, 2, 1), False),
    ("tiny6", (1, 1, 1), True), ("tiny6", (2, 2, 2), True), ("tiny6_cheb", (2, 2, 2), True), ("tiny6_cheb", (1, 1, 2), True),
    # the benchmarked configurations and their neighbours: full solves at 128^3 and 256^3, the first 20
    # iterations at 512^3.  An integer instead of True stores x on the sub-lattice of every s-th global point ("x_sample").
    # (bench1024_it20 needs ~110 GB of host memory: tools/make_golden_1024.py runs it on the GPU box's host)
    ("d128", (1, 1, 1), 4), ("d128", (1, 1, 2), False), ("bench256", (2, 2, 2), 8), ("bench256", (1, 1, 8), False),
    ("bench512_it20", (1, 1, 1), 16), ("bench512_it20", (1, 1, 8), False),
]
BIG = {"d128", "bench256", "bench512_it20", "bench1024_it20"}   # ref_dump writes only the final x (PPS_DUMP_LIGHT)


def read_summary(path):
    s = {}
    for line in open(path):
        k, *v = line.split()
        s[k] = v
    return s


def read_meta(path):
    m = {}
    for line in open(path):
        k, *v = line.split()
        m[k] = [float(t) if "." in t or "e" in t else int(t) for t in v]
    return m


def assemble(outdir, world, npglobal, stride=1):
    """global data-range solution; stride s > 1: only the points whose global indices are all multiples of s"""
    s = stride
    g = np.zeros(tuple((npglobal[d] + s - 1) // s for d in (2, 1, 0)))
    for r in range(world):
        m = read_meta(f"{outdir}/rank{r}.meta")
        ng, nn, loc = m["nlocal_guards"], m["nlocal_noguards"], m["global_location"]
        x = np.memmap(f"{outdir}/rank{r}.x", dtype=np.float64, mode="r").reshape(ng[2], ng[1], ng[0])
        o = [loc[d] * nn[d] for d in range(3)]
        # axes >= DIM hold one point without guards (blockGrid.hpp:160-182)
        inner = tuple(slice(1, -1) if ng[d] > nn[d] else slice(None) for d in (2, 1, 0))
        xi = x[inner]
        first = [(-o[d]) % s for d in range(3)]                    # first local index on the sub-lattice
        dst = [(o[d] + first[d]) // s for d in range(3)]
        sub = np.array(xi[first[2]::s, first[1]::s, first[0]::s])
        g[dst[2]:dst[2] + sub.shape[0], dst[1]:dst[1] + sub.shape[1], dst[0]:dst[0] + sub.shape[2]] = sub
    return g


def run_case(name, ranks, store_x):
    c = br.CONFIGS[name]
    exe = os.path.join(br.OUT, "bin", "ref_dump_" + name)
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, PPS_DUMP_LIGHT="1") if name in BIG else None
        r = subprocess.run([exe, *map(str, ranks), td], check=True, capture_output=True, text=True, env=env)
        s = read_summary(td + "/summary.txt")
        hist = np.fromfile(td + "/history.bin")
        if c["solver"] == "cheb_main":
            # the reference never fills errorFromIterationHistory_ in this mode: keep only the final residual
            hist = np.array([float(s["error_operator"][0])])
        world = ranks[0] * ranks[1] * ranks[2]
        maxerr = [l for l in r.stdout.splitlines() if l.startswith("Max error local point")]
        out = dict(
            history=hist, iters=int(s["iters"][0]), norm_b=float(s["norm_b"][0]),
            error_iteration=float(s["error_iteration"][0]), error_operator=float(s["error_operator"][0]),
            tolerance=float(s["tolerance"][0]), max_iter=int(s["max_iter"][0]),
            np=np.array(c["np"]), nranks=np.array(ranks), ds=np.array(c["ds"], dtype=float),
            origin=np.array(c["origin"], dtype=float), bcs=np.array(c["bcs"]),
            solver=c["solver"].split("_")[0], precond=c["solver"].split("_")[1], cheb_max=c["cheb_max"],
            max_point_error=float(maxerr[0].split()[4]) if maxerr else np.nan,
            order_neumann=c.get("order_neumann", 2), dim=c.get("dim", 3),
            cheb_rescale_min=500.0 if c.get("rescale_min") is None else float(c["rescale_min"]),
            cheb_rescale_max=1 - 1e-4 if c.get("rescale_max") is None else float(c["rescale_max"]),
            precond_max_iter=150 if c.get("precond_iter_max") is None else int(c["precond_iter_max"]),
        )
        if store_x is True:
            out["x"] = assemble(td, world, c["np"])
        elif store_x:
            out["x_sample"] = assemble(td, world, c["np"], int(store_x))
            out["x_stride"] = int(store_x)
    fn = os.path.join(HERE, f"{name}_{ranks[0]}{ranks[1]}{ranks[2]}.npz")
    np.savez_compressed(fn, **out)
    return fn, out["iters"]


if __name__ == "__main__":
    br.build(sorted({c[0] for c in CASES}))
    only = set(sys.argv[1:])
    for name, ranks, store_x in CASES:
        if only and name not in only:
            continue
        fn, iters = run_case(name, ranks, store_x)
        print(os.path.basename(fn), "iters", iters, os.path.getsize(fn) // 1024, "KiB")

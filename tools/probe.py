#!/usr/bin/env python3
"""GPU probe: kernel-level bandwidth numbers for tuning (not the bench).  Writes JSON lines to stdout.

    python tools/probe.py stencil 512            # y = A x sweep over tilings
    python tools/probe.py solve 256 [cheb]       # one solve with per-kernel CUDA-event timings
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallelpoissonsolver_b200 as pps  # noqa: E402

PASSES = {  # vector passes (8 B per cell each) per launch of a kernel class
    "stencil_dot(v=A*p, r0.v)": 3, "s_update(r-=alpha*v)": 3, "stencil_dot2(t=A*s, s.t, t.t)": 2,
    "xr_update(x+=.., r-=omega*t, r0.r, r.r)": 7, "p_update(p=r+beta*(p-omega*v))": 4, "cheb_first": 3, "cheb_step": 4,
    "residual(r=b-A*x, r.r)": 3, "fused_p(p=r+beta*(p-omega*v), v=A*p, r0.v)": 6, "fused_s(s=r-alpha*v, t=A*s, s.t, t.t)": 4, "stencil(y=A*x)": 2, "cg_apply(Ap, r.z, p.Ap)": 3, "cg_xr": 6, "cg_p": 3, "dot": 2,
}


def manufactured(n, ds=0.1):
    """x0 (Dirichlet planes = u) and b = f on an n^3 all-Dirichlet grid, reference layout"""
    c = np.arange(n) * ds
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    u = np.sin(x) + np.cos(y) + 3 * np.sin(z) + x * x * y * z + x * x + 10
    f = -np.sin(x) - np.cos(y) - 3 * np.sin(z) + 2 * y * z + 2
    X = np.zeros((n + 2,) * 3)
    B = np.zeros((n + 2,) * 3)
    B[1:-1, 1:-1, 1:-1] = f
    inner = np.zeros((n,) * 3, bool)
    inner[1:-1, 1:-1, 1:-1] = True
    X[1:-1, 1:-1, 1:-1] = np.where(inner, 0.0, u)
    return X, B


def stencil_sweep(n):
    variants = [dict(PPS_STENCIL_TMA="0", PPS_TILE_ROWS="8"), dict(PPS_STENCIL_TMA="1", PPS_TMA_ROWS="8"),
                dict(PPS_STENCIL_TMA="1", PPS_TMA_ROWS="16")]
    for v in variants:
        for zc in (0, 16, 32, 64, 128, n):
            os.environ.update(v)
            os.environ["PPS_ZCHUNK_STENCIL"] = str(zc)
            s = pps.PoissonSolver(pps.make_config((n, n, n)))
            for dot in (False, True):
                ms = s.bench_operator(20, dot)
                nbytes = n ** 3 * 8 * (3 if dot else 2)
                print(json.dumps(dict(kind="stencil", n=n, variant=v, zchunk=zc, dot=dot, ms=ms, gbs=nbytes / ms / 1e6)), flush=True)
            s.close()


def solve(n, cheb=False, max_iter=4000):
    X, B = manufactured(n)
    cfg = pps.make_config((n, n, n), precond=pps.PRECOND_CHEBYSHEV if cheb else pps.PRECOND_NONE, max_iter=max_iter)
    s = pps.PoissonSolver(cfg)
    s.set_fields(0, X, B)
    s.save_fields()
    s.solve()                      # warm-up
    s.restore_fields()
    s.set_profiling(True)
    t0 = time.time()
    s.solve()
    wall = time.time() - t0
    cells = n ** 3
    out = dict(kind="solve", n=n, cheb=cheb, iters=s.iterations, err=s.error_iteration, err_true=s.error_operator,
               solver_s=s.solver_seconds, loop_s=s.loop_seconds, wall_s=wall, launches=s.launch_count,
               mlups=cells * s.iterations / s.solver_seconds / 1e6,
               gbs_algorithmic=136 * cells * s.iterations / s.loop_seconds / 1e9 if not cheb else None, kernels=[])
    for k in s.kernel_stats():
        p = PASSES.get(k["name"])
        k["gbs"] = p * 8 * cells / k["avg_ms"] / 1e6 if p and k["avg_ms"] > 0 else None
        out["kernels"].append(k)
    print(json.dumps(out), flush=True)
    s.set_profiling(False)
    s.restore_fields()
    s.solve()
    print(json.dumps(dict(kind="solve_noprof", n=n, iters=s.iterations, solver_s=s.solver_seconds, loop_s=s.loop_seconds,
                          mlups=cells * s.iterations / s.solver_seconds / 1e6,
                          gbs_algorithmic=136 * cells * s.iterations / s.loop_seconds / 1e9 if not cheb else None)), flush=True)
    s.close()


def iters(n, k):
    """K iterations of the bench problem at n^3 (short: for ncu captures)"""
    X, B = manufactured(n)
    s = pps.PoissonSolver(pps.make_config((n, n, n), max_iter=k))
    s.set_fields(0, X, B)
    s.solve()
    print(json.dumps(dict(kind="iters", n=n, iters=s.iterations, err=s.error_iteration, loop_s=s.loop_seconds,
                          kernels=[k_["name"] for k_ in s.kernel_stats()])), flush=True)
    s.close()


if __name__ == "__main__":
    what = sys.argv[1]
    n = int(sys.argv[2])
    if what == "stencil":
        stencil_sweep(n)
    elif what == "iters":
        iters(n, int(sys.argv[3]))
    else:
        solve(n, cheb=len(sys.argv) > 3 and sys.argv[3] == "cheb")

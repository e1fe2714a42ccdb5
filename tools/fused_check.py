#!/usr/bin/env python3
"""Diagnostics for the experimental 17-pass schedule (PPS_FUSE_FULL): FULL vs SPLIT, bitwise, on one GPU.

    python tools/fused_check.py [sizes ...]                 fresh handles, 40 fixed iterations (what round 1 verified up to 512^3)
    python tools/fused_check.py --converge [sizes ...]      tolerance 1e-8, THREE solves per handle (save/restore): the flow in which
                                                            round 1 saw the 512^3 anomaly (iteration counts 1459 / 1242, true residual
                                                            4e-7 vs recurrence 4e-9).  Bisect with PPS_FUSE_P=0 / PPS_FUSE_S=0.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# one reduction tree for both schedules: the defaults pick z-chunks and tile heights per kernel
os.environ.setdefault("PPS_ZCHUNK_STENCIL", "32")
os.environ.setdefault("PPS_FUSE_BY_S", "8")
import parallelpoissonsolver_b200 as pps  # noqa: E402
from tools.probe import manufactured  # noqa: E402


def fixed(n):
    X, B = manufactured(n)
    res = {}
    for name, fus in (("split", pps.FUSE_SPLIT), ("full1", pps.FUSE_FULL), ("full2", pps.FUSE_FULL)):
        s = pps.PoissonSolver(pps.make_config((n, n, n), max_iter=40, tolerance=1e-30, fusion=fus))
        s.set_fields(0, X, B)
        s.solve()
        res[name] = (s.history().copy(), s.get_solution(0).copy(), s.error_operator)
        s.close()
    h0 = res["split"][0]
    out = {"mode": "fixed40", "n": n}
    for k in ("full1", "full2"):
        h = res[k][0]
        neq = np.nonzero(h != h0)[0]
        out[k + "_first_diff_iter"] = int(neq[0]) if len(neq) else None
        out[k + "_x_maxdiff"] = float(np.abs(res[k][1] - res["split"][1]).max())
        out[k + "_true_res"] = res[k][2]
    out["split_true_res"] = res["split"][2]
    out["full1_eq_full2"] = bool(np.array_equal(res["full1"][0], res["full2"][0]))
    print(json.dumps(out), flush=True)


def converge(n):
    X, B = manufactured(n)
    out = {"mode": "converge_x3", "n": n, "PPS_FUSE_P": os.environ.get("PPS_FUSE_P", "1"), "PPS_FUSE_S": os.environ.get("PPS_FUSE_S", "1")}
    hists = {}
    for name, fus in (("split", pps.FUSE_SPLIT), ("full", pps.FUSE_FULL)):
        s = pps.PoissonSolver(pps.make_config((n, n, n), max_iter=6000, tolerance=1e-8, fusion=fus))
        s.set_fields(0, X, B)
        s.save_fields()
        runs = []
        for rep in range(3):
            if rep:
                s.restore_fields()
            s.solve()
            runs.append(dict(iters=s.iterations, err=s.error_iteration, err_true=s.error_operator))
            hists[(name, rep)] = s.history().copy()
        out[name] = runs
        s.close()
    ref = hists[("split", 0)]
    for key, h in hists.items():
        m = min(len(h), len(ref))
        neq = np.nonzero(h[:m] != ref[:m])[0]
        out[f"{key[0]}{key[1]}_first_diff_vs_split0"] = int(neq[0]) if len(neq) else (None if len(h) == len(ref) else m)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    sizes = [int(a) for a in args] or [96, 128, 192, 256]
    for n in sizes:
        (converge if "--converge" in sys.argv else fixed)(n)

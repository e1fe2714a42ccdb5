#!/usr/bin/env python3
"""Diagnostic for the experimental 17-pass schedule: FULL vs SPLIT, bitwise, at growing sizes (one GPU)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallelpoissonsolver_b200 as pps
from tools.probe import manufactured

for n in [int(a) for a in sys.argv[1:]] or [96, 128, 192, 256]:
    X, B = manufactured(n)
    res = {}
    for name, fus in (("split", pps.FUSE_SPLIT), ("full1", pps.FUSE_FULL), ("full2", pps.FUSE_FULL)):
        s = pps.PoissonSolver(pps.make_config((n, n, n), max_iter=40, tolerance=1e-30, fusion=fus))
        s.set_fields(0, X, B)
        s.solve()
        res[name] = (s.history().copy(), s.get_solution(0).copy(), s.error_operator)
        s.close()
    h0 = res["split"][0]
    out = {"n": n}
    for k in ("full1", "full2"):
        h = res[k][0]
        neq = np.nonzero(h != h0)[0]
        out[k + "_first_diff_iter"] = int(neq[0]) if len(neq) else None
        out[k + "_x_maxdiff"] = float(np.abs(res[k][1] - res["split"][1]).max())
        out[k + "_true_res"] = res[k][2]
    out["split_true_res"] = res["split"][2]
    out["full1_eq_full2"] = bool(np.array_equal(res["full1"][0], res["full2"][0]))
    print(json.dumps(out), flush=True)

#!/usr/bin/env python3
"""Forensics of the fused_s mismatch: stop at the first bad launch, then search the WHOLE s_ref / r / v arrays for the
value that the wrong t implies for each stencil neighbour -- where did the stale datum come from?"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallelpoissonsolver_b200 as pps  # noqa: E402

S_OUT, T_OUT, S_REF, T_REF, R, V = 0, 1, 2, 3, 4, 5
ALPHA = 0.37


def whole(s, arr, total):
    out = np.empty(total)
    step = 1 << 24
    for o in range(0, total, step):
        n = min(step, total - o)
        out[o:o + n] = s.debug_peek(arr, o, n)
    return out


def run(n, variant, zchunk, tries=8):
    os.environ["PPS_ZCHUNK_STENCIL"] = str(zchunk)
    for attempt in range(tries):
        s = pps.PoissonSolver(pps.make_config((n, n, n), max_iter=10, fusion=pps.FUSE_FULL))
        out = s.debug_fused(0, variant, -600)
        if out[3] == 0:
            s.close()
            continue
        pitch, plane = out[4], out[5]
        total = plane * (n + 2)
        tg, tr = whole(s, T_OUT, total), whole(s, T_REF, total)
        sr, rr, vv = whole(s, S_REF, total), whole(s, R, total), whole(s, V, total)
        so = whole(s, S_OUT, total)
        bad = np.nonzero(tg != tr)[0]
        dec = lambda q: (int(q % pitch - 15), int((q % plane) // pitch), int(q // plane))
        rec = dict(n=n, variant=variant, zchunk=zchunk, rep=out[8], n_bad=len(bad), s_out_equals_ref=bool(np.array_equal(so, sr)),
                   rows=sorted({(dec(q)[1], dec(q)[2]) for q in bad}), warp_rows=sorted({(dec(q)[1] - 1) % 8 for q in bad}),
                   plane_in_chunk=sorted({(dec(q)[2] - 2) % zchunk for q in bad}), cells=[])
        nbrs = (("xm", -1), ("xp", 1), ("ym", -pitch), ("yp", pitch), ("zm", -plane), ("zp", plane))
        for q in bad[:6]:
            implied = (tg[q] - tr[q]) * 0.01
            cell = dict(at=dec(q), implied=implied, found={})
            for nm, d in nbrs:
                want = sr[q + d] + implied
                tol = 2e-13
                hit = np.nonzero(np.abs(sr - want) <= tol)[0]
                if len(hit) and len(hit) < 20:
                    cell["found"][nm + ":s"] = [dec(h) for h in hit]
                hit = np.nonzero(np.abs(rr - (want + ALPHA * vv[q + d])) <= tol)[0]       # only r stale
                if len(hit) and len(hit) < 20:
                    cell["found"][nm + ":r_only"] = [dec(h) for h in hit]
                hit = np.nonzero(np.abs(vv - (rr[q + d] - want) / ALPHA) <= 4 * tol)[0]   # only v stale
                if len(hit) and len(hit) < 20:
                    cell["found"][nm + ":v_only"] = [dec(h) for h in hit]
                if abs(want) <= tol:
                    cell["found"][nm + ":zero"] = True
                if abs(want - rr[q + d]) <= tol:
                    cell["found"][nm + ":v_zero"] = True
                if abs(want + ALPHA * vv[q + d]) <= tol:
                    cell["found"][nm + ":r_zero"] = True
            rec["cells"].append(cell)
        print(json.dumps(rec), flush=True)
        s.close()
        return
    print(json.dumps(dict(n=n, variant=variant, zchunk=zchunk, result="no bad launch")), flush=True)


if __name__ == "__main__":
    for variant, zc in ((6, 32), (6, 32), (4, 32)):
        run(512, variant, zc)

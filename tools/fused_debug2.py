#!/usr/bin/env python3
"""Second-level diagnosis of the fused_s mismatch: stop at the first bad launch, dump the bad z-column of t and the
neighbourhood of s, and find which neighbour value the wrong t implies."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallelpoissonsolver_b200 as pps  # noqa: E402

S_OUT, T_OUT, S_REF, T_REF, R, V = 0, 1, 2, 3, 4, 5


def run(n, variant, zchunk, tries=6):
    os.environ["PPS_ZCHUNK_STENCIL"] = str(zchunk)
    for attempt in range(tries):
        s = pps.PoissonSolver(pps.make_config((n, n, n), max_iter=10, fusion=pps.FUSE_FULL))
        out = s.debug_fused(0, variant, -600)
        if out[3] == 0:
            s.close()
            continue
        pitch, plane = out[4], out[5]
        e = out[8:13]
        idx = e[4]
        inv = 1.0 / 0.01
        i, j, k = idx % pitch - 15, (idx % plane) // pitch, idx // plane
        rec = dict(n=n, variant=variant, zchunk=zchunk, rep=e[0], n_bad_t=e[3], first=dict(i=i, j=j, k=k), column=[])
        # the z-column of the first bad cell, and its x-row / y-column in the first bad plane
        for m in range(-2, zchunk + 2):
            o = idx + m * plane
            tg, tr = s.debug_peek(T_OUT, o, 1)[0], s.debug_peek(T_REF, o, 1)[0]
            if tg != tr or abs(m) <= 1:
                sref = {nm: s.debug_peek(S_REF, o + d, 1)[0] for nm, d in (("c", 0), ("xm", -1), ("xp", 1), ("ym", -pitch), ("yp", pitch), ("zm", -plane), ("zp", plane))}
                implied = (tg - tr) / inv      # wrong neighbour value minus right neighbour value
                # candidates: which stale / displaced value would explain it?
                cands = {}
                for nm in ("xm", "xp", "ym", "yp", "zm", "zp"):
                    want = sref[nm] + implied
                    hits = []
                    for dk in range(-8, 9):
                        for dj in range(-2, 3):
                            row = s.debug_peek(S_REF, o + dk * plane + dj * pitch - 4, 9)
                            for di, val in enumerate(row):
                                if abs(val - want) <= 1e-12 * max(1.0, abs(want)):
                                    hits.append((di - 4, dj, dk))
                    if abs(want) <= 1e-13:
                        hits.append("zero")
                    if hits:
                        cands[nm] = hits
                rec["column"].append(dict(dz=m, t_got=tg, t_ref=tr, implied_delta=implied, s_ref=sref, explains=cands))
        # geometry of all bad cells in the first bad plane (x-row and y-column through the first bad cell)
        rowg, rowr = s.debug_peek(T_OUT, idx - 70, 140), s.debug_peek(T_REF, idx - 70, 140)
        rec["bad_in_x_row"] = [int(q - 70) for q in np.nonzero(rowg != rowr)[0]]
        colbad = []
        for dj in range(-10, 11):
            if s.debug_peek(T_OUT, idx + dj * pitch, 1)[0] != s.debug_peek(T_REF, idx + dj * pitch, 1)[0]:
                colbad.append(dj)
        rec["bad_in_y_column"] = colbad
        print(json.dumps(rec), flush=True)
        s.close()
        return
    print(json.dumps(dict(n=n, variant=variant, zchunk=zchunk, result="no bad launch in %d x 600" % tries)), flush=True)


if __name__ == "__main__":
    for variant, zc in ((6, 32), (6, 32), (4, 32), (6, 16)):
        run(512, variant, zc)

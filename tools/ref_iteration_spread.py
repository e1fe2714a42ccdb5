#!/usr/bin/env python3
"""How far does the UNMODIFIED reference's own iteration count move when only its summation order changes?

Runs oracle/_ref/bin/ref_solver_<config> (the reference's main.cpp, threads-as-ranks) on many rank layouts of the same
problem.  A layout changes nothing but the order in which the dot products are summed (unpreconditioned BiCGSTAB; the
operator and the axpys are pointwise identical), so the spread of the printed iteration counts is the yardstick for
"same iteration count" between the reference and ANY implementation with another summation order -- ours included.

    python tools/ref_iteration_spread.py d64 d128 [--max-ranks 8] > profiles/r02_reference_iteration_spread.jsonl

CPU only (runs in the build container; nothing here touches the GPU path).
"""
import itertools
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIZES = {"d16": 16, "d32": 32, "d64": 64, "d128": 128, "bench256": 256}


def layouts(n, max_ranks):
    out = []
    for px, py, pz in itertools.product((1, 2, 4, 8), repeat=3):
        if px * py * pz <= max_ranks and all(n % p == 0 and n // p >= 3 for p in (px, py, pz)):
            out.append((px, py, pz))
    return sorted(out, key=lambda l: (l[0] * l[1] * l[2], l))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    max_ranks = 8
    if "--max-ranks" in sys.argv:
        max_ranks = int(sys.argv[sys.argv.index("--max-ranks") + 1])
        args = [a for a in args if a != str(max_ranks)]
    for name in args:
        exe = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_solver_" + name)
        n = SIZES[name]
        rows = []
        for lay in layouts(n, max_ranks):
            out = subprocess.run([exe] + [str(v) for v in lay], capture_output=True, text=True).stdout
            m = re.search(r"finished with iter: (\d+) error from algo (\S+) error r=b-Ax (\S+)", out)
            rows.append({"layout": list(lay), "iters": int(m.group(1)), "err": float(m.group(2)), "err_true": float(m.group(3))})
            print(json.dumps({"config": name, **rows[-1]}), flush=True)
        its = [r["iters"] for r in rows]
        print(json.dumps({"config": name, "summary": True, "layouts": len(its), "iters_min": min(its), "iters_max": max(its),
                          "spread_rel": (max(its) - min(its)) / min(its)}), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""BASELINE.json configs[4]: operator-apply-only bandwidth sweep, y = A x on one GPU, 64^3 up to the largest cube
that fits, GB/s against the HBM roofline (16 algorithmic bytes per cell: read x, write y).  JSON lines on stdout.

    python tools/bandwidth_sweep.py [sizes ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import parallelpoissonsolver_b200 as pps  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 192, 256, 384, 512, 640, 768, 1024]
    peak, src = bench.measured_peak()
    for n in sizes:
        try:
            s = pps.PoissonSolver(pps.make_config((n, n, n)))
        except pps.PpsError as e:
            print(json.dumps(dict(n=n, error=str(e)[:200])), flush=True)
            break
        reps = max(5, min(200, int(2e9 / n ** 3)))
        for dot in (False, True):
            ms = s.bench_operator(reps, dot)
            nbytes = n ** 3 * 8 * (3 if dot else 2)
            print(json.dumps(dict(kind="apply+dot" if dot else "apply", n=n, cells=n ** 3, reps=reps, ms=ms, mlups=n ** 3 / ms / 1e3,
                                  gbs=nbytes / ms / 1e6, frac_of_peak=nbytes / ms / 1e6 / peak, peak_gbs=peak, peak_source=src,
                                  note="arrays smaller than the 126 MB L2 are served from L2 between repetitions" if n ** 3 * 16 < 126e6 else "")),
                  flush=True)
        s.close()


if __name__ == "__main__":
    main()

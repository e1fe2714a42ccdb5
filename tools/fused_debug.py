#!/usr/bin/env python3
"""Hunt for the nondeterminism of the 17-pass schedule: repeat ONE fused launch on frozen inputs and compare
with the split kernels (pps_debug_fused), over ring depths / z-chunks; then one converging solve with the in-solve checker.

    python tools/fused_debug.py [n ...]
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(n, which, variant, reps, zchunk):
    import parallelpoissonsolver_b200 as pps
    os.environ["PPS_ZCHUNK_STENCIL"] = str(zchunk)
    s = pps.PoissonSolver(pps.make_config((n, n, n), max_iter=10, fusion=pps.FUSE_FULL))
    out = s.debug_fused(which, variant, reps)
    s.close()
    ev = [out[8 + 5 * e: 13 + 5 * e] for e in range(4) if out[8 + 5 * e + 1] or out[8 + 5 * e + 3] or e < out[3]]
    pitch, plane = out[4], out[5]

    def dec(i):
        return None if i < 0 else dict(i=i % pitch - 15, j=(i % plane) // pitch, k=i // plane)
    print(json.dumps(dict(n=n, kernel="fused_s" if which == 0 else "fused_p", variant=variant, reps=reps, zchunk=out[6], ctas=out[7],
                          bad_operand_cells=out[0], bad_result_cells=out[1], bad_sum_launches=out[2], bad_launches=out[3],
                          events=[dict(rep=e[0], n_operand=e[1], first_operand=dec(e[2]), n_result=e[3], first_result=dec(e[4])) for e in ev[:out[3]]])),
          flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one(*map(int, sys.argv[2:7]))
        sys.exit(0)
    sizes = [int(a) for a in sys.argv[1:]] or [256, 512]
    for n in sizes:
        for which, variants in ((0, (6, 4, 3, 16)), (1, (4, 3))):
            for zc in (32, 16):
                for v in variants:
                    # a fresh process per configuration: a sticky CUDA error must not hide the rest
                    subprocess.run([sys.executable, __file__, "--one", str(n), str(which), str(v), "400", str(zc)], timeout=300)

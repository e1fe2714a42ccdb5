#!/usr/bin/env python3
"""First 20 iterations of the UNMODIFIED reference on the 1024^3 benchmark problem (BASELINE.json configs[2]), 1x1x8 ranks:
needs ~115 GB of host memory, so it runs on the GPU box's host (gpurun) with the binary built here by oracle/build_ref.py.
Writes gpurun_out/bench1024_it20_118.npz (copy it to tests/golden/)."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

if __name__ == "__main__":
    avail_gb = 0
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable"):
            avail_gb = int(line.split()[1]) / 1e6
    print("host memory available: %.0f GB, cores %d" % (avail_gb, os.cpu_count()), flush=True)
    if avail_gb < 140:
        print("not enough host memory for the 1024^3 reference run; skipped")
        sys.exit(0)
    import make_golden as mg
    ranks = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 1, 8)
    fn, iters = mg.run_case("bench1024_it20", ranks, 32)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    shutil.copy(fn, os.path.join(ROOT, "gpurun_out", os.path.basename(fn)))
    print(os.path.basename(fn), "iters", iters)

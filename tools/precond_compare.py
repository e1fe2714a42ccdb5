#!/usr/bin/env python3
"""BASELINE.json configs[3]: preconditioned (block-Jacobi Chebyshev(11), the stack solverSetup.hpp / inputParam.hpp ship)
against unpreconditioned BiCGSTAB on the same grid and rank layout: iteration count and time to tolerance.

    python tools/precond_compare.py [N]                                   one GPU, N^3 (default 256)
    torchrun --nproc-per-node 8 tools/precond_compare.py 768              1x1x8 slabs of 768^3

The block-Jacobi preconditioner makes the iteration count depend on the layout, so the CPU
reference to compare with is the oracle on the SAME layout: pass --oracle to run it too (sizes the host can hold).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import parallelpoissonsolver_b200 as pps  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if args else 256
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import torch
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    D = bench.Dist(rank, world, "cuda")
    npglobal = (n, n, n)
    X, B = bench.manufactured_slab(npglobal, world, rank)
    out = {"npglobal": npglobal, "layout": [1, 1, world]}
    variants = [("unpreconditioned", dict(precond=pps.PRECOND_NONE)),
                ("chebyshev11_block_jacobi", dict(precond=pps.PRECOND_CHEBYSHEV, cheb_block=0)),                   # one kernel per sweep: 280 B/cell per call
                ("chebyshev11_block_jacobi_blocked3", dict(precond=pps.PRECOND_CHEBYSHEV, cheb_block=3)),          # 3 sweeps per HBM pass: 96 B/cell per call
                ("chebyshev11_block_jacobi_blocked3_fp32", dict(precond=pps.PRECOND_CHEBYSHEV, cheb_block=3, cheb_precision=pps.CHEB_FP32)),
                ("chebyshev11_block_jacobi_blocked4_fp32", dict(precond=pps.PRECOND_CHEBYSHEV, cheb_block=4, cheb_precision=pps.CHEB_FP32)),
                ("chebyshev11_local_eig_blocked3", dict(precond=pps.PRECOND_CHEBYSHEV, cheb_block=3, cheb_eigenvalues=pps.CHEB_EIG_LOCAL)),
                ("chebyshev11_global_comm", dict(precond=pps.PRECOND_CHEBYSHEV, precond_communication=1))]
    if "--quick" in sys.argv:
        variants = variants[:3] + variants[3:4]
    for name, kw in variants:
        uid = D.bcast_bytes(pps.get_unique_id() if rank == 0 else None, 128) if world > 1 else None
        s = pps.PoissonSolver(pps.make_config(npglobal, nranks=(1, 1, world), bcs=(0,) * 6, tolerance=1e-8, max_iter=6000, device=local, **kw),
                              rank=rank, world_size=world, unique_id=uid)
        my = rank if world > 1 else 0
        s.set_fields(my, X, B)
        s.save_fields()
        s.set_max_iterations(30)
        s.solve()                      # warm-up: 30 iterations of the same loop
        s.set_max_iterations(6000)
        s.restore_fields()
        D.barrier()
        s.solve()
        D.barrier()
        secs = D.max(s.solver_seconds)
        out[name] = dict(iterations=s.iterations, seconds=secs, ms_per_iteration=D.max(s.loop_seconds) / max(1, s.iterations) * 1e3,
                         true_residual=s.error_operator, mlups=n ** 3 * s.iterations / secs / 1e6, launches=s.launch_count)
        if rank == 0:
            print("#", name, json.dumps(out[name]), file=sys.stderr, flush=True)
        s.close()
    a = out["unpreconditioned"]
    for name, _ in variants[1:]:
        out[name]["iteration_ratio_vs_unpreconditioned"] = a["iterations"] / out[name]["iterations"]
        out[name]["speedup_vs_unpreconditioned"] = a["seconds"] / out[name]["seconds"]
    if "--oracle" in sys.argv and rank == 0:
        from oracle import pyoracle as po
        for name, pre in (("unpreconditioned", po.PRECOND_NONE), ("chebyshev11_block_jacobi", po.PRECOND_CHEBYSHEV)):
            o = po.Oracle(po.make_config(npglobal, (1, 1, world), bcs=(0,) * 6, precond=pre, tolerance=1e-8, max_iter=6000))
            o.set_problem()
            o.solve()
            out[name]["oracle_iterations"] = o.iters
            o.close()
    if rank == 0:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

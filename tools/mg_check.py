#!/usr/bin/env python3
"""Multi-GPU parity check, one process per GPU (run under torchrun; tests/test_gpu_multi.py drives it).

    torchrun --nproc-per-node N tools/mg_check.py PX PY PZ [cheb | chebg | nbg] [cg] [fusecmp]
        cheb   block-Jacobi Chebyshev(11) preconditioner        chebg  the same with communicationON (global polynomial)
        nbg    GLOBAL nested BiCGSTAB preconditioner, nested solves capped at 8 iterations (validated on one GPU hosting all
               blocks at the end of round 2; this flag is how its one-block-per-GPU NCCL leg gets its first multi-GPU run)

Every rank hosts its block of the PX x PY x PZ decomposition on its own GPU (NCCL halo exchange + allreduce);
the CPU oracle runs the same layout in one process; rank 0 compares ||b||, the residual history and, after a
gather, the solution.  Prints one JSON line and exits non-zero on mismatch.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import parallelpoissonsolver_b200 as pps  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    px, py, pz = (int(v) for v in sys.argv[1:4])
    flags = sys.argv[4:]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert px * py * pz == world
    import threading
    wd = threading.Timer(240, lambda: os._exit(3))   # never hold a GPU box on a stalled collective
    wd.daemon = True
    wd.start()
    backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group(backend, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    dev = "cuda" if backend == "nccl" else "cpu"

    np_ = (24 * px if px > 1 else 40, 20 * py if py > 1 else 24, 16 * pz if pz > 1 else 20)
    ocfg = po.make_config(np_, (px, py, pz), ds=(0.1, 0.12, 0.09), origin=(0.3, -0.2, 0.1), bcs=(0, 1, 0, 1, 0, 1),
                          solver=po.SOLVER_CG if "cg" in flags else po.SOLVER_BICGSTAB,
                          precond=(po.PRECOND_BICGSTAB_LOCAL if "nbg" in flags else
                                   po.PRECOND_CHEBYSHEV if ("cheb" in flags or "chebg" in flags) else po.PRECOND_NONE),
                          precond_comm=int("nbg" in flags or "chebg" in flags), precond_max_iter=8 if "nbg" in flags else 150, tolerance=1e-8)
    if "cg" in flags:
        ocfg.bcs[:] = [0] * 6
    o = po.Oracle(ocfg)
    o.set_problem()

    def solve_with(fusion):
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(pps.get_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().numpy().tobytes())
        s = pps.PoissonSolver(H.pps_config_from_oracle(ocfg, device=local, fusion=fusion), rank=rank, world_size=world, unique_id=uid)
        H.hand_over_problem(o, s, ranks=[rank])
        s.solve()
        return s, s.get_solution(rank)

    if "fusecmp" in flags:
        os.environ["PPS_ZCHUNK_STENCIL"] = "8"   # one reduction tree for both schedules (the default picks z-chunks per kernel)
        os.environ["PPS_FUSE_BY_S"] = "8"        # ... and tile heights
    s, x = solve_with(pps.FUSE_AUTO)
    fused_equals_split = None
    if "fusecmp" in flags:
        # the 17-pass schedule (AUTO for unpreconditioned BiCGSTAB) against the 19-pass one on the same ranks: same bits
        s2, x2 = solve_with(pps.FUSE_SPLIT)
        same = bool(np.array_equal(s.history(), s2.history()) and np.array_equal(x, x2) and s.iterations == s2.iterations)
        names = [k["name"] for k in s.kernel_stats()]
        same = same and any(n.startswith("fused_s") for n in names)
        f = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        fused_equals_split = bool(int(f.item()))
        s2.close()

    # gather the per-rank data-range blocks on rank 0
    mine = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    ok = True
    if rank == 0:
        o.solve()
        ho, hs = o.history(), s.history()
        n10, n20 = min(11, len(ho), len(hs)), min(21, len(ho), len(hs))
        d10 = float(np.max(np.abs(hs[:n10] - ho[:n10]) / ho[:n10]))
        d20 = float(np.max(np.abs(hs[:n20] - ho[:n20]) / ho[:n20]))
        blocks = [(list(o.block(r).nlocal), list(o.block(r).loc)) for r in range(world)]
        xs = H.assemble_global([p.cpu().numpy() for p in parts], blocks, list(ocfg.np))
        rel = H.rel_l2(xs, H.oracle_global_solution(o))
        # guard planes after the final halo exchange (BiCGSTAB.hpp:317-321): my z- / x- guard = neighbour's data
        out = dict(layout=[px, py, pz], flags=flags, iters=s.iterations, iters_oracle=o.iters, norm_b_rel=abs(s.norm_b - o.norm_b) / o.norm_b,
                   hist10=d10, hist20=d20, sol_rel_l2=rel, true_residual=s.error_operator, seconds=s.solver_seconds)
        out["fused_equals_split"] = fused_equals_split
        # same bars as tests/test_gpu_parity.py (history 1e-11 / 1e-7, iterations inside the reference's own 15 % band)
        ok = (out["norm_b_rel"] <= 1e-13 and d10 <= 1e-11 and d20 <= 1e-7 and rel <= 2e-6 and
              0.85 * o.iters - 2 <= s.iterations <= 1.15 * o.iters + 2 and s.error_operator < 1.5e-8 and
              fused_equals_split is not False)
        if "nbg" in flags:
            # capped nested solves: lock-step over the first entries, iteration counts scatter (tests/test_gpu_next.py)
            n4 = min(4, len(ho), len(hs))
            out["hist4"] = float(np.max(np.abs(hs[:n4] - ho[:n4]) / ho[:n4]))
            out["nested_iterations"] = s.preconditioner_iterations
            ok = (out["norm_b_rel"] <= 1e-13 and out["hist4"] <= 1e-6 and rel <= 2e-6 and s.error_operator < 1.5e-8 and
                  abs(s.iterations - o.iters) <= max(2, o.iters // 4))
        try:
            H.record_margin("multi_gpu_vs_oracle", **{k: v for k, v in out.items() if k != "flags"}, flags=list(flags))
        except Exception:
            pass
        out["ok"] = bool(ok)
        print(json.dumps(out), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    s.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()

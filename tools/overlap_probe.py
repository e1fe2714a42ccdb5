#!/usr/bin/env python3
"""How much of the halo exchange is hidden behind interior compute (north_star: >= 80 %).  Run under torchrun.

    torchrun --nproc-per-node N tools/overlap_probe.py [NX NY NZ] [iters]

Three timed variants of the same fixed-length BiCGSTAB loop on 1x1xN z-slabs (SURVEY.md section 8d "Overlap metric"):
  no_comm   PPS_DEBUG_NO_HALO=1   faces never travel (wrong numbers, timing only)
  serial    PPS_OVERLAP=0         exchange, then the whole operator, on one stream
  overlap   PPS_OVERLAP=1         exchange (NCCL send/recv) on the halo stream while the interior box is computed
  p2p       + PPS_HALO_P2P=1      the default: the same schedule with the faces pushed into the neighbours' guard planes by copy engines
hidden = 1 - (t_overlap - t_no_comm) / (t_serial - t_no_comm), times = device loop time per iteration, max over ranks.
Optional extra variants: --p2p (PPS_HALO_P2P=1: CUDA-IPC peer pushes on copy engines, with the 3-stream schedule and with the
single-launch in-kernel wait PPS_OVERLAP=3), --inkernel (PPS_OVERLAP=2).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import parallelpoissonsolver_b200 as pps  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    args = [int(a) for a in sys.argv[1:] if not a.startswith("--")]
    npglobal = tuple(args[:3]) if len(args) >= 3 else (1024, 1024, 1024)
    iters = args[3] if len(args) >= 4 else 200
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    D = bench.Dist(rank, world, "cuda")
    X, B = bench.manufactured_slab(npglobal, world, rank)
    out = {"npglobal": npglobal, "world": world, "iters": iters}
    base = {"PPS_DEBUG_NO_HALO": "0", "PPS_OVERLAP": "1", "PPS_HALO_P2P": "1", "PPS_ALLREDUCE_P2P": "0"}
    variants = [("no_comm", dict(base, PPS_DEBUG_NO_HALO="1", PPS_OVERLAP="0", PPS_HALO_P2P="0")),
                ("serial", dict(base, PPS_OVERLAP="0", PPS_HALO_P2P="0")),
                ("overlap", dict(base, PPS_HALO_P2P="0")),       # three streams, faces over NCCL send/recv
                ("p2p", dict(base))]                              # three streams, faces pushed by the copy engines (the default)
    if "--p2p" in sys.argv:
        variants.append(("p2p_inkernel", dict(base, PPS_OVERLAP="3")))
        variants.append(("p2p_arp2p", dict(base, PPS_ALLREDUCE_P2P="1")))
    if "--inkernel" in sys.argv:
        variants.append(("inkernel", dict(base, PPS_OVERLAP="2", PPS_HALO_P2P="0")))
    for name, env in variants:
        os.environ.update(env)
        uid = D.bcast_bytes(pps.get_unique_id() if rank == 0 else None, 128)
        cfg = pps.make_config(npglobal, nranks=(1, 1, world), bcs=(0,) * 6, tolerance=1e-300, max_iter=iters, device=local)
        s = pps.PoissonSolver(cfg, rank=rank, world_size=world, unique_id=uid)
        s.set_fields(rank, X, B)
        s.save_fields()
        ts = []
        for rep in range(3):
            s.restore_fields()
            D.barrier()
            s.solve()
            D.barrier()
            ts.append(D.max(s.loop_seconds) / max(1, s.iterations))
        out[name + "_ms_per_iter"] = float(np.min(ts[1:])) * 1e3
        out[name + "_iters"] = s.iterations
        s.close()
    tn, tsr, to = out["no_comm_ms_per_iter"], out["serial_ms_per_iter"], out["overlap_ms_per_iter"]
    out["exposed_comm_ms_serial"] = tsr - tn
    out["exposed_comm_ms_overlap"] = to - tn
    out["hidden_fraction"] = 1 - (to - tn) / (tsr - tn) if tsr > tn else None
    for extra in ("p2p", "p2p_inkernel", "p2p_arp2p", "inkernel"):
        if extra + "_ms_per_iter" in out:
            out["hidden_fraction_" + extra] = 1 - (out[extra + "_ms_per_iter"] - tn) / (tsr - tn) if tsr > tn else None
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

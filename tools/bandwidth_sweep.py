#!/usr/bin/env python3
"""BASELINE.json configs[4]: operator-apply-only bandwidth sweep, y = A x on one GPU, 64^3 up to the largest cube
that fits, GB/s against the HBM roofline (16 algorithmic bytes per cell: read x, write y).  JSON lines on stdout.

    python tools/bandwidth_sweep.py [sizes ...]                      one GPU (operator-only handles: 3 vectors per size)
    torchrun --nproc-per-node N tools/bandwidth_sweep.py [sizes ...]  weak scaling: every rank owns an n^3 block of a 1x1xN
                                                                       slab decomposition; each apply is preceded by the z-face exchange
The largest cube that fits: 3 vectors * 8 B * (n+17)(n+2)^2 <= ~170 GB  ->  n ~ 1850; with x and y only (no fused dot) 2176^3 = 165 GB.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import parallelpoissonsolver_b200 as pps  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 192, 256, 384, 512, 640, 768, 1024, 1280, 1536, 1792, 2048, 2176]
    peak, src = bench.measured_peak()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    D = bench.Dist(rank, world, "cuda")
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for n in sizes:
        uid = D.bcast_bytes(pps.get_unique_id() if rank == 0 else None, 128) if world > 1 else None
        two_vectors = 3 * 8 * (n + 17) * (n + 2) ** 2 > 165e9      # beyond ~1850^3 only x and y fit: no fused-dot variant
        try:
            s = pps.PoissonSolver(pps.make_config((n, n, n * world), nranks=(1, 1, world), device=local,
                                                  flags=pps.FLAG_OPERATOR_ONLY | (pps.FLAG_NO_DOT_VECTOR if two_vectors else 0)),
                                  rank=rank, world_size=world, unique_id=uid)
        except pps.PpsError as e:
            print(json.dumps(dict(n=n, error=str(e)[:200])), flush=True)
            break
        reps = max(5, min(200, int(2e9 / n ** 3)))
        for dot in ((False,) if two_vectors else (False, True)):
            D.barrier()
            ms = D.max(s.bench_operator(reps, dot, with_halo=world > 1))
            nbytes = n ** 3 * 8 * (3 if dot else 2) * world
            if rank == 0:
                print(json.dumps(dict(kind="apply+dot" if dot else "apply", n=n, gpus=world, cells=n ** 3 * world, reps=reps, ms=ms,
                                      mlups=n ** 3 * world / ms / 1e3, gbs=nbytes / ms / 1e6, gbs_per_gpu=nbytes / ms / 1e6 / world,
                                      frac_of_peak=nbytes / ms / 1e6 / world / peak, peak_gbs=peak, peak_source=src,
                                      note="arrays smaller than the 126 MB L2 are served from L2 between repetitions" if n ** 3 * 16 < 126e6 else "")),
                      flush=True)
        s.close()


if __name__ == "__main__":
    main()

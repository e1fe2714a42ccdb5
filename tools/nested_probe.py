#!/usr/bin/env python3
"""Timing probe of the preconditioner stacks on ONE GPU hosting 1x1x2 blocks (shipped default problem, 128x128x256, mixed BCs):
block-Jacobi Chebyshev(11), block-local nested BiCGSTAB, GLOBAL nested BiCGSTAB (communicationON in the preconditioner slot),
the nested solves capped at 8 iterations.  Prints one JSON line per stack."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallelpoissonsolver_b200 as pps  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests import helpers as H  # noqa: E402

ocfg0 = po.make_config((128, 128, 256), (1, 1, 2), bcs=(0, 1, 0, 1, 0, 1))
o = po.Oracle(ocfg0)
o.set_problem()
STACKS = [("none", dict(precond=pps.PRECOND_NONE)),
          ("chebyshev11_block_jacobi", dict(precond=pps.PRECOND_CHEBYSHEV)),
          ("nested_bicgstab_local_cap8", dict(precond=pps.PRECOND_BICGSTAB_LOCAL, precond_max_iter=8)),
          ("nested_bicgstab_global_cap8", dict(precond=pps.PRECOND_BICGSTAB_LOCAL, precond_max_iter=8, precond_communication=1))]
for name, kw in STACKS:
    s = pps.PoissonSolver(H.pps_config_from_oracle(ocfg0, **kw))
    H.hand_over_problem(o, s)
    s.solve()
    print(json.dumps({"stack": name, "iterations": s.iterations, "nested_iterations": s.preconditioner_iterations,
                      "solver_seconds": s.solver_seconds, "true_residual": s.error_operator, "launches": s.launch_count}), flush=True)
    s.close()

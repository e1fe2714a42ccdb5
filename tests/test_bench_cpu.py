"""CPU tests of bench.py's host logic: the parity gate (first 20 iterations against the reference's fixture, per-rank slab of the
sub-lattice sample) and the labels of the reference arm.  No GPU, no library call: the solver is a stand-in that replays the fixture."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


class _Out:
    def __init__(self, a):
        self.a = a

    def numpy(self):
        return self.a


class _ReplaySolver:
    """hands back the fixture's history and, for its slab, a field whose sub-lattice equals the fixture's sample"""

    def __init__(self, g, npglobal, rank, world, hist_scale=1.0, x_bias=0.0):
        self.g, self.np_, self.rank, self.world = g, npglobal, rank, world
        self.hist_scale, self.x_bias = hist_scale, x_bias
        self.norm_b = float(g["norm_b"])
        self.iterations = 20

    def set_max_iterations(self, n):
        assert n == 20

    def restore_fields(self):
        pass

    def solve(self):
        pass

    def history(self):
        h = np.array(self.g["history"], dtype=float)
        h[5:] *= self.hist_scale
        return h

    def get_solution(self, my, outh):
        nx, ny, nzg = self.np_
        nz = nzg // self.world
        k0 = self.rank * nz
        s = int(self.g["x_stride"])
        x = outh.a
        x[:] = 0
        for k in range(k0, k0 + nz):
            if k % s == 0:
                x[1 + k - k0, 1:-1:s, 1:-1:s] = self.g["x_sample"][k // s] + self.x_bias


@pytest.mark.parametrize("world", [1, 8])
def test_parity_gate_accepts_the_fixture_and_rejects_a_perturbed_run(world):
    npglobal = (512, 512, 512) if world == 1 else (1024, 1024, 1024)     # 0.5 GB for the stand-in field; the 8-rank case checks indices only
    g = np.load(os.path.join(ROOT, "tests", "golden", bench.GOLDEN_FOR[npglobal] + ".npz"))
    nz = npglobal[2] // world

    def run(rank, **kw):
        outh = _Out(np.zeros((nz + 2, npglobal[1] + 2, npglobal[0] + 2), dtype=np.float32))   # float32 keeps the test small
        s = _ReplaySolver(g, npglobal, rank, world, **kw)
        # rank_sum over the (sequentially simulated) ranks: first pass collects, second pass uses the totals
        return s, outh

    # single simulated rank is enough to exercise the slab indexing: rank r sees only its planes of the sample
    if world == 1:
        s, outh = run(0)
        res = bench.parity_gate(s, npglobal, 0, 1, 0, lambda v: v, outh)
        assert res["ok"] and res["hist_rel_it10"] == 0 and res["x_sample_rel_l2"] < 1e-6    # float32 storage of the stand-in
        s, outh = run(0, hist_scale=1 + 1e-6)
        with pytest.raises(SystemExit):
            bench.parity_gate(s, npglobal, 0, 1, 0, lambda v: v, outh)
    else:
        # every rank's slab contributes a disjoint, non-empty part of the 32^3 sample (stride 32, 128 planes per rank -> 4 planes)
        s_ = int(g["x_stride"])
        planes = [[k for k in range(r * nz, (r + 1) * nz) if k % s_ == 0] for r in range(world)]
        assert all(len(p) == 4 for p in planes) and sum(len(p) for p in planes) == g["x_sample"].shape[0]


def test_reference_arm_line_says_what_it_ran(monkeypatch, capsys):
    monkeypatch.setattr(bench, "run_reference_sample", lambda npg: dict(value=600.0, unit="MLUP/s", cores=16, kind="reference", seconds=1.8, iters=8,
                                                                        grid="512x512x512", ranks="1x1x16", sample="stub"))
    monkeypatch.setattr(bench, "_REAL_STDOUT", None)

    class A:
        gpus, steps, warmup = 8, 2, 1
    bench.reference_arm(A, (1024, 1024, 1024), 0)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["sample_of"] == "1024x1024x1024"
    assert line["sample_ran"]["grid"] == "512x512x512" and line["sample_ran"]["iterations"] == 8 and line["sample_ran"]["to_convergence"] is False
    assert line["config"] == bench.workload_config((1024, 1024, 1024), 8)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["cores"] == 16
    bench.reference_arm(A, (1024, 1024, 1024), 3)          # other ranks print nothing
    assert capsys.readouterr().out == ""

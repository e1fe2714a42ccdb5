"""CPU test of the result-file writers (include/reference_compat/output.hpp): the formats of the reference's alpaka
driver -- residualHistory.txt (iterativeSolverBaseAlpaka.hpp:620-638) and solution.dat (src/main.cpp:135-146)."""
import os
import subprocess
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_writers(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(textwrap.dedent('''
        #include "output.hpp"
        #include <thread>
        #include <vector>
        int main() {
            const double hist[6] = {1.0, 0.5, 0.25, 1e-9, 0.0, 7.0};          // stops at the first non-positive entry
            pps_compat::write_residual_history("residualHistory.txt", 1.25, 3, 66, hist, 6, 1700);
            std::vector<std::thread> th;                                      // 4 "ranks" write their blocks concurrently
            for (int r = 0; r < 4; r++) th.emplace_back([r] {
                std::vector<double> blk(10, double(r) + 0.5);
                pps_compat::write_solution_block("solution.dat", r, 10, blk.data());
            });
            for (auto& t : th) t.join();
            return 0;
        }
    '''))
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", "-I" + os.path.join(ROOT, "include", "reference_compat"), str(src), "-o", str(exe)],
                   check=True)
    subprocess.run([str(exe)], check=True, cwd=tmp_path, capture_output=True)
    lines = (tmp_path / "residualHistory.txt").read_text().split()
    assert [float(v) for v in lines] == [1.25, 3, 66, 1.0, 0.5, 0.25, 1e-9]
    sol = np.fromfile(tmp_path / "solution.dat")
    assert sol.shape == (40,)
    assert np.array_equal(sol.reshape(4, 10), np.repeat(np.arange(4) + 0.5, 10).reshape(4, 10))


def test_writers_against_the_files_of_the_unmodified_alpaka_tree(tmp_path):
    """tests/golden/alpaka/files_*.npz holds residualHistory.txt and solution.dat exactly as the reference's alpaka tree wrote them
    (its own src/main.cpp with writeResidual / writeSolution = true, built by oracle/build_ref_alpaka.py, 1x1x2 ranks) next to the
    full-precision history and the per-rank blocks of the same solve.  Fed with those numbers, the writers of output.hpp must
    produce the same bytes (the first line is the solve time: re-emitted from the parsed value)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "alpaka", "files_alp_files_m24_112.npz"))
    want_txt = bytes(g["residual_history_txt"]).decode()
    seconds = float(want_txt.split("\n")[0])
    hist = g["history"]
    blocks = g["blocks"]
    ntot = blocks.shape[1]
    src = tmp_path / "w.cpp"
    src.write_text(textwrap.dedent('''
        #include "output.hpp"
        #include <cstdio>
        #include <vector>
        int main() {
            const double hist[] = {%s};
            pps_compat::write_residual_history("residualHistory.txt", %s, %d, %d, hist, %d, %d);
            std::vector<double> blk(%d);
            FILE* f = std::fopen("blocks.bin", "rb");
            for (int r = %d - 1; r >= 0; r--) {                       // any order: the offsets are disjoint
                std::fseek(f, long(sizeof(double)) * %d * r, SEEK_SET);
                if (std::fread(blk.data(), sizeof(double), blk.size(), f) != blk.size()) return 1;
                pps_compat::write_solution_block("solution.dat", r, %d, blk.data());
            }
            return 0;
        }
    ''') % (", ".join(float(v).hex() for v in hist), float(seconds).hex(), int(g["iters"]), int(g["precond_iters"]), len(hist), int(g["max_iter"]),
            ntot, blocks.shape[0], ntot, ntot))
    blocks.tofile(tmp_path / "blocks.bin")
    exe = tmp_path / "w"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include", "reference_compat"), str(src), "-o", str(exe)], check=True)
    subprocess.run([str(exe)], check=True, cwd=tmp_path, capture_output=True)
    assert (tmp_path / "residualHistory.txt").read_text() == want_txt
    # solution.dat: block r at byte offset r * ntot * 8 (src/main.cpp:139-144).  The reference writes fieldX after
    # checkSolutionLocalGlobal has refreshed its guard cells, the dump harness before: compare the data range.
    ours = np.fromfile(tmp_path / "solution.dat")
    assert ours.shape == g["solution_dat"].shape
    n = [int(v) for v in g["np"]]
    shp = (blocks.shape[0], n[2] // blocks.shape[0] + 2, n[1] + 2, n[0] + 2)
    inner = (slice(None), slice(1, -1), slice(1, -1), slice(1, -1))
    assert np.array_equal(ours.reshape(shp)[inner], g["solution_dat"].reshape(shp)[inner])
    # the per-phase report of the alpaka driver (solverSetup.hpp:247-268): the aggregate rows the drop-in driver prints too
    labels = [str(v) for v in g["report_labels"]]
    assert labels[-6:] == ["timeTotPreconditionerTot", "timeTotCommunicationTot", "timeTotAllReductionTot", "timeTotKernelsBBiCGstabTot",
                           "timeTotResetNeumanBCs", "timeTotal"]
    compat = open(os.path.join(ROOT, "include", "reference_compat", "iterativeSolverBase.hpp")).read()
    for lab in labels[-6:]:
        assert lab in compat, lab


def test_report_lines_match_the_archived_reference_log(tmp_path):
    """include/reference_compat/report.hpp must reproduce the reference's stdout byte for byte: checked against the lines
    of the run the reference archives (solverPoissonMPI_CPU/run/solverScoreP.o:2-8,25,28-31; 4x4x4 ranks, 128x128x256)."""
    src = tmp_path / "r.cpp"
    src.write_text(textwrap.dedent('''
        #include "report.hpp"
        int main() {
            pps_compat::RunGeometry g{3, {4, 4, 4}, {128, 128, 256}, {32, 32, 64}, {34, 34, 66}, {1, 1, 1}, {0, 0, 0}, {0.1, 0.1, 0.1},
                                      {0, 1, 0, 1, 0, 1}, 65536, 76296};
            pps_compat::print_banner(g, 0);
            pps_compat::print_result(155, 7.55263e-09, 7.55263e-09, 128LL * 128 * 256);
            pps_compat::print_timings(614.681, 614.06, 615.756);
            return 0;
        }
    '''))
    exe = tmp_path / "r"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include", "reference_compat"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()
    expected = [
        "Domain DIM = 3 - Number of MPI tasks 4 4 4 - Tot MPI ranks 64 - Max threads per MPI rank 1 - Tot threads 64",
        "Global grid size from block 128 128 256 - Global number of points 4194304",
        "Domain local Np xyz no guards 32 32 64 - Domain local Np xyz guards = 34 34 66 - Guards size 1 1 1",
        "Total local number of points noguards 65536 - total local number of points guards 76296",
        "Total local number of points noguards per thread 65536 - total local number of points guards per thread 76296",
        "Domain global origin xyz 0 0 0 - domain global extension xyz 12.7 12.7 25.5 - Ds xyz  = 0.1 0.1 0.1",
        "Boundary condition type 0 1 0 1 0 1",
        "Iterative solver finished with iter: 155 error from algo 7.55263e-09 error r=b-Ax 7.55263e-09 errorAvgtot 1.80069e-15",
        "Solver time: 614.681 seconds",
        "SolverInFunction time: 614.06 seconds",
        "Elapsed time: 615.756 seconds",
        "End program. ",
    ]
    assert out[0].startswith("Current local time and date: ")
    assert out[1:] == expected

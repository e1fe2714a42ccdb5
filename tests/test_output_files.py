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

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

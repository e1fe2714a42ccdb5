// output.hpp (reference_compat) -- the two result files the reference's alpaka driver can write
// (solverPoissonMPI_alpaka/include/inputParam.hpp:39-47 switches writeResidual / writeSolution):
//   residualHistory.txt  iterativeSolverBaseAlpaka.hpp:620-638: solve seconds, main iterations, preconditioner
//                        iterations (both applications per cycle counted), then one residual norm per line
//   solution.dat         src/main.cpp:135-146: raw fp64, the guard-padded block of rank r at byte offset r * ntot * 8
// Host-only helpers; SolverAdapter::writeResidualHistory() and the drivers use them.
#pragma once

#include <fcntl.h>
#include <unistd.h>

#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

namespace pps_compat {

inline bool write_residual_history(const std::string& path, double solve_seconds, int iterations, int precond_iterations,
                                   const double* history, int n_history, int max_iteration) {
    std::ofstream out(path);
    if (!out) {
        std::cerr << "Error: Could not open the file!" << std::endl;
        return false;
    }
    out << solve_seconds << "\n" << iterations << "\n" << precond_iterations << "\n";
    for (int i = 0; i < n_history && i < max_iteration && history[i] > 0; i++) out << history[i] << "\n";
    out.close();
    std::cout << "residualHistory has been written to " << path << std::endl;
    return true;
}

// every rank writes its own block; offsets are disjoint, so ranks (threads or processes) need no ordering
inline bool write_solution_block(const std::string& path, int rank, long long ntot_local_guards, const double* field) {
    const int fd = ::open(path.c_str(), O_CREAT | O_WRONLY, 0644);
    if (fd < 0) {
        std::cerr << "Error: Could not open " << path << std::endl;
        return false;
    }
    const size_t bytes = sizeof(double) * static_cast<size_t>(ntot_local_guards);
    const char* p = reinterpret_cast<const char*>(field);
    size_t done = 0;
    off_t off = static_cast<off_t>(bytes) * rank;
    while (done < bytes) {
        const ssize_t w = ::pwrite(fd, p + done, bytes - done, off + static_cast<off_t>(done));
        if (w <= 0) {
            ::close(fd);
            return false;
        }
        done += static_cast<size_t>(w);
    }
    ::close(fd);
    return true;
}

}  // namespace pps_compat

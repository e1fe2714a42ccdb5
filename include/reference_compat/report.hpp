// report.hpp (reference_compat) -- the stdout lines of the reference driver, in one place, so that log scrapers written
// for solverPoissonMPI_CPU keep working: banner (main.cpp:64-74), result line (:105-109), timings (:120-126).
// The solver adapters print the START / every-10-iterations lines themselves (iterativeSolverBase.hpp here).
#pragma once

#include <array>
#include <ctime>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>

namespace pps_compat {

template <typename T, size_t N>
inline std::string joined(const std::array<T, N>& a, size_t n = N) {
    std::ostringstream os;
    for (size_t i = 0; i < n; i++) os << (i ? " " : "") << a[i];
    return os.str();
}

struct RunGeometry {
    int dim;
    std::array<int, 3> ranks, npglobal, nlocal, nlocalGuards, guards;
    std::array<double, 3> origin, ds;
    std::array<int, 6> bcs;
    long long ntotLocal, ntotLocalGuards;
};

inline void print_banner(const RunGeometry& g, std::time_t when) {
    const int total = g.ranks[0] * g.ranks[1] * g.ranks[2];
    std::array<double, 3> extent{};
    for (int d = 0; d < 3; d++) extent[d] = g.origin[d] + (g.npglobal[d] - 1) * g.ds[d];
    std::ostream& o = std::cout;
    o << "Current local time and date: " << std::put_time(std::localtime(&when), "%Y-%m-%d %H:%M:%S") << "\n";
    o << "Domain DIM = " << g.dim << " - Number of MPI tasks " << joined(g.ranks) << " - Tot MPI ranks " << total
      << " - Max threads per MPI rank " << 1 << " - Tot threads " << total << "\n";
    o << "Global grid size from block " << joined(g.npglobal) << " - Global number of points " << total * g.ntotLocal << "\n";
    o << "Domain local Np xyz no guards " << joined(g.nlocal) << " - Domain local Np xyz guards = " << joined(g.nlocalGuards)
      << " - Guards size " << joined(g.guards) << "\n";
    for (const char* per : {"", " per thread"})
        o << "Total local number of points noguards" << per << " " << g.ntotLocal << " - total local number of points guards" << per << " "
          << g.ntotLocalGuards << "\n";
    o << "Domain global origin xyz " << joined(g.origin) << " - domain global extension xyz " << joined(extent) << " - Ds xyz  = "
      << joined(g.ds) << "\n";
    o << "Boundary condition type " << joined(g.bcs) << std::endl;
}

inline void print_result(int iterations, double errAlgo, double errTrue, long long globalPoints) {
    std::cout << "Iterative solver finished with iter: " << iterations << " error from algo " << errAlgo << " error r=b-Ax " << errTrue
              << " errorAvgtot " << errTrue / static_cast<double>(globalPoints) << std::endl;
}

inline void print_timings(double solver, double inFunction, double elapsed) {
    const char* label[3] = {"Solver time: ", "SolverInFunction time: ", "Elapsed time: "};
    const double value[3] = {solver, inFunction, elapsed};
    for (int i = 0; i < 3; i++) std::cout << label[i] << value[i] << " seconds" << std::endl;
    std::cout << "End program. " << std::endl;
}

}  // namespace pps_compat

// ref_dump_alpaka.cpp -- golden-data extractor around the UNMODIFIED alpaka tree of the reference
// (solverPoissonMPI_alpaka) with alpaka's OpenMP-blocks CPU accelerator.  TEST INFRASTRUCTURE ONLY.
//
// Compiled by oracle/build_ref_alpaka.py against a per-config copy of that tree's inputParam.hpp / solverSetup.hpp
// (only constants and typedefs edited: the reference's own configuration mechanism), the other headers and the
// vendored alpaka 1.2.0 + mdspan where they lie, oracle/boost_shim (predef / demangle only) and the
// threads-as-ranks oracle/mpi_shim.  It pins the alpaka-only configuration surface of SURVEY.md section 8 (f1):
// T_data_chebyshev = float (solverSetup.hpp:14) and the `local` eigenvalue switch (inputParam.hpp:21-22,27).
//
//   mode "precond": apply the preconditioner class under test (PPS_ALP_PRECOND, e.g. T_PreconditionerChebGlobal)
//                   ONCE to a deterministic field B (integer hash of the global indices, solver range only) --
//                   ChebyshevIterationAlpaka::operator()(bufX, bufB), chebyshevIterationAlpaka.hpp:119-310.
//                     rank<r>.pb / .px   raw fp64, guard-padded: B after the call (its Neumann ghosts are reset), X
//   mode "solve":   what src/main.cpp:83-101 does: T_Solver on zeroed arrays (the solver sets the problem itself)
//                     rank<r>.x, history.bin, summary.txt as oracle/ref_dump.cpp writes them
//   both:           rank<r>.meta (geometry, eigenvalue bounds)
// usage: ref_dump_alpaka px py pz outdir precond|solve
#include <mpi.h>

#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <alpaka/alpaka.hpp>

#include "inputParam.hpp"
#include "communicationMPI.hpp"
#include "blockGrid.hpp"
#include "alpakaHelper.hpp"

#ifndef PPS_ALP_PRECOND
#    define PPS_ALP_PRECOND T_PreconditionerChebGlobal
#endif

namespace {

struct Probe : public T_Solver {
    using T_Solver::T_Solver;
    const T_data* history() const { return this->errorFromIterationHistory_; }
    T_data normB() const { return this->normFieldB_; }
};

void writeRaw(const std::string& path, const T_data* p, size_t n) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(p), static_cast<std::streamsize>(n * sizeof(T_data)));
}

// the test field of "precond" mode; tests/test_oracle.py (alpaka_test_field) builds the same numbers with numpy.
// Exact in fp64 (a 12-bit and a 10-bit integer scaled by powers of two), NOT representable in fp32: the cast rounds.
double testField(long gi, long gj, long gk) {
    const long h1 = (gi * 73856093L + gj * 19349663L + gk * 83492791L) % 4001L;
    const long h2 = (gi * 2654435761L + gj * 40503L + gk * 9973L) % 1021L;
    return static_cast<double>(h1 - 2000L) / 4096.0 + static_cast<double>(h2) / 1099511627776.0;  // 2^-40
}

void writeMeta(const std::string& base, const BlockGrid<DIM, T_data>& grid) {
    std::ofstream m(base + ".meta");
    auto ng = grid.getNlocalGuards();
    auto nn = grid.getNlocalNoGuards();
    auto loc = grid.getGlobalLocation();
    auto ld = grid.getIndexLimitsData();
    auto ls = grid.getIndexLimitsSolver();
    auto hb = grid.getHasBoundary();
    auto hc = grid.getHasCommunication();
    m << "nlocal_guards " << ng[0] << " " << ng[1] << " " << ng[2] << "\n";
    m << "nlocal_noguards " << nn[0] << " " << nn[1] << " " << nn[2] << "\n";
    m << "global_location " << loc[0] << " " << loc[1] << " " << loc[2] << "\n";
    m << "limits_data";
    for (int v : ld) m << " " << v;
    m << "\nlimits_solver";
    for (int v : ls) m << " " << v;
    m << "\nhas_boundary";
    for (bool v : hb) m << " " << int(v);
    m << "\nhas_comm";
    for (bool v : hc) m << " " << int(v);
    m.precision(17);
    m << "\neig_global " << grid.getEigenValuesGlobal()[0] << " " << grid.getEigenValuesGlobal()[1];
    m << "\neig_local " << grid.getEigenValuesLocal()[0] << " " << grid.getEigenValuesLocal()[1] << "\n";
}

int dumpMain(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    int world = 1, rank = 0;
    MPI_Comm_size(MPI_COMM_WORLD, &world);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    const std::array<int, 3> nranks = {std::atoi(argv[1]), std::atoi(argv[2]), std::atoi(argv[3])};
    const std::string out = argv[4];
    const std::string mode = argv[5];

    BlockGrid<DIM, T_data> grid(nranks, rank, npglobal, ds, origin, guards, bcsType, bcsValue);
    CommunicatorMPI<DIM, T_data> comm(grid);
    ExactSolutionAndBCs<DIM, T_data> exact;
    AlpakaHelper<DIM, T_data> helper(grid);
    const size_t ntot = static_cast<size_t>(grid.getNtotLocalGuards());
    const std::string base = out + "/rank" + std::to_string(rank);
    writeMeta(base, grid);

    if (mode == "precond") {
        PPS_ALP_PRECOND precond(grid, exact, comm, helper);
        auto bufX = alpaka::allocBuf<T_data, Idx>(helper.devAcc_, helper.extent_);
        auto bufB = alpaka::allocBuf<T_data, Idx>(helper.devAcc_, helper.extent_);
        T_data* x = alpaka::getPtrNative(bufX);
        T_data* b = alpaka::getPtrNative(bufB);
        std::fill(x, x + ntot, 0.0);
        std::fill(b, b + ntot, 0.0);
        const auto ng = grid.getNlocalGuards();
        const auto nn = grid.getNlocalNoGuards();
        const auto loc = grid.getGlobalLocation();
        const auto ls = grid.getIndexLimitsSolver();
        const auto gd = grid.getGuards();
        for (int k = ls[4]; k < ls[5]; k++)
            for (int j = ls[2]; j < ls[3]; j++)
                for (int i = ls[0]; i < ls[1]; i++) {
                    const long gi = static_cast<long>(loc[0]) * nn[0] + (i - gd[0]);
                    const long gj = static_cast<long>(loc[1]) * nn[1] + (j - gd[1]);
                    const long gk = static_cast<long>(loc[2]) * nn[2] + (k - gd[2]);
                    b[i + static_cast<size_t>(ng[0]) * (j + static_cast<size_t>(ng[1]) * k)] = testField(gi, gj, gk);
                }
        precond(bufX, bufB);
        writeRaw(base + ".pb", b, ntot);
        writeRaw(base + ".px", x, ntot);
        MPI_Barrier(MPI_COMM_WORLD);
        if (rank == 0) {
            std::ofstream s(out + "/summary.txt");
            s << "world " << world << "\n";
            s << "precond_iters " << precond.getNumIterationFinal() << "\n";
        }
        MPI_Finalize();
        return 0;
    }

    Probe solver(grid, exact, comm, helper);
    std::vector<T_data> x(ntot, 0), b(ntot, 0);
    auto t0 = std::chrono::high_resolution_clock::now();
    solver(x.data(), b.data());
    auto t1 = std::chrono::high_resolution_clock::now();
    MPI_Barrier(MPI_COMM_WORLD);
    writeRaw(base + ".x", x.data(), ntot);
    if (rank == 0) {
        const int iters = solver.getNumIterationFinal();
        writeRaw(out + "/history.bin", solver.history(), static_cast<size_t>(iters) + 1);
        std::ofstream s(out + "/summary.txt");
        s.precision(17);
        s << "world " << world << "\n";
        s << "iters " << iters << "\n";
        s << "precond_iters " << solver.getNumIterationPreconditionerFinal() << "\n";
        s << "error_iteration " << solver.getErrorFromIteration() << "\n";
        s << "error_operator " << solver.getErrorComputeOperator() << "\n";
        s << "norm_b " << solver.normB() << "\n";
        s << "tolerance " << static_cast<T_data>(tollMainSolver) * tollScalingFactor << "\n";
        s << "max_iter " << iterMaxMainSolver << "\n";
        s << "solver_seconds " << std::chrono::duration<double>(t1 - t0).count() << "\n";
    }
    solver.checkSolutionLocalGlobal(x.data());
    MPI_Finalize();
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s px py pz outdir precond|solve\n", argv[0]);
        return 2;
    }
    const int world = std::atoi(argv[1]) * std::atoi(argv[2]) * std::atoi(argv[3]);
    return pps_shim_run(world, dumpMain, argc, argv);
}

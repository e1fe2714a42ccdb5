// ref_dump.cpp -- golden-data extractor around the UNMODIFIED reference classes.
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/build_ref.py against a per-config copy of
// the reference headers (oracle/_ref/cfg/<name>/, generated, never committed) and the
// threads-as-ranks mpi.h.  It repeats what solverPoissonMPI_CPU/src/main.cpp:58-99 does
// (BlockGrid, CommunicatorMPI, ExactSolutionAndBCs, MatrixFreeOperatorA, T_Solver,
// setProblem, solve) and writes the arrays the reference never writes to disk:
//   rank<r>.meta           text: geometry of the block
//   rank<r>.x0 / .b0       raw fp64, guard-padded, after setProblem (solver inputs)
//   rank<r>.x  / .b        raw fp64, guard-padded, after the solve
//   history.bin            rank 0: errorFromIterationHistory_[0..iters]
//   summary.txt            rank 0: iters, errors, norm of the BC-adjusted b, timings
// usage: ref_dump px py pz outdir
#include <mpi.h>

#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "inputParam.hpp"
#include "communicationMPI.hpp"
#include "solvers.hpp"
#include "blockGrid.hpp"
#include "matrixFreeOperatorA.hpp"

namespace {

// exposes the protected bits of the reference solver without touching its sources
struct Probe : public T_Solver {
    using T_Solver::T_Solver;
    const T_data* history() const { return this->errorFromIterationHistory_; }
    T_data normOfAdjustedB(T_data* x, T_data* b) {
        return this->template normalizeProblemToFieldBNorm<true, true>(x, b);
    }
};

void writeRaw(const std::string& path, const T_data* p, size_t n) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(p), static_cast<std::streamsize>(n * sizeof(T_data)));
}

int dumpMain(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    int world = 1, rank = 0;
    MPI_Comm_size(MPI_COMM_WORLD, &world);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    const std::array<int, 3> nranks = {std::atoi(argv[1]), std::atoi(argv[2]), std::atoi(argv[3])};
    const std::string out = argv[4];

    BlockGrid<DIM, T_data> grid(nranks, rank, npglobal, ds, origin, guards, bcsType, bcsValue);
    CommunicatorMPI<DIM, T_data> comm(grid);
    ExactSolutionAndBCs<DIM, T_data> exact;
    MatrixFreeOperatorA<DIM, T_data> opA(grid);
    Probe solver(grid, exact, comm);

    const size_t ntot = static_cast<size_t>(grid.getNtotLocalGuards());
    std::vector<T_data> x(ntot, 0), b(ntot, 0);
    solver.setProblem(x.data(), b.data());

    const std::string base = out + "/rank" + std::to_string(rank);
    {
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
    // PPS_DUMP_LIGHT=1 (large grids): only the final x is written, not the inputs and not b
    const bool light = std::getenv("PPS_DUMP_LIGHT") != nullptr;
    if (!light) {
        writeRaw(base + ".x0", x.data(), ntot);
        writeRaw(base + ".b0", b.data(), ntot);
    }

    // norm of the BC-adjusted right-hand side, computed on copies (collective call)
    T_data normB = 1;
#ifndef PPS_DUMP_SKIP_NORM
    // (ChebyshevIteration as main solver never normalises and never resets normFieldB_: the probe call would leave a norm
    //  != 1 behind and change its Neumann ghosts, so build_ref.py defines PPS_DUMP_SKIP_NORM for that solver)
    {
        std::vector<T_data> xc(x), bc(b);
        normB = solver.normOfAdjustedB(xc.data(), bc.data());
    }
#endif

    auto t0 = std::chrono::high_resolution_clock::now();
    solver(x.data(), b.data(), opA);
    auto t1 = std::chrono::high_resolution_clock::now();
    MPI_Barrier(MPI_COMM_WORLD);

    writeRaw(base + ".x", x.data(), ntot);
    if (!light) writeRaw(base + ".b", b.data(), ntot);

    if (rank == 0) {
        const int iters = solver.getNumIterationFinal();
        writeRaw(out + "/history.bin", solver.history(), static_cast<size_t>(iters) + 1);
        std::ofstream s(out + "/summary.txt");
        s.precision(17);
        s << "world " << world << "\n";
        s << "nranks " << nranks[0] << " " << nranks[1] << " " << nranks[2] << "\n";
        s << "npglobal " << npglobal[0] << " " << npglobal[1] << " " << npglobal[2] << "\n";
        s << "ds " << ds[0] << " " << ds[1] << " " << ds[2] << "\n";
        s << "origin " << origin[0] << " " << origin[1] << " " << origin[2] << "\n";
        s << "bcs";
        for (int v : bcsType) s << " " << v;
        s << "\niters " << iters << "\n";
        s << "error_iteration " << solver.getErrorFromIteration() << "\n";
        s << "error_operator " << solver.getErrorComputeOperator() << "\n";
        s << "norm_b " << normB << "\n";
        s << "tolerance " << static_cast<T_data>(tollMainSolver) * tollScalingFactor << "\n";
        s << "max_iter " << iterMaxMainSolver << "\n";
        s << "solver_seconds " << std::chrono::duration<double>(t1 - t0).count() << "\n";
        s << "loop_seconds " << solver.getDurationSolver().count() << "\n";
    }
    // the reference's own post-solve check (prints the two "Max error local" lines)
    solver.checkSolutionLocalGlobal(x.data());
    MPI_Finalize();
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s px py pz outdir\n", argv[0]);
        return 2;
    }
    const int world = std::atoi(argv[1]) * std::atoi(argv[2]) * std::atoi(argv[3]);
    return pps_shim_run(world, dumpMain, argc, argv);
}

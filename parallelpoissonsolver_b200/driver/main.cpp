// solverPoisson -- C++ host driver of the B200 path: `solverPoisson px py pz`, configured at compile time by
// inputParam.hpp / solverSetup.hpp like the reference driver it stands in for (solverPoissonMPI_CPU/src/main.cpp), printing
// the same log (include/reference_compat/report.hpp).  Differences: the ranks are threads of this process -- one per GPU, or
// all blocks on one GPU when there are more ranks than GPUs -- and the solver classes are the GPU-backed templates of
// include/reference_compat.
#include <mpi.h>   // include/reference_compat/mpi.h

#include <chrono>
#include <cstdlib>
#include <vector>

#include "inputParam.hpp"
#include "report.hpp"

namespace {

using Clock = std::chrono::high_resolution_clock;
double seconds(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

int runRank(const std::array<int, 3>& layout) {
    const auto t0 = Clock::now();
    int rank = 0;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    const bool root = rank == 0;

    BlockGrid<DIM, T_data> grid(layout, rank, npglobal, ds, origin, guards, bcsType, bcsValue);
    if (root) {
        pps_compat::RunGeometry g{DIM, layout, npglobal, grid.getNlocalNoGuards(), grid.getNlocalGuards(), guards, {}, {}, grid.getBcsType(),
                                  grid.getNtotLocalNoGuards(), grid.getNtotLocalGuards()};
        for (int d = 0; d < 3; d++) { g.origin[d] = origin[d]; g.ds[d] = ds[d]; }
        pps_compat::print_banner(g, std::chrono::system_clock::to_time_t(std::chrono::system_clock::now()));
    }

    CommunicatorMPI<DIM, T_data> halo(grid);
    ExactSolutionAndBCs<DIM, T_data> problem;
    MatrixFreeOperatorA<DIM, T_data> laplacian(grid);
    T_Solver solver(grid, problem, halo);

    std::vector<T_data> x(static_cast<size_t>(grid.getNtotLocalGuards()), 0), b(x.size(), 0);
    solver.setProblem(x.data(), b.data());

    const auto t1 = Clock::now();
    solver(x.data(), b.data(), laplacian);
    const auto t2 = Clock::now();
    MPI_Barrier(MPI_COMM_WORLD);

    if (root) pps_compat::print_result(solver.getNumIterationFinal(), solver.getErrorFromIteration(), solver.getErrorComputeOperator(),
                                       grid.getNtotNpglobal());
    solver.checkSolutionLocalGlobal(x.data());
    std::cout.flush();
    MPI_Barrier(MPI_COMM_WORLD);
    const auto t3 = Clock::now();

    // optional result files of the reference's alpaka driver (its main.cpp:124-146)
    if constexpr (writeResidual) {
        if (root) solver.writeResidualHistory();
    }
    if constexpr (writeSolution) pps_compat::write_solution_block("solution.dat", rank, grid.getNtotLocalGuards(), x.data());
    MPI_Barrier(MPI_COMM_WORLD);

    // per-phase report of the alpaka driver (its main.cpp:156), when the solve ran with phase timers (PPS_PHASE_TIMERS=1)
    if (root && std::getenv("PPS_PHASE_TIMERS") != nullptr && std::atoi(std::getenv("PPS_PHASE_TIMERS")) != 0)
        solver.timeCounter.printAverageTime(solver.getNumIterationFinal());
    if (root) pps_compat::print_timings(seconds(t1, t2), solver.getDurationSolver().count(), seconds(t0, t3));
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    std::array<int, 3> layout = {1, 1, 1};
    if (argc < 1 + DIM) {
        std::cerr << "usage: " << argv[0] << " px py pz" << std::endl;
        return 2;
    }
    for (int d = 0; d < DIM; d++) layout[d] = std::atoi(argv[1 + d]);
    const int world = layout[0] * layout[1] * layout[2];
    bool divisible = world >= 1;
    for (int d = 0; d < 3 && divisible; d++) divisible = layout[d] >= 1 && npglobal[d] % layout[d] == 0;
    if (!divisible) {
        // the reference exits on an incoherent rank grid (main.cpp:51-55) and silently drops points when the grid is not
        // divisible by the layout (blockGrid.hpp:165); we refuse both
        std::cerr << "Error: configuration of ranks not coherent! DIM = " << DIM << " ranks " << pps_compat::joined(layout) << " grid "
                  << pps_compat::joined(npglobal) << std::endl;
        return -1;
    }
    return pps_compat::run_ranks(world, [&](int) { return runRank(layout); });
}

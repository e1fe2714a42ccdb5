// solverPoisson -- C++ host driver of the B200 path.  Same command line (`solverPoisson px py pz`), same
// configuration headers and the same stdout lines as the reference driver (solverPoissonMPI_CPU/src/main.cpp);
// the ranks are threads of this process (one per GPU; or all blocks on one GPU when there are more ranks than
// GPUs) instead of MPI processes, and the solver classes are the GPU-backed templates of include/reference_compat.
#include <mpi.h>   // include/reference_compat/mpi.h

#include <array>
#include <chrono>
#include <cstdlib>
#include <ctime>
#include <iomanip>
#include <iostream>
#include <vector>

#include "inputParam.hpp"
#include "communicationMPI.hpp"
#include "solvers.hpp"
#include "blockGrid.hpp"
#include "matrixFreeOperatorA.hpp"

static int rankMain(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    int worldSize = 1, myRank = 0;
    MPI_Comm_size(MPI_COMM_WORLD, &worldSize);
    MPI_Comm_rank(MPI_COMM_WORLD, &myRank);
    const auto wallStart = std::chrono::high_resolution_clock::now();
    const std::time_t now = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());

    const std::array<int, 3> nranks = {std::atoi(argv[1]), DIM > 1 ? std::atoi(argv[2]) : 1, DIM > 2 ? std::atoi(argv[3]) : 1};
    BlockGrid<DIM, T_data> blockGrid(nranks, myRank, npglobal, ds, origin, guards, bcsType, bcsValue);
    const auto nl = blockGrid.getNlocalNoGuards();
    const auto ng = blockGrid.getNlocalGuards();
    const int total = nranks[0] * nranks[1] * nranks[2];

    if (myRank == 0) {
        std::cout << "Current local time and date: " << std::put_time(std::localtime(&now), "%Y-%m-%d %H:%M:%S") << std::endl;
        std::cout << "Domain DIM = " << DIM << " - Number of MPI tasks " << nranks[0] << " " << nranks[1] << " " << nranks[2]
                  << " - Tot MPI ranks " << total << " - Max threads per MPI rank " << 1 << " - Tot threads " << total << std::endl;
        std::cout << "Global grid size from block " << npglobal[0] << " " << npglobal[1] << " " << npglobal[2]
                  << " - Global number of points " << total * blockGrid.getNtotLocalNoGuards() << std::endl;
        std::cout << "Domain local Np xyz no guards " << nl[0] << " " << nl[1] << " " << nl[2] << " - Domain local Np xyz guards = " << ng[0]
                  << " " << ng[1] << " " << ng[2] << " - Guards size " << guards[0] << " " << guards[1] << " " << guards[2] << std::endl;
        std::cout << "Total local number of points noguards " << blockGrid.getNtotLocalNoGuards()
                  << " - total local number of points guards " << blockGrid.getNtotLocalGuards() << std::endl;
        std::cout << "Total local number of points noguards per thread " << blockGrid.getNtotLocalNoGuards()
                  << " - total local number of points guards per thread " << blockGrid.getNtotLocalGuards() << std::endl;
        std::cout << "Domain global origin xyz " << origin[0] << " " << origin[1] << " " << origin[2] << " - domain global extension xyz "
                  << origin[0] + (npglobal[0] - 1) * ds[0] << " " << origin[1] + (npglobal[1] - 1) * ds[1] << " "
                  << origin[2] + (npglobal[2] - 1) * ds[2] << " - Ds xyz  = " << ds[0] << " " << ds[1] << " " << ds[2] << std::endl;
        std::cout << "Boundary condition type";
        for (int f = 0; f < 6; f++) std::cout << " " << blockGrid.getBcsType()[f];
        std::cout << std::endl;
    }

    CommunicatorMPI<DIM, T_data> communicator(blockGrid);
    ExactSolutionAndBCs<DIM, T_data> exactSolutionAndBCs;
    MatrixFreeOperatorA<DIM, T_data> operatorA(blockGrid);
    T_Solver solver(blockGrid, exactSolutionAndBCs, communicator);

    std::vector<T_data> fieldX(static_cast<size_t>(blockGrid.getNtotLocalGuards()), 0);
    std::vector<T_data> fieldB(static_cast<size_t>(blockGrid.getNtotLocalGuards()), 0);
    solver.setProblem(fieldX.data(), fieldB.data());

    const auto solveStart = std::chrono::high_resolution_clock::now();
    solver(fieldX.data(), fieldB.data(), operatorA);
    const auto solveEnd = std::chrono::high_resolution_clock::now();
    MPI_Barrier(MPI_COMM_WORLD);

    if (myRank == 0)
        std::cout << "Iterative solver finished with iter: " << solver.getNumIterationFinal() << " error from algo "
                  << solver.getErrorFromIteration() << " error r=b-Ax " << solver.getErrorComputeOperator() << " errorAvgtot "
                  << solver.getErrorComputeOperator() / static_cast<T_data>(blockGrid.getNtotNpglobal()) << std::endl;

    solver.checkSolutionLocalGlobal(fieldX.data());
    std::cout.flush();
    MPI_Barrier(MPI_COMM_WORLD);
    const auto wallEnd = std::chrono::high_resolution_clock::now();
    // optional result files, as in the reference's alpaka driver (solverPoissonMPI_alpaka/src/main.cpp:124-146)
    if constexpr (writeResidual) {
        if (myRank == 0) solver.writeResidualHistory();
    }
    if constexpr (writeSolution) {
        pps_compat::write_solution_block("solution.dat", myRank, blockGrid.getNtotLocalGuards(), fieldX.data());
    }
    MPI_Barrier(MPI_COMM_WORLD);
    if (myRank == 0) {
        std::cout << "Solver time: " << std::chrono::duration<double>(solveEnd - solveStart).count() << " seconds" << std::endl;
        std::cout << "SolverInFunction time: " << solver.getDurationSolver().count() << " seconds" << std::endl;
        std::cout << "Elapsed time: " << std::chrono::duration<double>(wallEnd - wallStart).count() << " seconds" << std::endl;
        std::cout << "End program. " << std::endl;
    }
    MPI_Finalize();
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 1 + DIM) {
        std::cerr << "usage: " << argv[0] << " px py pz" << std::endl;
        return 2;
    }
    const int px = std::atoi(argv[1]), py = DIM > 1 ? std::atoi(argv[2]) : 1, pz = DIM > 2 ? std::atoi(argv[3]) : 1;
    const int world = px * py * pz;
    if (world < 1 || npglobal[0] % px || npglobal[1] % py || npglobal[2] % pz) {
        // the reference exits on an incoherent rank grid (main.cpp:51-55); it silently drops points when the grid
        // is not divisible (blockGrid.hpp:165) -- we refuse instead
        std::cerr << "Error: configuration of ranks not coherent! DIM = " << DIM << " ranks " << px << " " << py << " " << pz
                  << " grid " << npglobal[0] << " " << npglobal[1] << " " << npglobal[2] << std::endl;
        return -1;
    }
    return pps_compat::run_ranks(world, [&](int) { return rankMain(argc, argv); });
}

// chebyshevIteration.hpp (reference_compat) -- ChebyshevIteration<DIM, T_data, tolerance, maxIteration, isMainLoop,
// communicationON, T_Preconditioner> (chebyshevIteration.hpp:14-156).  Two roles:
//   preconditioner  !isMainLoop, communicationOFF (inputParam.hpp:28): block-Jacobi, maxIteration = chebyshevMax; a tag, never built
//   main solver     isMainLoop, communicationON (chebyshevIteration.hpp:61-67,132-139): maxIteration sweeps, no history
#pragma once
#include "iterativeSolverBase.hpp"

template <int DIM, typename T_data, int tolerance, int maxIteration, bool isMainLoop, bool communicationON, typename T_Preconditioner>
class ChebyshevIteration : public pps_compat::SolverAdapter<DIM, T_data, maxIteration> {
  public:
    static constexpr int kAsSolver = (isMainLoop && communicationON) ? static_cast<int>(PPS_SOLVER_CHEBYSHEV) : -1;
    static constexpr pps_compat::StackInfo kStack{kAsSolver, PPS_PRECOND_CHEBYSHEV, maxIteration, communicationON, tolerance, 0};
    ChebyshevIteration(const BlockGrid<DIM, T_data>& blockGrid, const ExactSolutionAndBCs<DIM, T_data>& exactSolutionAndBCs,
                       CommunicatorMPI<DIM, T_data>& communicatorMPI)
        : pps_compat::SolverAdapter<DIM, T_data, maxIteration>(blockGrid, exactSolutionAndBCs, communicatorMPI, kStack,
                                                               T_Preconditioner::kStack, tolerance, "chebyshevIteration") {
        static_assert(isMainLoop && communicationON, "as T_Solver, ChebyshevIteration needs isMainLoop and communicationON");
    }
};

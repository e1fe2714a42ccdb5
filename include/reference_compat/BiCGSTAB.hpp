// BiCGSTAB.hpp (reference_compat) -- BiCGSTAB<DIM, T_data, tolerance, maxIteration, isMainLoop, communicationON,
// T_Preconditioner> with the reference's parameter list (BiCGSTAB.hpp:14); the iteration of BiCGSTAB.hpp:131-292
// runs in libpps_b200.so.  Two roles:
//   main solver            isMainLoop, communicationON                       (inputParam.hpp:32-33)
//   nested preconditioner  !isMainLoop, NoneSolver inside  (T_Preconditioner, inputParam.hpp:31): a tag, never built.  communicationOFF = the
//                          block-local solve of inputParam.hpp:31; communicationON = a GLOBAL nested solve (face exchanges and allreduces
//                          inside the preconditioner; the alpaka tree's T_PreconditionerBiCGStabGlobal, its inputParam.hpp:33)
#pragma once
#include "iterativeSolverBase.hpp"

template <int DIM, typename T_data, int tolerance, int maxIteration, bool isMainLoop, bool communicationON, typename T_Preconditioner>
class BiCGSTAB : public pps_compat::SolverAdapter<DIM, T_data, maxIteration> {
  public:
    static constexpr int kAsSolver = (isMainLoop && communicationON) ? static_cast<int>(PPS_SOLVER_BICGSTAB) : -1;
    static constexpr int kAsPreconditioner =
        (!isMainLoop && T_Preconditioner::kStack.precond_kind == PPS_PRECOND_NONE) ? static_cast<int>(PPS_PRECOND_BICGSTAB_LOCAL) : -1;
    static constexpr pps_compat::StackInfo kStack{kAsSolver, kAsPreconditioner, maxIteration, communicationON, tolerance, 0};
    BiCGSTAB(const BlockGrid<DIM, T_data>& blockGrid, const ExactSolutionAndBCs<DIM, T_data>& exactSolutionAndBCs,
             CommunicatorMPI<DIM, T_data>& communicatorMPI)
        : pps_compat::SolverAdapter<DIM, T_data, maxIteration>(blockGrid, exactSolutionAndBCs, communicatorMPI, kStack,
                                                               T_Preconditioner::kStack, tolerance, "BiCGSTAB") {
        static_assert(isMainLoop && communicationON, "as T_Solver, BiCGSTAB needs isMainLoop and communicationON");
    }
};

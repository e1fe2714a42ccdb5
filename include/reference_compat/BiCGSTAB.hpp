// BiCGSTAB.hpp (reference_compat) -- BiCGSTAB<DIM, T_data, tolerance, maxIteration, isMainLoop, communicationON,
// T_Preconditioner> with the reference's parameter list (BiCGSTAB.hpp:14); the iteration of BiCGSTAB.hpp:131-292
// runs in libpps_b200.so.
#pragma once
#include "iterativeSolverBase.hpp"

template <int DIM, typename T_data, int tolerance, int maxIteration, bool isMainLoop, bool communicationON, typename T_Preconditioner>
class BiCGSTAB : public pps_compat::SolverAdapter<DIM, T_data, maxIteration> {
  public:
    // as a preconditioner (inputParam.hpp:29,31) it is a "next" item: precond_kind = -1
    static constexpr pps_compat::StackInfo kStack{PPS_SOLVER_BICGSTAB, -1, maxIteration, communicationON};
    BiCGSTAB(const BlockGrid<DIM, T_data>& blockGrid, const ExactSolutionAndBCs<DIM, T_data>& exactSolutionAndBCs,
             CommunicatorMPI<DIM, T_data>& communicatorMPI)
        : pps_compat::SolverAdapter<DIM, T_data, maxIteration>(blockGrid, exactSolutionAndBCs, communicatorMPI, kStack,
                                                               T_Preconditioner::kStack, tolerance, "BiCGSTAB") {
        static_assert(isMainLoop && communicationON, "BiCGSTAB is implemented as the main solver (isMainLoop, communicationON)");
    }
};

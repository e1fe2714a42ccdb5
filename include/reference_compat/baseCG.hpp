// baseCG.hpp (reference_compat) -- BaseCG<...> with the reference's parameter list (baseCG.hpp:14); the
// iteration of baseCG.hpp:115-228 runs in libpps_b200.so.
#pragma once
#include "iterativeSolverBase.hpp"

template <int DIM, typename T_data, int tolerance, int maxIteration, bool isMainLoop, bool communicationON, typename T_Preconditioner>
class BaseCG : public pps_compat::SolverAdapter<DIM, T_data, maxIteration> {
  public:
    static constexpr pps_compat::StackInfo kStack{PPS_SOLVER_CG, -1, maxIteration, communicationON};
    BaseCG(const BlockGrid<DIM, T_data>& blockGrid, const ExactSolutionAndBCs<DIM, T_data>& exactSolutionAndBCs,
           CommunicatorMPI<DIM, T_data>& communicatorMPI)
        : pps_compat::SolverAdapter<DIM, T_data, maxIteration>(blockGrid, exactSolutionAndBCs, communicatorMPI, kStack,
                                                               T_Preconditioner::kStack, tolerance, "baseCG") {
        static_assert(isMainLoop && communicationON, "BaseCG is implemented as the main solver (isMainLoop, communicationON)");
    }
};

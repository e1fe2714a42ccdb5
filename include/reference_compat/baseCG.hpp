// baseCG.hpp (reference_compat) -- BaseCG<...> with the reference's parameter list (baseCG.hpp:14); the
// iteration of baseCG.hpp:115-228 runs in libpps_b200.so.  Two roles:
//   main solver            isMainLoop, communicationON
//   nested preconditioner  !isMainLoop, communicationOFF, ChebyshevIteration(communicationOFF) inside  (T_Preconditioner3, inputParam.hpp:29)
#pragma once
#include "iterativeSolverBase.hpp"

template <int DIM, typename T_data, int tolerance, int maxIteration, bool isMainLoop, bool communicationON, typename T_Preconditioner>
class BaseCG : public pps_compat::SolverAdapter<DIM, T_data, maxIteration> {
  public:
    static constexpr int kAsSolver = (isMainLoop && communicationON) ? static_cast<int>(PPS_SOLVER_CG) : -1;
    static constexpr int kAsPreconditioner = (!isMainLoop && !communicationON && T_Preconditioner::kStack.precond_kind == PPS_PRECOND_CHEBYSHEV &&
                                              !T_Preconditioner::kStack.communication)
                                                 ? static_cast<int>(PPS_PRECOND_CG_CHEB_LOCAL)
                                                 : -1;
    static constexpr pps_compat::StackInfo kStack{kAsSolver, kAsPreconditioner, maxIteration, communicationON, tolerance,
                                                  T_Preconditioner::kStack.iterations};
    BaseCG(const BlockGrid<DIM, T_data>& blockGrid, const ExactSolutionAndBCs<DIM, T_data>& exactSolutionAndBCs,
           CommunicatorMPI<DIM, T_data>& communicatorMPI)
        : pps_compat::SolverAdapter<DIM, T_data, maxIteration>(blockGrid, exactSolutionAndBCs, communicatorMPI, kStack,
                                                               T_Preconditioner::kStack, tolerance, "baseCG") {
        static_assert(isMainLoop && communicationON, "as T_Solver, BaseCG needs isMainLoop and communicationON");
    }
};

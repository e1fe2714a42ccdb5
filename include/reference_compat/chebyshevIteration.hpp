// chebyshevIteration.hpp (reference_compat) -- ChebyshevIteration<DIM, T_data, tolerance, maxIteration, isMainLoop,
// communicationON, T_Preconditioner> (chebyshevIteration.hpp:14-156) as it is used by the reference: in the
// preconditioner slot with communicationOFF (inputParam.hpp:28), i.e. block-Jacobi.  maxIteration = chebyshevMax.
#pragma once
#include "iterativeSolverBase.hpp"

template <int DIM, typename T_data, int tolerance, int maxIteration, bool isMainLoop, bool communicationON, typename T_Preconditioner>
class ChebyshevIteration {
  public:
    static constexpr pps_compat::StackInfo kStack{-1, PPS_PRECOND_CHEBYSHEV, maxIteration, communicationON};
    ChebyshevIteration(const BlockGrid<DIM, T_data>&, const ExactSolutionAndBCs<DIM, T_data>&, CommunicatorMPI<DIM, T_data>&) {}
};

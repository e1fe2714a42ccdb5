// noneSolver.hpp (reference_compat) -- NoneSolver<DIM, T_data, tolerance, maxIteration> (noneSolver.hpp:14-28):
// the identity preconditioner.  On the GPU it is free: M(p) aliases p instead of being memcpy'd.
#pragma once
#include "iterativeSolverBase.hpp"

template <int DIM, typename T_data, int tolerance, int maxIteration>
class NoneSolver {
  public:
    static constexpr pps_compat::StackInfo kStack{-1, PPS_PRECOND_NONE, maxIteration, false};
    NoneSolver(const BlockGrid<DIM, T_data>&, const ExactSolutionAndBCs<DIM, T_data>&, CommunicatorMPI<DIM, T_data>&) {}
};

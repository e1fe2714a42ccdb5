// solvers.hpp (reference_compat) -- every solver-stack building block an inputParam.hpp may name.
//
//   main solver slot (isMainLoop, communicationON):   BiCGSTAB<...>, BaseCG<...>
//   preconditioner slot:                              NoneSolver<...>, ChebyshevIteration<..., communicationOFF, ...>
//
// The templates keep the reference's parameter lists; what they do is described in iterativeSolverBase.hpp.
#pragma once

#include "BiCGSTAB.hpp"
#include "baseCG.hpp"
#include "chebyshevIteration.hpp"
#include "noneSolver.hpp"

static_assert(sizeof(T_data) == 8, "the B200 path computes in fp64");

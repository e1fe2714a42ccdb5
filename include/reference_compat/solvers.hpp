// solvers.hpp (reference_compat) -- umbrella include, like the reference's solvers.hpp:4-7
#pragma once
#include "noneSolver.hpp"
#include "baseCG.hpp"
#include "BiCGSTAB.hpp"
#include "chebyshevIteration.hpp"

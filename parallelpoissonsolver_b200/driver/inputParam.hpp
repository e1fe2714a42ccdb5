// inputParam.hpp -- problem definition and solver-stack selection, compile time.
// Same names as the reference's solverPoissonMPI_CPU/include/inputParam.hpp:16-46; the solver class templates
// come from include/reference_compat/ and run on the GPU.  Override the grid at build time with
// -DPPS_NPX=.. -DPPS_NPY=.. -DPPS_NPZ=.., the boundary types with -DPPS_BCS=0,0,0,0,0,0 and the stack with
// -DPPS_UNPRECONDITIONED / -DPPS_USE_CG.
#ifndef PARAM_HPP
#define PARAM_HPP
#pragma once

#include <array>

#include "solvers.hpp"

constexpr int DIM = 3;

constexpr bool ischebyshevMainLoop = false;
constexpr bool isbiCGMainLoop1 = true;
constexpr bool isbiCGMainLoop2 = false;
constexpr bool communicationON = true;
constexpr bool communicationOFF = false;

using T_NoneSolver = NoneSolver<DIM, T_data, tollPreconditionerSolver, iterMaxPreconditioner>;
// block-Jacobi Chebyshev: no halo exchange inside the preconditioner
using T_Preconditioner2 = ChebyshevIteration<DIM, T_data, tollPreconditionerSolver, chebyshevMax, ischebyshevMainLoop, communicationOFF, T_NoneSolver>;

#if defined(PPS_UNPRECONDITIONED)
using T_ActivePreconditioner = T_NoneSolver;
#else
using T_ActivePreconditioner = T_Preconditioner2;
#endif

#if defined(PPS_USE_CG)
using T_Solver = BaseCG<DIM, T_data, tollMainSolver, iterMaxMainSolver, isbiCGMainLoop1, communicationON, T_ActivePreconditioner>;
#else
using T_Solver = BiCGSTAB<DIM, T_data, tollMainSolver, iterMaxMainSolver, isbiCGMainLoop1, communicationON, T_ActivePreconditioner>;
#endif

// result files of the reference's alpaka driver (its inputParam.hpp:46-47); -DPPS_WRITE_RESIDUAL / -DPPS_WRITE_SOLUTION
#ifdef PPS_WRITE_RESIDUAL
constexpr bool writeResidual = true;    // residualHistory.txt
#else
constexpr bool writeResidual = false;
#endif
#ifdef PPS_WRITE_SOLUTION
constexpr bool writeSolution = true;    // solution.dat, raw guard-padded blocks in rank order
#else
constexpr bool writeSolution = false;
#endif

#ifndef PPS_NPX
#define PPS_NPX 128
#define PPS_NPY 128
#define PPS_NPZ 256
#endif
#ifndef PPS_BCS
#define PPS_BCS 0, 1, 0, 1, 0, 1
#endif

constexpr std::array<int, 3> npglobal = {PPS_NPX, PPS_NPY, PPS_NPZ};
constexpr std::array<T_data, 3> ds = {0.1, 0.1, 0.1};
constexpr std::array<T_data, 3> origin = {0, 0, 0};
constexpr std::array<int, 3> guards = {1, 1, 1};
constexpr std::array<int, 6> bcsType = {PPS_BCS};          // 0 Dirichlet, 1 Neumann; x- x+ y- y+ z- z+
const std::array<T_data, 6> bcsValue = {0, 0, 0, 0, 0, 0};

#endif

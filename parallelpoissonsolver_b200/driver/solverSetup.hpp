// solverSetup.hpp -- compile-time numerics of the solve and the manufactured problem.
// Same names and meaning as the reference's solverPoissonMPI_CPU/include/solverSetup.hpp (the configuration
// surface README.md:36 describes), so a user's own copy of that header drops in unchanged.
#ifndef SOLVERSETUP_HPP
#define SOLVERSETUP_HPP
#pragma once

#include <cmath>

using T_data = double;

// kept for source compatibility with headers that ask for the MPI datatype of T_data
template <typename T>
inline MPI_Datatype getMPIType();
template <>
inline MPI_Datatype getMPIType<float>() { return MPI_FLOAT; }
template <>
inline MPI_Datatype getMPIType<double>() { return MPI_DOUBLE; }

constexpr T_data PI = 3.141592653589793;

// integer template arguments carry tolerances; the effective value is <int> * tollScalingFactor
constexpr T_data tollScalingFactor = 1e-10;
constexpr int orderNeumanBcs = 2;              // 2nd-order (mirror) ghost cells on Neumann faces

constexpr int tollMainSolver = 1e2;            // -> 1e-8 relative residual
constexpr int iterMaxMainSolver = 1700;
constexpr bool trackErrorFromIterationHistory = 1;

constexpr int tollPreconditionerSolver = 1e4;
constexpr int iterMaxPreconditioner = 150;

// Chebyshev preconditioner: spectrum window [rescaleEigMin * lambda_min, rescaleEigMax * lambda_max] * (1 + epsilon)
constexpr T_data epsilon = 1e-4;
constexpr T_data rescaleEigMin = 500;
constexpr T_data rescaleEigMax = 1 - 1e-4;
constexpr int chebyshevMax = 11;

// Manufactured solution u, right-hand side f = laplace(u) and the face-normal derivatives of u.
template <int DIM, typename T>
class ExactSolutionAndBCs {
  public:
    inline T setFieldB(const T x, const T y, const T z) const { return -sin(x) - cos(y) - 3 * sin(z) + 2 * y * z + 2; }
    inline T trueSolutionFxyz(const T x, const T y, const T z) const { return sin(x) + cos(y) + 3 * sin(z) + x * x * y * z + x * x + 10; }
    inline T trueSolutionDdir(const T x, const T y, const T z, const int dir) const {
        switch (dir) {
            case 0: return cos(x) + 2 * x * y * z + 2 * x;
            case 1: return -sin(y) + x * x * z;
            case 2: return 3 * cos(z) + x * x * y;
            default: return -100;
        }
    }
};

#endif

// matrixFreeOperatorA.hpp (reference_compat) -- MatrixFreeOperatorA keeps its constructor (main.cpp:80) and is
// passed to T_Solver::operator() like in the reference; the 7-point operator itself
// (matrixFreeOperatorA.hpp:22-39) is evaluated by the stencil kernels of libpps_b200.so.
#pragma once

#include "blockGrid.hpp"

template <int DIM, typename T_data>
class MatrixFreeOperatorA {
  public:
    explicit MatrixFreeOperatorA(const BlockGrid<DIM, T_data>& blockGrid) : blockGrid_(blockGrid) {}

  private:
    const BlockGrid<DIM, T_data>& blockGrid_;
};

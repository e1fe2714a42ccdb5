// communicationMPI.hpp (reference_compat) -- CommunicatorMPI keeps its constructor (main.cpp:78) but owns
// nothing: the face halo exchange of communicationMPI.hpp:51-316 runs inside libpps_b200.so (NCCL send/recv
// between GPUs, copy kernels between blocks on one GPU).
#pragma once

#include "blockGrid.hpp"

template <int DIM, typename T_data>
class CommunicatorMPI {
  public:
    explicit CommunicatorMPI(const BlockGrid<DIM, T_data>& blockGrid) : blockGrid_(blockGrid) {}
    void waitAllandCheckSend() const {}
    void waitAllandCheckRcv() const {}

  private:
    const BlockGrid<DIM, T_data>& blockGrid_;
};

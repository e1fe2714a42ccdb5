// blockGrid.hpp (reference_compat) -- BlockGrid with the reference's constructor and getters
// (solverPoissonMPI_CPU/include/blockGrid.hpp:15-145), host-only bookkeeping written from scratch.
#pragma once

#include <array>
#include <cmath>

#include "solverSetup.hpp"

template <int DIM, typename T_data>
class BlockGrid {
  public:
    BlockGrid(const std::array<int, 3>& nranks, const int my_rank, const std::array<int, 3>& npglobal,
              const std::array<T_data, 3>& ds, const std::array<T_data, 3>& origin, const std::array<int, 3>& guards,
              const std::array<int, 6>& bcsType, const std::array<T_data, 6>& bcsValue)
        : nranks_(nranks), rank_(my_rank), npglobal_(npglobal), ds_(ds), origin_(origin), guards_(guards), bcsType_(bcsType),
          bcsValue_(bcsValue) {
        static_assert(DIM >= 1 && DIM <= 3, "DIM must be 1, 2 or 3");
        location_ = {my_rank % nranks[0], (my_rank / nranks[0]) % nranks[1], my_rank / (nranks[0] * nranks[1])};
        ntotGuards_ = ntotNoGuards_ = 1;
        numComm_ = 0;
        for (int d = 0; d < 3; d++) {
            if (d >= DIM) {
                // an unused axis holds one point, no guards, no faces (blockGrid.hpp:166-167,178-179,193-204)
                nlocal_[d] = nlocalGuards_[d] = 1;
                limitsData_[2 * d] = limitsSolver_[2 * d] = 0;
                limitsData_[2 * d + 1] = limitsSolver_[2 * d + 1] = 1;
                hasBoundary_[2 * d] = hasBoundary_[2 * d + 1] = hasComm_[2 * d] = hasComm_[2 * d + 1] = false;
                continue;
            }
            nlocal_[d] = npglobal[d] / nranks[d];
            nlocalGuards_[d] = nlocal_[d] + 2 * guards[d];
            ntotGuards_ *= nlocalGuards_[d];
            ntotNoGuards_ *= nlocal_[d];
            limitsData_[2 * d] = guards[d];
            limitsData_[2 * d + 1] = nlocal_[d] + guards[d];
            const bool first = location_[d] == 0, last = location_[d] == nranks[d] - 1;
            hasBoundary_[2 * d] = first;
            hasBoundary_[2 * d + 1] = last;
            hasComm_[2 * d] = nranks[d] > 1 && !first;
            hasComm_[2 * d + 1] = nranks[d] > 1 && !last;
            numComm_ += int(hasComm_[2 * d]) + int(hasComm_[2 * d + 1]);
            limitsSolver_[2 * d] = limitsData_[2 * d] + ((bcsType[2 * d] == 0 && first) ? 1 : 0);
            limitsSolver_[2 * d + 1] = limitsData_[2 * d + 1] - ((bcsType[2 * d + 1] == 0 && last) ? 1 : 0);
        }
        limitsComm_ = limitsData_;
        numElementsComm_ = {0, 0, 0};
        for (int d = 0; d < DIM; d++) {
            if (hasComm_[2 * d + 1]) limitsComm_[2 * d] = limitsData_[2 * d + 1] - guards[d];
            if (hasComm_[2 * d]) limitsComm_[2 * d + 1] = limitsData_[2 * d] + guards[d];
            numElementsComm_[d] = 1;
            for (int e = 0; e < DIM; e++)
                if (e != d) numElementsComm_[d] *= nlocal_[e];
        }
        eigen(true);
        eigen(false);
    }

    const std::array<int, 3> getNranks() const { return nranks_; }
    const std::array<int, 3> getNpglobal() const { return npglobal_; }
    const std::array<T_data, 3> getDs() const { return ds_; }
    const std::array<T_data, 3> getOrigin() const { return origin_; }
    const std::array<int, 3> getGuards() const { return guards_; }
    int getMyrank() const { return rank_; }
    const std::array<int, 3> getGlobalLocation() const { return location_; }
    const std::array<int, 3> getNlocalNoGuards() const { return nlocal_; }
    const std::array<int, 3> getNlocalGuards() const { return nlocalGuards_; }
    const std::array<int, 6> getIndexLimitsData() const { return limitsData_; }
    const std::array<int, 6> getIndexLimitsSolver() const { return limitsSolver_; }
    const std::array<int, 6> getIndexLimitsComm() const { return limitsComm_; }
    int getNumCommunication() const { return numComm_; }
    const std::array<int, 3> getNumElementsComm() const { return numElementsComm_; }
    int getNtotLocalGuards() const { return ntotGuards_; }
    int getNtotLocalNoGuards() const { return ntotNoGuards_; }
    const std::array<bool, 6> getHasBoundary() const { return hasBoundary_; }
    const std::array<bool, 6> getHasCommunication() const { return hasComm_; }
    const std::array<int, 6> getBcsType() const { return bcsType_; }
    const std::array<T_data, 6> getBcsValue() const { return bcsValue_; }
    bool checkBCsSet() const { return false; }
    const std::array<T_data, 2> getEigenValuesLocal() const { return eigLocal_; }
    const std::array<T_data, 2> getEigenValuesGlobal() const { return eigGlobal_; }
    int getNtotNpglobal() const {
        int n = 1;
        for (int d = 0; d < DIM; d++) n *= npglobal_[d];
        return n;
    }

  private:
    void eigen(bool global) {
        T_data lo = 0, hi = 0;
        for (int d = 0; d < DIM; d++) {
            const int n = global ? npglobal_[d] - (bcsType_[2 * d] == 0) - (bcsType_[2 * d + 1] == 0)
                                 : limitsSolver_[2 * d + 1] - limitsSolver_[2 * d];
            const T_data a = std::sin(1 * PI / 2 / (n + 1)), b = std::sin(n * PI / 2 / (n + 1));
            lo += 4 * a * a / (ds_[d] * ds_[d]);
            hi += 4 * b * b / (ds_[d] * ds_[d]);
        }
        (global ? eigGlobal_ : eigLocal_) = {lo, hi};
    }

    std::array<int, 3> nranks_;
    int rank_;
    std::array<int, 3> npglobal_;
    std::array<T_data, 3> ds_, origin_;
    std::array<int, 3> guards_;
    std::array<int, 6> bcsType_;
    std::array<T_data, 6> bcsValue_;
    std::array<int, 3> location_{}, nlocal_{}, nlocalGuards_{}, numElementsComm_{};
    std::array<int, 6> limitsData_{}, limitsSolver_{}, limitsComm_{};
    std::array<bool, 6> hasBoundary_{}, hasComm_{};
    int ntotGuards_ = 1, ntotNoGuards_ = 1, numComm_ = 0;
    std::array<T_data, 2> eigLocal_{}, eigGlobal_{};
};

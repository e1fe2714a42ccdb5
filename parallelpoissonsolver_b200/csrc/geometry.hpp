// geometry.hpp -- host-side block geometry (the facts BlockGrid holds, blockGrid.hpp:15-366), re-derived
// for the pitched device layout.  Tiny, host only.
#pragma once

#include <array>
#include <cmath>

#include "../../include/pps_b200.h"
#include "common.cuh"

namespace pps {

constexpr double kPi = 3.141592653589793;   // solverSetup.hpp:20

// DIM < 3 (inputParam.hpp:16, blockGrid.hpp:160-206): the reference gives the unused axes one point, no guards and no
// faces.  On the device such an axis keeps its two guard planes (always zero) so that every kernel stays the 3-D kernel:
// the operator's term along it is (0 - 2u + 0) / inf = -0 in PARITY arithmetic and (..) * 0 in FAST arithmetic (Coef is
// built that way in solver.cu).  BlockGeom uses DEVICE numbering throughout (the data point of an unused axis is index 1);
// pps_block_info_get converts to the reference's numbering (index 0, extent 1).
struct BlockGeom {
    int rank = 0;
    int dim = 3;
    int loc[3] = {0, 0, 0};      // globalLocation_
    int n[3] = {0, 0, 0};        // nlocal_noguards_
    int ld[6] = {0};             // indexLimitsData_
    int ls[6] = {0};             // indexLimitsSolver_
    bool hb[6] = {false};        // hasBoundary_
    bool hc[6] = {false};        // hasCommunication_
    int nbr[6] = {-1, -1, -1, -1, -1, -1};   // neighbour rank per face (communicationMPI.hpp:79)
    double eig_global[2] = {0, 0}, eig_local[2] = {0, 0};
    Dims dims{};

    Box solver_box() const { return Box{ls[0], ls[1], ls[2], ls[3], ls[4], ls[5]}; }
    Box data_box() const { return Box{ld[0], ld[1], ld[2], ld[3], ld[4], ld[5]}; }
    // extent of axis d in the reference's host layout (blockGrid.hpp:172-182)
    int ref_extent(int d) const { return d < dim ? n[d] + 2 : 1; }
    long long ref_total() const { return static_cast<long long>(ref_extent(0)) * ref_extent(1) * ref_extent(2); }
    // element offset of reference cell (i, j, k) in the device layout
    long long at(int i, int j, int k) const { return kOff + i + dims.pitch * (j + static_cast<long long>(n[1] + 2) * k); }
    long long stride(int axis) const { return axis == 0 ? 1 : (axis == 1 ? dims.pitch : dims.plane); }
    // tangential axes of a face, fast one first
    void tangential(int face, int& u, int& v) const {
        const int d = face / 2;
        u = d == 0 ? 1 : 0;
        v = d == 2 ? 1 : 2;
    }
};

inline void eigen_pair(int dim, const double ds[3], const int n[3], double out[2]) {
    // blockGrid.hpp:301-340: sum over d < DIM of 4 sin^2(pi/(2(n_d+1)))/ds_d^2 and 4 sin^2(n_d pi/(2(n_d+1)))/ds_d^2
    double lo = 0, hi = 0;
    for (int i = 0; i < dim; i++) {
        const double a = std::sin(1 * kPi / 2 / (n[i] + 1));
        const double b = std::sin(n[i] * kPi / 2 / (n[i] + 1));
        lo += 4 * a * a / (ds[i] * ds[i]);
        hi += 4 * b * b / (ds[i] * ds[i]);
    }
    out[0] = lo;
    out[1] = hi;
}

inline BlockGeom make_block(const pps_config& c, int rank) {
    BlockGeom g;
    g.rank = rank;
    g.dim = (c.dim == 1 || c.dim == 2) ? c.dim : 3;
    g.loc[0] = rank % c.nranks[0];                                  // blockGrid.hpp:151-158
    g.loc[1] = (rank / c.nranks[0]) % c.nranks[1];
    g.loc[2] = rank / (c.nranks[0] * c.nranks[1]);
    for (int d = 0; d < 3; d++) {
        if (d >= g.dim) {                                           // :166-167,193-204,237,261: one point, no faces
            g.n[d] = 1;
            g.ld[2 * d] = g.ls[2 * d] = 1;
            g.ld[2 * d + 1] = g.ls[2 * d + 1] = 2;
            continue;
        }
        g.n[d] = c.npglobal[d] / c.nranks[d];                       // :160-170 (integer division, like the reference)
        g.ld[2 * d] = 1;                                            // :184-206 with guards = 1
        g.ld[2 * d + 1] = g.n[d] + 1;
        const bool first = g.loc[d] == 0, last = g.loc[d] == c.nranks[d] - 1, many = c.nranks[d] > 1;
        g.hb[2 * d] = first;                                        // :234-254
        g.hb[2 * d + 1] = last;
        g.hc[2 * d] = many && !first;                               // :256-299
        g.hc[2 * d + 1] = many && !last;
        g.ls[2 * d] = g.ld[2 * d] + ((c.bcs_type[2 * d] == 0 && g.hb[2 * d]) ? 1 : 0);            // :208-222
        g.ls[2 * d + 1] = g.ld[2 * d + 1] - ((c.bcs_type[2 * d + 1] == 0 && g.hb[2 * d + 1]) ? 1 : 0);
        for (int up = 0; up < 2; up++) {
            if (!g.hc[2 * d + up]) continue;
            int l[3] = {g.loc[0], g.loc[1], g.loc[2]};
            l[d] += up ? 1 : -1;
            g.nbr[2 * d + up] = l[0] + l[1] * c.nranks[0] + l[2] * c.nranks[0] * c.nranks[1];
        }
    }
    int nl[3], ng[3];
    for (int d = 0; d < 3; d++) {
        nl[d] = g.ls[2 * d + 1] - g.ls[2 * d];
        ng[d] = c.npglobal[d] - (c.bcs_type[2 * d] == 0) - (c.bcs_type[2 * d + 1] == 0);
    }
    eigen_pair(g.dim, c.ds, nl, g.eig_local);
    eigen_pair(g.dim, c.ds, ng, g.eig_global);
    g.dims = make_dims(g.n[0], g.n[1], g.n[2]);
    return g;
}

}  // namespace pps

// setproblem.cpp -- host-side setProblem() of the driver as a C-callable helper (libpps_driver_host.so).
//
// Replaces, for callers that are not C++ (bench.py, tools/): IterativeSolverBase::setProblem
// (solverPoissonMPI_CPU/include/iterativeSolverBase.hpp:51-55) = applyDirichletBCsFromFunction (:557-603) +
// setFieldValuefromFunction (:537-555), evaluated with the driver's own ExactSolutionAndBCs (driver/solverSetup.hpp,
// the user-editable manufactured solution; same formulas and libm calls as the reference's solverSetup.hpp:48-59) and the
// reference's coordinate expression (:547-549), so the arrays are bit-identical to what the reference's setProblem
// writes.  Pure host code: untimed set-up outside the hot path, exactly where the reference does it.  The planes of a
// block are filled by `nthreads` host threads.
#include <algorithm>
#include <array>
#include <cmath>
#include <thread>
#include <vector>

#include "mpi.h"   // include/reference_compat: the type names solverSetup.hpp mentions
#include "solverSetup.hpp"

namespace {

struct Geom {
    int n[3];      // local points without guards (blockGrid.hpp:165)
    int loc[3];    // globalLocation, rank -> (lx, ly, lz) x fastest (blockGrid.hpp:151-158)
    bool lo[3], hi[3];   // physical boundary on the lower / upper face
};

Geom make_geom(const int np[3], const int nranks[3], int rank) {
    Geom g;
    g.loc[0] = rank % nranks[0];
    g.loc[1] = (rank / nranks[0]) % nranks[1];
    g.loc[2] = rank / (nranks[0] * nranks[1]);
    for (int d = 0; d < 3; d++) {
        g.n[d] = np[d] / nranks[d];
        g.lo[d] = g.loc[d] == 0;
        g.hi[d] = g.loc[d] == nranks[d] - 1;
    }
    return g;
}

}  // namespace

extern "C" int pps_driver_set_problem(const int npglobal[3], const int nranks[3], int rank, const double ds[3], const double origin[3],
                                      const int bcs[6], double* x, double* b, int nthreads) {
    const Geom g = make_geom(npglobal, nranks, rank);
    const ExactSolutionAndBCs<3, double> exact;
    const long sj = g.n[0] + 2, sk = sj * (g.n[1] + 2);
    // iterativeSolverBase.hpp:547-549 with indexLimitsData_[2d] = guards = 1
    auto coord = [&](int d, int i) { return origin[d] + (i - 1) * ds[d] + g.loc[d] * (g.n[d]) * ds[d]; };
    nthreads = std::max(1, std::min(nthreads, g.n[2]));
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++) {
        pool.emplace_back([&, t]() {
            const int k0 = 1 + static_cast<int>(static_cast<long>(g.n[2]) * t / nthreads);
            const int k1 = 1 + static_cast<int>(static_cast<long>(g.n[2]) * (t + 1) / nthreads);
            for (int k = k0; k < k1; k++) {
                const double z = coord(2, k);
                const bool zface = (k == 1 && g.lo[2] && bcs[4] == 0) || (k == g.n[2] && g.hi[2] && bcs[5] == 0);
                for (int j = 1; j <= g.n[1]; j++) {
                    const double y = coord(1, j);
                    const bool yface = (j == 1 && g.lo[1] && bcs[2] == 0) || (j == g.n[1] && g.hi[1] && bcs[3] == 0);
                    double* xr = x + sj * j + sk * k;
                    double* br = b + sj * j + sk * k;
                    for (int i = 1; i <= g.n[0]; i++) {
                        const double xx = coord(0, i);
                        br[i] = exact.setFieldB(xx, y, z);                                              // :537-555
                        const bool xface = (i == 1 && g.lo[0] && bcs[0] == 0) || (i == g.n[0] && g.hi[0] && bcs[1] == 0);
                        if (xface || yface || zface) xr[i] = exact.trueSolutionFxyz(xx, y, z);          // :557-603
                    }
                }
            }
        });
    }
    for (auto& th : pool) th.join();
    return 0;
}

// du/dn of the manufactured solution on one face (solverSetup.hpp:61-85 trueSolutionDdir), data range, lower axis fastest:
// what pps_set_neumann_face takes
extern "C" int pps_driver_neumann_face(const int npglobal[3], const int nranks[3], int rank, const double ds[3], const double origin[3],
                                       int face, double* out) {
    const Geom g = make_geom(npglobal, nranks, rank);
    const ExactSolutionAndBCs<3, double> exact;
    auto coord = [&](int d, int i) { return origin[d] + (i - 1) * ds[d] + g.loc[d] * (g.n[d]) * ds[d]; };
    const int dir = face / 2;
    int lim[6] = {1, g.n[0] + 1, 1, g.n[1] + 1, 1, g.n[2] + 1};
    if (face % 2 == 0) lim[2 * dir + 1] = lim[2 * dir] + 1;
    else lim[2 * dir] = lim[2 * dir + 1] - 1;
    long q = 0;
    for (int k = lim[4]; k < lim[5]; k++)
        for (int j = lim[2]; j < lim[3]; j++)
            for (int i = lim[0]; i < lim[1]; i++) out[q++] = exact.trueSolutionDdir(coord(0, i), coord(1, j), coord(2, k), dir);
    return 0;
}

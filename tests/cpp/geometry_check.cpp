// CPU harness for parallelpoissonsolver_b200/csrc/geometry.hpp (pure host code): prints the block geometry the CUDA library
// derives for one rank, in the REFERENCE's numbering, so that tests/test_geometry_cpu.py can compare it with the oracle
// (which is pinned to the unmodified reference).  usage: geometry_check dim nx ny nz px py pz b0..b5 dsx dsy dsz rank
#include <cstdio>
#include <cstdlib>

#include "../../parallelpoissonsolver_b200/csrc/geometry.hpp"

int main(int argc, char** argv) {
    if (argc < 18) return 2;
    pps_config c{};
    int a = 1;
    c.dim = std::atoi(argv[a++]);
    for (int d = 0; d < 3; d++) c.npglobal[d] = std::atoi(argv[a++]);
    for (int d = 0; d < 3; d++) c.nranks[d] = std::atoi(argv[a++]);
    for (int f = 0; f < 6; f++) c.bcs_type[f] = std::atoi(argv[a++]);
    for (int d = 0; d < 3; d++) c.ds[d] = std::atof(argv[a++]);
    const int rank = std::atoi(argv[a++]);
    const pps::BlockGeom g = pps::make_block(c, rank);
    std::printf("loc %d %d %d\n", g.loc[0], g.loc[1], g.loc[2]);
    std::printf("nlocal %d %d %d\n", g.n[0], g.n[1], g.n[2]);
    std::printf("nguards %d %d %d\n", g.ref_extent(0), g.ref_extent(1), g.ref_extent(2));
    std::printf("limits_data");
    for (int f = 0; f < 6; f++) std::printf(" %d", g.ld[f] - (f / 2 >= g.dim ? 1 : 0));
    std::printf("\nlimits_solver");
    for (int f = 0; f < 6; f++) std::printf(" %d", g.ls[f] - (f / 2 >= g.dim ? 1 : 0));
    std::printf("\nhas_boundary");
    for (int f = 0; f < 6; f++) std::printf(" %d", int(g.hb[f]));
    std::printf("\nhas_comm");
    for (int f = 0; f < 6; f++) std::printf(" %d", int(g.hc[f]));
    std::printf("\nntot %lld\n", g.ref_total());
    std::printf("eig %.17g %.17g %.17g %.17g\n", g.eig_global[0], g.eig_global[1], g.eig_local[0], g.eig_local[1]);
    // device layout invariants
    std::printf("pitch %lld plane %lld total %lld first_data_col %lld\n", g.dims.pitch, g.dims.plane, g.dims.total, g.at(1, 0, 0));
    return 0;
}

// ref_launcher.cpp -- starts px*py*pz rank-threads, each running the reference's own
// main() (compiled unmodified with -Dmain=ref_main against oracle/mpi_shim/mpi.h).
// TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline).
#include <cstdio>
#include <cstdlib>

#include "mpi.h"

int ref_main(int argc, char** argv);

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s px py pz\n", argv[0]);
        return 2;
    }
    const int world = std::atoi(argv[1]) * std::atoi(argv[2]) * std::atoi(argv[3]);
    if (world < 1) {
        std::fprintf(stderr, "bad rank grid\n");
        return 2;
    }
    return pps_shim_run(world, ref_main, argc, argv);
}

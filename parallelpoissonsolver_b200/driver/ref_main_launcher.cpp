// ref_main_launcher.cpp -- runs an UNMODIFIED reference main() (solverPoissonMPI_CPU/src/main.cpp compiled with
// -Dmain=ref_main against include/reference_compat) on px*py*pz rank-threads: the drop-in demonstration of
// INTEGRATION.md.  Built by oracle/build_ref.py into oracle/_ref/bin/ref_main_on_b200_<config>.
#include <mpi.h>

#include <cstdlib>
#include <iostream>

int ref_main(int argc, char** argv);

int main(int argc, char** argv) {
    if (argc < 4) {
        std::cerr << "usage: " << argv[0] << " px py pz" << std::endl;
        return 2;
    }
    const int world = std::atoi(argv[1]) * std::atoi(argv[2]) * std::atoi(argv[3]);
    return pps_compat::run_ranks(world, [&](int) { return ref_main(argc, argv); });
}

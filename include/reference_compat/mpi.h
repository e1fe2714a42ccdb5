// mpi.h (reference_compat) -- the sliver of MPI the reference DRIVER touches (main.cpp:21-29,103,116,132;
// solverSetup.hpp:8-18), provided without MPI: the ranks of `solverPoisson px py pz` are threads of one
// process, one per GPU, and everything the reference's SOLVERS did over MPI (halo exchange, allreduce,
// error gather) happens inside libpps_b200.so over NCCL.  Not an MPI implementation.
#pragma once

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_DOUBLE 1
#define MPI_FLOAT 2

struct pps_handle;

namespace pps_compat {

struct World {
    int size = 1;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    long generation = 0;
    // shared between the rank threads by the solver adapters
    unsigned char unique_id[128] = {0};
    pps_handle* shared_handle = nullptr;   // "virtual ranks" mode: more ranks than GPUs, one handle hosts all blocks
    std::vector<double> gather;            // error gather in that mode
};

inline World& world() {
    static World w;
    return w;
}
inline int& this_rank() {
    static thread_local int r = 0;
    return r;
}

inline void barrier() {
    World& w = world();
    if (w.size == 1) return;
    std::unique_lock<std::mutex> lk(w.m);
    const long gen = w.generation;
    if (++w.arrived == w.size) {
        w.arrived = 0;
        w.generation++;
        w.cv.notify_all();
    } else {
        w.cv.wait(lk, [&] { return w.generation != gen; });
    }
}

// run fn(rank) on `n` rank-threads (the replacement of `mpirun -n`)
inline int run_ranks(int n, const std::function<int(int)>& fn) {
    World& w = world();
    w.size = n;
    std::vector<int> rc(n, 0);
    if (n == 1) {
        this_rank() = 0;
        return fn(0);
    }
    std::vector<std::thread> th;
    for (int r = 0; r < n; r++)
        th.emplace_back([&, r] {
            this_rank() = r;
            rc[r] = fn(r);
        });
    for (auto& t : th) t.join();
    for (int v : rc)
        if (v) return v;
    return 0;
}

}  // namespace pps_compat

inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int* n) { *n = pps_compat::world().size; return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = pps_compat::this_rank(); return MPI_SUCCESS; }
inline int MPI_Barrier(MPI_Comm) { pps_compat::barrier(); return MPI_SUCCESS; }

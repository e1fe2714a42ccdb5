// common.cuh -- shared device/host definitions of the B200 Poisson hot path.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace pps {

// ---------------------------------------------------------------------------------------------
// Device layout of one field of one block (DESIGN.md "Data layout in HBM").
// Reference layout: dense (nx+2)(ny+2)(nz+2), x fastest (blockGrid.hpp:172-182).  Ours keeps x fastest
// and the guards in-array, but pads every row to `pitch` doubles (a multiple of 16 = 128 B) and shifts
// it by kOff so that the first DATA cell of every row (reference i = 1) is 128-byte aligned:
//     addr(i, j, k) = base + kOff + i + pitch * (j + (ny + 2) * k)            i, j, k in reference numbering
// ---------------------------------------------------------------------------------------------
constexpr int kOff = 15;          // column of reference i = 0 (x- guard); i = 1 sits at column 16
constexpr int kFirstDataCol = 16;
constexpr int kTailPad = 64;      // slack doubles after the last row

struct Dims {
    int nx, ny, nz;        // local points without guards
    long long pitch;       // doubles per row
    long long plane;       // pitch * (ny + 2)
    long long total;       // plane * (nz + 2) + kTailPad
};

inline Dims make_dims(int nx, int ny, int nz) {
    Dims d;
    d.nx = nx; d.ny = ny; d.nz = nz;
    d.pitch = ((static_cast<long long>(nx) + 2 + kOff) + 15) / 16 * 16;
    d.plane = d.pitch * (ny + 2);
    d.total = d.plane * (nz + 2) + kTailPad;
    return d;
}

// half-open index box in reference numbering (guards at 0 and n+1)
struct Box {
    int i0, i1, j0, j1, k0, k1;
};

// where the CTA grid of a launch starts: tile indices in x / y and the first z-plane (launches may cover any sub-box)
struct TileOrigin {
    int bx0, by0, kfirst;
};

// In-kernel wait for a halo exchange that runs concurrently on the halo stream (z-slab decompositions): the CTAs of
// the first / last z-chunk are dispatched last (chunk index rotated by `shift`) and their TMA producer spins on
// `flag` (set to `epoch` by the halo stream when the faces have landed) before it loads guard plane lo / hi.
struct HaloWait {
    const unsigned int* flag;   // nullptr: no waiting
    unsigned int epoch;
    int lo_plane, hi_plane;     // guard planes that are being received (-1: none)
    int shift;                  // rotation of blockIdx.z
    // peer-memory transport (PPS_OVERLAP=3): one flag per face, written by the NEIGHBOUR's copy engine -> system scope.
    // Trailing members, so the existing five-value initialisers leave them null / 0.
    const unsigned int* flag_hi;   // flag of the upper face (nullptr: `flag` covers both faces)
    int sys_scope;                 // 1: acquire at system scope
};

struct Coef {
    double ds[3];
    double ds2[3];     // ds*ds, as the reference evaluates it (matrixFreeOperatorA.hpp:35-37)
    double inv[3];     // 1/(ds*ds), fast mode
};

// Krylov scalars that live on the device for a whole solve (BiCGSTAB.hpp:75-83 locals).
struct Ctl {
    double rho0, alpha, omega, beta, err;
    double rz;            // CG: r.z of the current iterate (baseCG.hpp totSum1)
    double norm_b;        // normFieldB_
    double tol;
    double sums[8];       // raw results of the last fused reductions (after the allreduce when world > 1)
    int iter;
    int done;             // set when err < tol (BiCGSTAB.hpp:288) or iter == max_iter; later kernels return at once
    int max_iter;
    int pad;
    double* hist_err;     // host-mapped: errorFromIterationHistory_ (iterativeSolverBase.hpp:43)
    double* hist_alpha;
    double* hist_omega;
    double* hist_rho;
};

// scalar updates that follow a fused reduction (run by the last CTA of the reducing kernel, or by a
// 1-thread kernel after the NCCL allreduce when world > 1)
enum ScalarOp : int {
    OP_NONE = 0,
    OP_BICG_ALPHA = 1,     // alpha = rho0 / sum(r0.v)                      BiCGSTAB.hpp:156-164
    OP_BICG_OMEGA = 2,     // omega = sum(s.t) / sum(t.t)                   BiCGSTAB.hpp:216-225
    OP_BICG_RHO = 3,       // rho1, err, beta, rho0 <- rho1, iter++         BiCGSTAB.hpp:247-259,274-291
    OP_NORM_B = 4,         // norm_b = sqrt(sum)                            iterativeSolverBase.hpp:216-225
    OP_RESIDUAL0 = 5,      // err = sqrt(sum); hist[0] = err                BiCGSTAB.hpp:113-123
    OP_RESIDUAL_FINAL = 6, // sums[0] <- sqrt(sum) (errorComputeOperator_)  BiCGSTAB.hpp:305-308
    OP_CG_ALPHA = 7,       // alpha = sum(r.z) / sum(p.Ap)                  baseCG.hpp:141-151
    OP_CG_BETA = 8,        // beta = sum(r.z)new / (r.z)old, err, iter++    baseCG.hpp:183-194,211-227
};

// Allreduce fused into the reducing kernel over peer memory (PPS_ALLREDUCE_P2P=1): the last CTA of the reduction writes
// its sums into every rank's mailbox, waits for the mailboxes it receives and adds them up in rank order (identical bits
// on every rank).  Mailbox of a rank: [2 parities][world][4] doubles, slot 3 = epoch (as a 64-bit integer), written last.
struct PeerReduce {
    double* const* peer_mail;   // device array of `world` pointers: every rank's mailbox (mine included)
    double* my_mail;
    unsigned int epoch;         // 0: disabled
    int world, rank;
};

struct RedCtx {
    double* partials;          // [kMaxAcc][capacity]
    unsigned int* counter;     // ticket
    long long capacity;        // stride between accumulator rows in `partials`
    unsigned int total_ctas;   // CTAs of all launches that feed this reduction
    unsigned int cta_offset;   // first partial slot of this launch
    int nacc;                  // accumulators in use
    int op;                    // ScalarOp applied by the last CTA (OP_NONE when an allreduce follows)
    Ctl* ctl;
    PeerReduce pr;             // in-kernel allreduce across GPUs (pr.epoch != 0)
};

constexpr int kMaxAcc = 3;

#define PPS_CUDA_CHECK(expr)                                                                         \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            throw std::runtime_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) +     \
                                     " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")");        \
        }                                                                                            \
    } while (0)

}  // namespace pps

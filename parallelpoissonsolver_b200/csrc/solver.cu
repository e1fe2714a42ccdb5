// solver.cu -- host side of libpps_b200.so: handle, Krylov drivers, halo exchange, C ABI (include/pps_b200.h).
//
// One handle = one GPU = one CUDA stream.  With world_size == 1 the handle hosts every block of the
// px*py*pz decomposition ("virtual ranks"); with world_size == px*py*pz it hosts block `rank` and talks to
// its neighbours through NCCL.  The Krylov scalars (alpha, omega, beta, rho, ||r||) never leave the device
// during a solve: the last CTA of each reducing kernel computes them (world 1) or a one-thread kernel does
// after the NCCL allreduce; the host only reads the residual history (host-mapped) a few iterations late to
// decide when to stop launching, and every kernel returns at once after the device has set `done`.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../../include/pps_b200.h"
#include "common.cuh"
#include "geometry.hpp"
#include "kernels.cuh"
#include "nccl_dyn.hpp"
#include "stencil_tma.cuh"
#include "cheb_blocked.cuh"

namespace pps {

#define PPS_NCCL_CHECK(expr)                                                                          \
    do {                                                                                              \
        ncclResult_t _r = (expr);                                                                     \
        if (_r != ncclSuccess) {                                                                      \
            throw std::runtime_error(std::string(#expr) + " failed: " + nccl().GetErrorString(_r) +      \
                                     " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")");         \
        }                                                                                             \
    } while (0)

enum KernelClass : int {
    KC_APPLY_DOT = 0,   // v = A p, r0.v
    KC_S_UPDATE,        // r -= alpha v
    KC_APPLY_DOT2,      // t = A s, s.t, t.t
    KC_XR_UPDATE,       // x += .., r -= omega t, r0.r, r.r
    KC_P_UPDATE,        // p = r + beta (p - omega v)
    KC_HALO,
    KC_GHOST,
    KC_CHEB_FIRST,
    KC_CHEB_STEP,
    KC_RESIDUAL,
    KC_SETUP,
    KC_CG_APPLY,
    KC_CG_XR,
    KC_CG_P,
    KC_DOT,
    KC_SCALAR,
    KC_APPLY,
    KC_FUSED_P,   // p = r + beta (p - omega v), v = A p, r0.v
    KC_FUSED_S,   // s = r - alpha v, t = A s, s.t, t.t
    KC_CHEB_BLOCKED,   // several Chebyshev sweeps per pass (cheb_blocked.cuh)
    KC_COUNT
};
static const char* kKernelNames[KC_COUNT] = {
    "stencil_dot(v=A*p, r0.v)", "s_update(r-=alpha*v)", "stencil_dot2(t=A*s, s.t, t.t)",
    "xr_update(x+=.., r-=omega*t, r0.r, r.r)", "p_update(p=r+beta*(p-omega*v))", "halo", "neumann_ghost",
    "cheb_first", "cheb_step", "residual(r=b-A*x, r.r)", "setup", "cg_apply(Ap, r.z, p.Ap)", "cg_xr", "cg_p", "dot",
    "scalar_op", "stencil(y=A*x)", "fused_p(p=r+beta*(p-omega*v), v=A*p, r0.v)", "fused_s(s=r-alpha*v, t=A*s, s.t, t.t)",
    "cheb_blocked(several sweeps per pass)"};

struct KernelStat {
    double ms = 0;
    long long launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

struct Block {
    BlockGeom g;
    double *x = nullptr, *b = nullptr, *r = nullptr, *r0 = nullptr, *p = nullptr, *v = nullptr, *t = nullptr;
    double *mp = nullptr, *z = nullptr;              // alias p / r without a preconditioner (noneSolver.hpp:24-27)
    double *cy = nullptr, *cz = nullptr, *cw = nullptr;
    double *c4 = nullptr;                            // fourth iterate buffer of the temporally blocked Chebyshev schedule
    double *check_buf = nullptr;                     // per-CTA partial sums / maxima of pps_check_solution
    double theta = 0, delta = 0, sigma = 0;          // Chebyshev constants of the PRECONDITIONER on this block (global, or the block's own)
    double *x_saved = nullptr, *b_saved = nullptr;
    double *p2 = nullptr, *v2 = nullptr, *s = nullptr;   // PPS_FUSE_FULL: ping-pong p / v, separate s
    // nested (block-local) Krylov preconditioner: its own work vectors, control block and residual history
    double *ip = nullptr, *ir = nullptr, *ir0 = nullptr, *iv = nullptr, *it = nullptr, *iz = nullptr;
    Ctl* ictl = nullptr;                 // device
    double* ihist_host = nullptr;        // host-mapped, 4 x (precond_max_iter + 2): err, alpha, omega, rho
    double* dudn[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double* sendbuf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double* recvbuf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::vector<double*> owned;
};

}  // namespace pps

using namespace pps;

struct pps_handle {
    pps_config cfg{};
    int rank = 0, world = 1, device = 0;
    bool parity = false;
    int by = 8;                 // tile rows of the plain-load kernels
    int stencil_impl = 1;       // 0 plain loads, 1 TMA ring (stencil_tma.cuh)
    int by_tma = 8;             // tile rows of the TMA operator kernels (8 or 16)
    int by_hint = 0;            // tile rows of the operator kernel whose tiling is being made, 0 = by_tma (PPS_FUSE_BY_S: fused_s sweep)
    int fuse_by_s = 16;
    int zchunk_hint = 0;        // z-chunk of the operator kernel whose tiling is being made, 0 = default (see make_tiling)
    int tma_l2_promo = 3;       // PPS_TMA_L2PROMO: 0 none, 1 64 B, 2 128 B, 3 256 B (tuning sweeps)
    int zchunk_fused_p = 0, zchunk_fused_s = 0;   // PPS_ZCHUNK_FUSED_P / _S: z-chunk of the two fused kernels (0 = the operator kernels' value)
    std::map<std::pair<const void*, int>, CUtensorMap> tmaps;   // key: (field, rows) main box; (field, -rows) aux box
    std::set<const void*> smem_opt_in;   // kernels whose dynamic shared-memory limit was raised on THIS device
    int zchunk_stencil = 0, zchunk_point = 0;   // 0 = heuristic
    int zchunk_cheb = 0;                        // PPS_ZCHUNK_CHEB: output planes per CTA of the blocked Chebyshev kernel (0 = heuristic)
    int lag = 3;
    bool operator_only = false; // PPS_FLAG_OPERATOR_ONLY: only p, v, r0 exist
    bool fuse_full = false;     // 17-pass schedule (single block, all-Dirichlet, no preconditioner)
    bool fuse_p = true, fuse_s = true;   // PPS_FUSE_P / PPS_FUSE_S: enable the two fused kernels separately (diagnostics)
    int fuse_stages_p = 3, fuse_stages_s = 4;   // ring depth of the two fused kernels (PPS_FUSE_STAGES_P = 3|4, PPS_FUSE_STAGES_S = 3|4|6: tuning sweeps)
    int fuse_check = 0;                  // PPS_FUSE_CHECK=1 (with PPS_FUSE_P=0): verify every fused_s launch against the split kernels
    int fuse_check_events = 0;
    int iter_in_solve = 0;      // host-side count of enqueued iterations of the running solve
    std::vector<Block> blocks;
    cudaStream_t stream = nullptr;
    Ctl* ctl = nullptr;         // device
    Ctl* actl = nullptr;        // control block the launchers hand to the kernels: ctl, or a block's ictl inside a nested solve
    bool local_reduce = false;  // nested solve: reductions stay on this block (no allreduce), the last CTA applies the scalar op
    Ctl ctl_host{};
    int precond_max_iter = 150;          // iterMaxPreconditioner, solverSetup.hpp:32
    double precond_tolerance = 1e-6;     // tollPreconditionerSolver * tollScalingFactor, solverSetup.hpp:31
    long long precond_iters = 0;         // nested iterations of the last solve, summed over calls and local blocks
    std::vector<cudaEvent_t> inner_events;
    double* hist_host[4] = {nullptr, nullptr, nullptr, nullptr};   // host-mapped
    double* hist_dev[4] = {nullptr, nullptr, nullptr, nullptr};
    int hist_len = 0;
    double* partials = nullptr;
    long long partial_capacity = 0;
    unsigned int* counter = nullptr;
    ncclComm_t comm = nullptr;        // scalar allreduces, on the compute stream
    ncclComm_t comm_halo = nullptr;   // face exchange, on the (high-priority) halo stream when overlap is on
    cudaStream_t halo_stream = nullptr;
    cudaStream_t bnd_stream = nullptr;   // boundary-shell launches of an overlapped operator, concurrent with the interior launch
    cudaStream_t launch_stream = nullptr;   // stream the operator launchers use right now (stream or bnd_stream)
    cudaEvent_t ev_pre = nullptr, ev_bnd_done = nullptr;
    cudaEvent_t ev_field_ready = nullptr, ev_halo_done = nullptr;
    int overlap = 1;                  // 0 serial; 1 interior first, boundary chunks after the exchange; 2 in-kernel wait over NCCL
                                      // (experimental, can deadlock); 3 in-kernel wait over the SM-free peer transport (experimental)
    unsigned int* halo_flag = nullptr;   // device: epoch of the last exchange that has landed
    unsigned int halo_epoch = 0;
    HaloWait wait_next{nullptr, 0, -1, -1, 0};   // consumed by the next TMA operator launch
    int debug_no_halo = 0;            // timing experiments only: skip the exchange (wrong results)
    int cheb_block = 0;               // sweeps per pass of the temporally blocked Chebyshev kernel (0: one kernel per sweep)
    bool cheb_f32 = false;            // mixed-precision preconditioner: iterates in fp32 (alpaka tree, T_data_chebyshev = float)
    bool precond_comm = false;        // Chebyshev preconditioner with communicationON: faces of B and of every iterate are exchanged
    bool cheb_eig_local = false;      // block-local, not rescaled eigenvalue bounds (alpaka tree inputParam.hpp:21-22 `local`)
    int batch_ghosts = 1;             // all Neumann faces of a block in one launch (PPS_BATCH_GHOSTS=0: one launch per face)
    // PPS_GRAPH (default for launch-bound problems on one GPU): one Krylov iteration is captured into a CUDA graph at its first launch and replayed;
    // every iteration enqueues the same kernels with the same arguments (the scalars live in `ctl` on the device)
    bool use_graph = false;
    cudaGraphExec_t iter_graph = nullptr;
    long long iter_graph_launches = 0;   // kernels per replay (for pps_get_launch_count)
    // peer-memory halo path (PPS_HALO_P2P=1, z-slabs): neighbours' field / flag arrays mapped through CUDA IPC
    bool p2p = false;
    std::vector<const double*> p2p_local;            // my arrays that neighbours may push into (r, p, p2, v, v2, Mp, z)
    std::vector<double*> p2p_peer[2];                // the same arrays of my lower / upper z-neighbour, mapped here
    unsigned int* peer_flags[2] = {nullptr, nullptr};                      // neighbour's recv_epoch array
    unsigned int* recv_epoch = nullptr;                                    // mine: [exchange slot][face lo/hi], written by the neighbours
    unsigned int field_epoch[2] = {0, 0};            // per exchange slot: 0 = before the first operator of an iteration, 1 = before the second
    unsigned int* epoch_ring = nullptr;   // pinned host: source words of the flag DMAs (the host runs a few exchanges ahead)
    unsigned int epoch_ring_next = 0;
    std::vector<void*> ipc_opened;
    // peer-memory allreduce fused into the reducing kernels (PPS_ALLREDUCE_P2P=1)
    bool ar_p2p = false;
    double* ar_mail = nullptr;            // my mailbox [2][world][4]
    double** ar_peer_table = nullptr;     // device: world mailbox pointers
    unsigned int ar_epoch = 0;
    Coef coef{};
    // Chebyshev constants (chebyshevIteration.hpp:22-26)
    double theta = 0, delta = 0, sigma = 0;
    // results
    int iters = 0;
    double err_iter = -1, err_op = -1, norm_b = 1, solver_seconds = 0, loop_seconds = 0;
    long long launch_count = 0;
    bool profiling = false;
    int profile_only = -1;      // >= 0: only this kernel class is bracketed with events
    KernelStat stats[KC_COUNT];
    std::vector<cudaEvent_t> event_pool;
    size_t event_next = 0;
    std::vector<cudaEvent_t> iter_events;
    cudaEvent_t ev_start = nullptr, ev_loop0 = nullptr, ev_loop1 = nullptr, ev_end = nullptr;
};

namespace pps {

static thread_local std::string g_last_error;

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
    const char* s = std::getenv(name);
    return s ? std::atoi(s) : dflt;
}

static double* dalloc(Block& b, long long n, cudaStream_t s) {
    double* p = nullptr;
    PPS_CUDA_CHECK(cudaMalloc(&p, sizeof(double) * static_cast<size_t>(n)));
    PPS_CUDA_CHECK(cudaMemsetAsync(p, 0, sizeof(double) * static_cast<size_t>(n), s));
    b.owned.push_back(p);
    return p;
}

static Block* find_block(pps_handle* h, int rank) {
    for (auto& b : h->blocks)
        if (b.g.rank == rank) return &b;
    throw std::runtime_error("block of rank " + std::to_string(rank) + " is not hosted by this handle");
}

static cudaEvent_t pool_event(pps_handle* h) {
    if (h->event_next == h->event_pool.size()) {
        cudaEvent_t e;
        PPS_CUDA_CHECK(cudaEventCreate(&e));
        h->event_pool.push_back(e);
    }
    return h->event_pool[h->event_next++];
}

// brackets the launches of one kernel class with events when profiling is on, and counts launches
struct LaunchScope {
    pps_handle* h;
    int kc;
    cudaEvent_t e0 = nullptr;
    bool on;
    LaunchScope(pps_handle* h_, int kc_) : h(h_), kc(kc_) {
        on = h->profiling && (h->profile_only < 0 || h->profile_only == kc);
        if (on) {
            e0 = pool_event(h);
            cudaEventRecord(e0, h->stream);
        }
    }
    void count(int n = 1) {
        h->launch_count += n;
        h->stats[kc].launches += n;
    }
    ~LaunchScope() {
        if (on) {
            cudaEvent_t e1 = pool_event(h);
            cudaEventRecord(e1, h->stream);
            h->stats[kc].pending.emplace_back(e0, e1);
        }
    }
};

static void collect_stats(pps_handle* h) {
    for (int k = 0; k < KC_COUNT; k++) {
        for (auto& pr : h->stats[k].pending) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) h->stats[k].ms += ms;
        }
        h->stats[k].pending.clear();
    }
    h->event_next = 0;
}

static void check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + " launch failed: " + cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------------------
// launch geometry
// ------------------------------------------------------------------------------------------------
struct Tiling {
    dim3 grid, block;
    int zchunk;
    TileOrigin org;
    unsigned int ctas() const { return grid.x * grid.y * grid.z; }
};

// CTA grid that covers `box` (any sub-box of the block): 64-wide x tiles, `by` rows, z-chunks
static Tiling make_tiling(const pps_handle* h, const BlockGeom& g, const Box& box, bool stencil) {
    Tiling t;
    const bool tma = stencil && h->stencil_impl == 1;
    const int by = tma ? (h->by_hint > 0 ? h->by_hint : h->by_tma) : h->by;
    t.block = dim3(32, tma ? by + 1 : by, 1);
    const int bx0 = (std::max(box.i0, 1) - 1) / 64, bx1 = (std::min(box.i1, g.n[0] + 1) - 2) / 64 + 1;
    const int by0 = (std::max(box.j0, 1) - 1) / by, by1 = (std::min(box.j1, g.n[1] + 1) - 2) / by + 1;
    const int gx = std::max(1, bx1 - bx0), gy = std::max(1, by1 - by0);
    const int nzb = std::max(1, box.k1 - box.k0);
    int zc = stencil ? h->zchunk_stencil : h->zchunk_point;
    if (zc <= 0) {
        if (stencil) {
            // measured on B200 at 512^3 (profiles/README.md, r01 / r02 sweeps): fused_s 16 / 32 / 43 / 64 / 85 / 128 planes per chunk
            // -> 0.814 / 0.799 / 0.810 / 0.817 / 0.859 / 0.892 ms, plain operator alike: short chunks win although every chunk
            // re-reads two halo planes (they are still in L2 from the neighbouring chunk; z is the slowest grid dimension).  A model
            // of wave quantisation + halo overhead picked 85 / 128 and was measurably worse.  fused_p (three halo'd inputs, its ring
            // limits the SM to 2-3 CTAs) is the exception: 1.094 / 1.076 / 1.063 / 1.085 ms at 32 / 48 / 85 / 102 planes (3 ring stages) -> `zchunk_hint`.
            zc = h->zchunk_hint > 0 ? h->zchunk_hint : 32;
        } else {
            // pointwise kernels have no halo: ~6 waves of CTAs (148 SMs x 8 CTAs of 256 threads), chunks >= 8 planes
            const long long target = 148LL * 8 * 6;
            long long nz_chunks = std::max<long long>(1, target / std::max(1, gx * gy));
            zc = std::max(static_cast<int>((nzb + nz_chunks - 1) / nz_chunks), 8);
        }
    }
    zc = std::min(zc, nzb);
    if (stencil) {
        // balanced chunks: never a sliver at the end (a 1-plane chunk still streams 3 planes)
        const int nch = (nzb + zc - 1) / zc;
        zc = (nzb + nch - 1) / nch;
    }
    t.zchunk = zc;
    t.grid = dim3(gx, gy, (nzb + zc - 1) / zc);
    t.org = TileOrigin{bx0, by0, box.k0};
    return t;
}

static RedCtx make_red(pps_handle* h, int nacc, unsigned int total, unsigned int offset, int op) {
    RedCtx r;
    r.partials = h->partials;
    r.counter = h->counter;
    r.capacity = h->partial_capacity;
    r.total_ctas = total;
    r.cta_offset = offset;
    r.nacc = nacc;
    r.op = (h->world > 1 && !h->local_reduce) ? static_cast<int>(OP_NONE) : op;
    r.ctl = h->actl;
    r.pr = PeerReduce{nullptr, nullptr, 0, 0, 0};
    if (h->ar_p2p && !h->local_reduce && nacc > 0 && op != OP_NONE) {
        // the allreduce happens inside the reducing kernel: every launch that feeds this reduction carries the same epoch
        // (finish_reduction bumps it once the launches of the reduction have been issued)
        r.op = op;
        r.pr = PeerReduce{h->ar_peer_table, h->ar_mail, h->ar_epoch + 1, h->world, h->rank};
    }
    if (total > h->partial_capacity) throw std::runtime_error("partials buffer too small");
    return r;
}

// after a fused reduction: allreduce the raw sums over NVLink and apply the scalar update (world > 1 only)
static void finish_reduction(pps_handle* h, int nacc, int op, bool ignore_done) {
    if (h->world == 1 || h->local_reduce) return;
    if (h->ar_p2p && op != OP_NONE) {
        h->ar_epoch++;   // done in-kernel (PeerReduce); next reduction, next epoch
        return;
    }
    LaunchScope ls(h, KC_SCALAR);
    // (the active control block: `ctl`, or the nested one inside a GLOBAL nested Krylov preconditioner)
    PPS_NCCL_CHECK(nccl().AllReduce(h->actl->sums, h->actl->sums, nacc, ncclDouble, ncclSum, h->comm, h->stream));
    scalar_op_kernel<<<1, 1, 0, h->stream>>>(op, h->actl, ignore_done ? 1 : 0);
    ls.count(1);
}

// 3-D tensor map of one field (pitch x (ny+2) x (nz+2) doubles), box = halo'd tile of one plane
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {   // thread-safe static initialisation (rank threads of the C++ driver)
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        PPS_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled not available");
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

static const CUtensorMap& tensor_map(pps_handle* h, const Block& b, const double* field, int by, bool aux) {
    auto key = std::make_pair(static_cast<const void*>(field), aux ? -by : by);
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) return it->second;
    CUtensorMap m;
    const Dims& d = b.g.dims;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(d.pitch), static_cast<cuuint64_t>(d.ny + 2), static_cast<cuuint64_t>(d.nz + 2)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(d.pitch) * 8, static_cast<cuuint64_t>(d.plane) * 8};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(aux ? 64 : kTmaBoxX), static_cast<cuuint32_t>(aux ? by : by + 2), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    static const CUtensorMapL2promotion promo[4] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_64B,
                                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
    CUresult r = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(field), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo[h->tma_l2_promo & 3],
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
    return h->tmaps.emplace(key, m).first->second;
}

template <int BY, int STAGES, bool PAR, class Epi>
static void launch_tma_inst(pps_handle* h, const Block& b, const double* u, const Box& box, const Epi& epi,
                            const RedCtx& red, const Tiling& t, const Ctl* ctl) {
    auto kern = stencil_tma_kernel<BY, STAGES, PAR, Epi>;
    constexpr int smem = TmaSmem<BY, STAGES, Epi::NAUX>::kBytes;
    // the opt-in above 48 KB is a per-device function attribute: remember it per handle (= per device), not per process
    if (h->smem_opt_in.insert(reinterpret_cast<const void*>(kern)).second)
        PPS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const CUtensorMap& tm = tensor_map(h, b, u, BY, false);
    const CUtensorMap& a0 = Epi::NAUX > 0 ? tensor_map(h, b, epi.aux(0), BY, true) : tm;
    const CUtensorMap& a1 = Epi::NAUX > 1 ? tensor_map(h, b, epi.aux(1), BY, true) : tm;
    kern<<<t.grid, t.block, smem, h->launch_stream>>>(tm, a0, a1, b.g.dims, box, h->coef, t.zchunk, t.org, h->wait_next, epi, red, ctl);
    h->wait_next = HaloWait{nullptr, 0, -1, -1, 0};
}

template <int BY, int STAGES, bool PAR, int DBG = 0, class Pre, class Epi>
static void launch_tma_pre(pps_handle* h, int kc, const Block& b, const Box& box, const Pre& pre, const Epi& epi, const RedCtx& red,
                           const Tiling& t, bool check_done) {
    LaunchScope ls(h, kc);
    auto kern = stencil_tma_pre_kernel<BY, STAGES, PAR, Pre, Epi, DBG>;
    constexpr int smem = TmaPreSmem<BY, STAGES, Pre::NIN, Epi::NAUX>::kBytes;
    if (h->smem_opt_in.insert(reinterpret_cast<const void*>(kern)).second)
        PPS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const CUtensorMap& i0 = tensor_map(h, b, pre.input(0), BY, false);
    const CUtensorMap& i1 = Pre::NIN > 1 ? tensor_map(h, b, pre.input(1), BY, false) : i0;
    const CUtensorMap& i2 = Pre::NIN > 2 ? tensor_map(h, b, pre.input(2), BY, false) : i0;
    const CUtensorMap& a0 = Epi::NAUX > 0 ? tensor_map(h, b, epi.aux(0), BY, true) : i0;
    const CUtensorMap& a1 = Epi::NAUX > 1 ? tensor_map(h, b, epi.aux(1), BY, true) : i0;
    kern<<<t.grid, t.block, smem, h->launch_stream>>>(i0, i1, i2, a0, a1, b.g.dims, box, h->coef, t.zchunk, t.org, pre, epi, red,
                                                       check_done ? h->actl : nullptr);
    check_launch(kKernelNames[kc]);
    ls.count(1);
}

template <class Epi>
static void launch_stencil(pps_handle* h, int kc, const Block& b, const double* u, const Box& box, const Epi& epi,
                           const RedCtx& red, const Tiling& t, bool check_done) {
    LaunchScope ls(h, kc);
    const Ctl* ctl = check_done ? h->actl : nullptr;
    if (h->stencil_impl == 1) {
        if (h->by_tma == 16) {
            if (h->parity) launch_tma_inst<16, 4, true>(h, b, u, box, epi, red, t, ctl);
            else           launch_tma_inst<16, 4, false>(h, b, u, box, epi, red, t, ctl);
        } else {
            if (h->parity) launch_tma_inst<8, 6, true>(h, b, u, box, epi, red, t, ctl);
            else           launch_tma_inst<8, 6, false>(h, b, u, box, epi, red, t, ctl);
        }
    } else {
#define PPS_LAUNCH_ST(BYV, PAR) \
    stencil_kernel<BYV, PAR, Epi><<<t.grid, t.block, 0, h->launch_stream>>>(u, b.g.dims, box, h->coef, t.zchunk, t.org, epi, red, ctl)
        if (h->by == 4) { if (h->parity) PPS_LAUNCH_ST(4, true); else PPS_LAUNCH_ST(4, false); }
        else            { if (h->parity) PPS_LAUNCH_ST(8, true); else PPS_LAUNCH_ST(8, false); }
#undef PPS_LAUNCH_ST
    }
    check_launch(kKernelNames[kc]);
    ls.count(1);
}

template <class Op>
static void launch_pointwise(pps_handle* h, int kc, const Block& b, const Box& box, const Op& op, const RedCtx& red,
                             const Tiling& t, bool check_done) {
    LaunchScope ls(h, kc);
    // pointwise ops read their scalars from ctl in begin(), so they always get ctl and therefore always honour ctl->done;
    // they are only launched inside the iteration or during set-up, where done == 0 (`check_done` documents the call site)
    const Ctl* ctl = h->actl;
    (void)check_done;
    if (h->by == 4) pointwise_kernel<4, Op><<<t.grid, t.block, 0, h->stream>>>(b.g.dims, box, t.zchunk, t.org, op, red, ctl);
    else            pointwise_kernel<8, Op><<<t.grid, t.block, 0, h->stream>>>(b.g.dims, box, t.zchunk, t.org, op, red, ctl);
    check_launch(kKernelNames[kc]);
    ls.count(1);
}

// ------------------------------------------------------------------------------------------------
// faces
// ------------------------------------------------------------------------------------------------
// plane index along the face normal: 0 guard, 1 boundary data plane, 2 first interior plane (mirror)
static FaceGeom face_geom(const BlockGeom& g, int face, int plane_a, int plane_b) {
    const int d = face / 2, up = face % 2;
    auto coord = [&](int which) { return up ? g.n[d] + 1 - which : which; };
    int u, v;
    g.tangential(face, u, v);
    int ia[3] = {1, 1, 1}, ib[3] = {1, 1, 1};
    ia[d] = coord(plane_a);
    ib[d] = coord(plane_b);
    FaceGeom f;
    f.base_a = g.at(ia[0], ia[1], ia[2]);
    f.base_b = g.at(ib[0], ib[1], ib[2]);
    f.stride_u = g.stride(u);
    f.stride_v = g.stride(v);
    f.nu = g.n[u];
    f.nv = g.n[v];
    return f;
}

static int face_blocks(const FaceGeom& f) {
    const long long n = static_cast<long long>(f.nu) * f.nv;
    return static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8));
}

using FieldSel = double* (*)(Block&);
static double* sel_x(Block& b) { return b.x; }
static double* sel_p(Block& b) { return b.p; }
static double* sel_mp(Block& b) { return b.mp; }
static double* sel_z(Block& b) { return b.z; }
static double* sel_t(Block& b) { return b.t; }
static double* sel_cy(Block& b) { return b.cy; }

static double* sel_r(Block& b) { return b.r; }
static double* sel_v(Block& b) { return b.v; }

// up to three fields that travel together (the fused schedule exchanges the INPUTS of the fused operand: r, p, v)
struct FieldSet {
    FieldSel sel[3] = {nullptr, nullptr, nullptr};
    int n = 0;
    FieldSet() {}
    FieldSet(FieldSel a) : n(1) { sel[0] = a; }
    FieldSet(FieldSel a, FieldSel b) : n(2) { sel[0] = a; sel[1] = b; }
    FieldSet(FieldSel a, FieldSel b, FieldSel c) : n(3) { sel[0] = a; sel[1] = b; sel[2] = c; }
};
constexpr int kMaxExchangeFields = 3;

// CommunicatorMPI::operator() + waitAllandCheckRcv (communicationMPI.hpp:51-316) for every field of `fs`; all faces of all
// fields travel in ONE NCCL group.  `on_halo_stream`: issue the exchange on the high-priority halo stream with its own
// communicator (overlap with interior compute).
static void halo_exchange(pps_handle* h, const FieldSet& fs, bool check_done, bool on_halo_stream = false) {
    bool any = false;
    for (auto& b : h->blocks)
        for (int f = 0; f < 6; f++) any = any || b.g.hc[f];
    if (!any || h->debug_no_halo || fs.n == 0) return;
    LaunchScope ls(h, KC_HALO);
    const int ign = check_done ? 0 : 1;
    if (h->world == 1) {
        for (auto& b : h->blocks) {
            for (int f = 0; f < 6; f++) {
                if (!b.g.hc[f]) continue;
                Block* nb = find_block(h, b.g.nbr[f]);
                // my guard plane on face f <- neighbour's boundary data plane on the opposite face
                FaceGeom mine = face_geom(b.g, f, 0, 0);
                FaceGeom theirs = face_geom(nb->g, f ^ 1, 1, 1);
                FaceGeom g = mine;
                g.base_b = theirs.base_b;
                for (int q = 0; q < fs.n; q++) {
                    face_copy_kernel<<<face_blocks(g), 256, 0, h->stream>>>(fs.sel[q](b), fs.sel[q](*nb), g, h->ctl, ign);
                    ls.count(1);
                }
            }
        }
    } else {
        cudaStream_t st = on_halo_stream ? h->halo_stream : h->stream;
        ncclComm_t comm = on_halo_stream ? h->comm_halo : h->comm;
        Block& b = h->blocks[0];
        auto face_count = [&](int f) { return static_cast<size_t>(b.g.n[f / 2 == 0 ? 1 : 0]) * b.g.n[2]; };
        for (int q = 0; q < fs.n; q++) {
            for (int f = 0; f < 4; f++) {   // x and y faces are strided: pack first
                if (!b.g.hc[f]) continue;
                FaceGeom g = face_geom(b.g, f, 1, 1);
                face_pack_kernel<<<face_blocks(g), 256, 0, st>>>(b.sendbuf[f] + q * face_count(f), fs.sel[q](b), g, h->ctl, ign);
                ls.count(1);
            }
        }
        PPS_NCCL_CHECK(nccl().GroupStart());
        for (int f = 0; f < 6; f++) {
            if (!b.g.hc[f]) continue;
            const int peer = b.g.nbr[f];
            if (f < 4) {
                const size_t cnt = face_count(f) * fs.n;   // the fields of one face are packed back to back
                PPS_NCCL_CHECK(nccl().Send(b.sendbuf[f], cnt, ncclDouble, peer, comm, st));
                PPS_NCCL_CHECK(nccl().Recv(b.recvbuf[f], cnt, ncclDouble, peer, comm, st));
            } else {
                // z faces: a k-plane of the pitched layout is contiguous, padding and all -- no packing
                const int up = f % 2;
                const long long kdata = up ? b.g.n[2] : 1, kguard = up ? b.g.n[2] + 1 : 0;
                for (int q = 0; q < fs.n; q++) {
                    double* fld = fs.sel[q](b);
                    PPS_NCCL_CHECK(nccl().Send(fld + kdata * b.g.dims.plane, b.g.dims.plane, ncclDouble, peer, comm, st));
                    PPS_NCCL_CHECK(nccl().Recv(fld + kguard * b.g.dims.plane, b.g.dims.plane, ncclDouble, peer, comm, st));
                }
            }
        }
        PPS_NCCL_CHECK(nccl().GroupEnd());
        for (int q = 0; q < fs.n; q++) {
            for (int f = 0; f < 4; f++) {
                if (!b.g.hc[f]) continue;
                FaceGeom g = face_geom(b.g, f, 0, 0);
                face_unpack_kernel<<<face_blocks(g), 256, 0, st>>>(fs.sel[q](b), b.recvbuf[f] + q * face_count(f), g, h->ctl, ign);
                ls.count(1);
            }
        }
    }
    check_launch("halo");
}
static void halo_exchange(pps_handle* h, FieldSel sel, bool check_done, bool on_halo_stream = false) {
    halo_exchange(h, FieldSet(sel), check_done, on_halo_stream);
}

// ------------------------------------------------------------------------------------------------
// Peer-memory halo path for z-slabs (opt-in, PPS_HALO_P2P=1).  An exchange of field f is, per neighbour, one
// copy-engine cudaMemcpyAsync of my boundary data plane straight into the neighbour's guard plane followed by a
// one-thread st.release.sys of the exchange epoch into the neighbour's flag; the consumer side waits for its own
// flags with a one-thread kernel on the boundary stream.  No NCCL kernel runs on the SMs and no packing is needed
// (a z-plane of the pitched layout is contiguous).  Re-use of a guard plane is safe without a credit message: two
// exchanges of the same field are always separated by the scalar allreduces of the iteration, which every rank
// only passes after its operator launches that read the previous content have completed.
// ------------------------------------------------------------------------------------------------
constexpr unsigned int kEpochRing = 4096;   // flag DMAs in flight are bounded by the host's run-ahead (PPS_LAG iterations)

constexpr int kP2pMaxFields = 8;
struct PeerInfo {
    long long pid;
    int device;
    int nfields;
    int ok;
    int pad;
    cudaIpcMemHandle_t field[kP2pMaxFields], flags;
    unsigned long long raw_field[kP2pMaxFields], raw_flags;
};

// Every rank offers the arrays whose guard planes its z-neighbours fill (same list, same order on every rank) and maps its
// neighbours' arrays.  Any failure (CUDA IPC not permitted, no peer access) on ANY rank switches the transport off on ALL
// ranks -- the NCCL exchange stays the fallback -- so the ranks can never disagree about which path an exchange takes.
static void setup_p2p(pps_handle* h) {
    Block& b = h->blocks[0];
    for (int f = 0; f < 4; f++)
        if (b.g.hc[f]) return;   // x / y faces would need packing: NCCL path
    if (!(b.g.hc[4] || b.g.hc[5]) || h->cfg.solver != PPS_SOLVER_BICGSTAB) return;
    std::vector<double*> offer = {b.r, b.p, b.p2, b.v, b.v2, b.mp, b.z};
    std::vector<double*> uniq;
    for (double* q : offer)
        if (q != nullptr && std::find(uniq.begin(), uniq.end(), q) == uniq.end()) uniq.push_back(q);
    const size_t sz = sizeof(PeerInfo);
    if (sz * h->world > sizeof(double) * kMaxAcc * static_cast<size_t>(h->partial_capacity)) return;
    PeerInfo mine{};
    mine.pid = static_cast<long long>(getpid());
    mine.device = h->device;
    mine.nfields = static_cast<int>(uniq.size());
    mine.ok = 1;
    if (cudaMalloc(&h->recv_epoch, 4 * sizeof(unsigned int)) != cudaSuccess) { mine.ok = 0; h->recv_epoch = nullptr; }
    if (mine.ok) PPS_CUDA_CHECK(cudaMemsetAsync(h->recv_epoch, 0, 4 * sizeof(unsigned int), h->stream));
    if (mine.ok && cudaHostAlloc(&h->epoch_ring, kEpochRing * sizeof(unsigned int), cudaHostAllocDefault) != cudaSuccess) { mine.ok = 0; h->epoch_ring = nullptr; }
    for (int q = 0; mine.ok && q < mine.nfields; q++) {
        if (cudaIpcGetMemHandle(&mine.field[q], uniq[q]) != cudaSuccess) mine.ok = 0;
        mine.raw_field[q] = reinterpret_cast<unsigned long long>(uniq[q]);
    }
    if (mine.ok && cudaIpcGetMemHandle(&mine.flags, h->recv_epoch) != cudaSuccess) mine.ok = 0;
    mine.raw_flags = reinterpret_cast<unsigned long long>(h->recv_epoch);
    cudaGetLastError();
    char* scratch = reinterpret_cast<char*>(h->partials);
    PPS_CUDA_CHECK(cudaMemcpyAsync(scratch + sz * h->rank, &mine, sz, cudaMemcpyHostToDevice, h->stream));
    PPS_NCCL_CHECK(nccl().AllGather(scratch + sz * h->rank, scratch, sz, ncclChar, h->comm, h->stream));
    std::vector<PeerInfo> all(h->world);
    PPS_CUDA_CHECK(cudaMemcpyAsync(all.data(), scratch, sz * h->world, cudaMemcpyDeviceToHost, h->stream));
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    bool ok = true;
    for (auto& pi : all) ok = ok && pi.ok && pi.nfields == mine.nfields;
    for (int up = 0; ok && up < 2; up++) {
        if (!b.g.hc[4 + up]) continue;
        const PeerInfo& pi = all[b.g.nbr[4 + up]];
        h->p2p_peer[up].assign(uniq.size(), nullptr);
        if (pi.pid == mine.pid) {
            // rank-threads of one process (C++ driver): plain peer access
            if (pi.device != h->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(pi.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
                cudaGetLastError();
            }
            for (size_t q = 0; q < uniq.size(); q++) h->p2p_peer[up][q] = reinterpret_cast<double*>(pi.raw_field[q]);
            h->peer_flags[up] = reinterpret_cast<unsigned int*>(pi.raw_flags);
        } else {
            void* ptr = nullptr;
            for (size_t q = 0; ok && q < uniq.size(); q++) {
                if (cudaIpcOpenMemHandle(&ptr, pi.field[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
                h->p2p_peer[up][q] = static_cast<double*>(ptr); h->ipc_opened.push_back(ptr);
            }
            if (ok && cudaIpcOpenMemHandle(&ptr, pi.flags, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) {
                h->peer_flags[up] = static_cast<unsigned int*>(ptr); h->ipc_opened.push_back(ptr);
            } else {
                ok = false;
            }
            cudaGetLastError();
        }
    }
    // agree: one more collective (also the barrier after which nobody's flags are cleared any more)
    double flag = ok ? 0.0 : 1.0;
    PPS_CUDA_CHECK(cudaMemcpyAsync(h->ctl->sums + 7, &flag, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    PPS_NCCL_CHECK(nccl().AllReduce(h->ctl->sums + 7, h->ctl->sums + 7, 1, ncclDouble, ncclSum, h->comm, h->stream));
    PPS_CUDA_CHECK(cudaMemcpyAsync(&flag, h->ctl->sums + 7, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (flag != 0.0) {
        if (h->rank == 0 && env_int("PPS_VERBOSE", 0)) std::fprintf(stderr, "[pps] peer-memory halo transport not available: using NCCL send/recv\n");
        return;
    }
    h->p2p_local.assign(uniq.begin(), uniq.end());
    h->p2p = true;
}

static int p2p_index(const pps_handle* h, const double* f) {
    for (size_t q = 0; q < h->p2p_local.size(); q++)
        if (h->p2p_local[q] == f) return static_cast<int>(q);
    return -1;
}

// mailboxes of the in-kernel allreduce: exchange IPC handles with EVERY rank
static void setup_allreduce_p2p(pps_handle* h) {
    struct Info { long long pid; int device; int pad; cudaIpcMemHandle_t mail; unsigned long long raw; };
    const size_t bytes = sizeof(double) * 2 * 4 * static_cast<size_t>(h->world);
    PPS_CUDA_CHECK(cudaMalloc(&h->ar_mail, bytes));
    PPS_CUDA_CHECK(cudaMemsetAsync(h->ar_mail, 0, bytes, h->stream));
    Info mine{};
    mine.pid = static_cast<long long>(getpid());
    mine.device = h->device;
    PPS_CUDA_CHECK(cudaIpcGetMemHandle(&mine.mail, h->ar_mail));
    mine.raw = reinterpret_cast<unsigned long long>(h->ar_mail);
    const size_t sz = sizeof(Info);
    char* scratch = reinterpret_cast<char*>(h->partials);
    PPS_CUDA_CHECK(cudaMemcpyAsync(scratch + sz * h->rank, &mine, sz, cudaMemcpyHostToDevice, h->stream));
    PPS_NCCL_CHECK(nccl().AllGather(scratch + sz * h->rank, scratch, sz, ncclChar, h->comm, h->stream));
    std::vector<Info> all(h->world);
    PPS_CUDA_CHECK(cudaMemcpyAsync(all.data(), scratch, sz * h->world, cudaMemcpyDeviceToHost, h->stream));
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    std::vector<double*> table(h->world, nullptr);
    for (int r = 0; r < h->world; r++) {
        if (r == h->rank) { table[r] = h->ar_mail; continue; }
        if (all[r].pid == mine.pid) {
            if (all[r].device != h->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) PPS_CUDA_CHECK(e);
                cudaGetLastError();
            }
            table[r] = reinterpret_cast<double*>(all[r].raw);
        } else {
            void* q = nullptr;
            PPS_CUDA_CHECK(cudaIpcOpenMemHandle(&q, all[r].mail, cudaIpcMemLazyEnablePeerAccess));
            table[r] = static_cast<double*>(q);
            h->ipc_opened.push_back(q);
        }
    }
    PPS_CUDA_CHECK(cudaMalloc(&h->ar_peer_table, sizeof(double*) * h->world));
    PPS_CUDA_CHECK(cudaMemcpyAsync(h->ar_peer_table, table.data(), sizeof(double*) * h->world, cudaMemcpyHostToDevice, h->stream));
    // nobody may post into a mailbox before it has been cleared: one more collective as a barrier
    PPS_NCCL_CHECK(nccl().AllReduce(h->ctl->sums + 7, h->ctl->sums + 7, 1, ncclDouble, ncclSum, h->comm, h->stream));
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    h->ar_p2p = true;
}

// push my boundary planes of every field of `fs` into the neighbours' guard planes on the halo stream; `slot` names the exchange
// (0: before the first operator of an iteration, 1: before the second) and with it the pair of epoch flags
static unsigned int halo_push_p2p(pps_handle* h, int slot, const FieldSet& fs) {
    Block& b = h->blocks[0];
    const unsigned int epoch = ++h->field_epoch[slot];
    LaunchScope ls(h, KC_HALO);
    for (int up = 0; up < 2; up++) {
        if (!b.g.hc[4 + up]) continue;
        const long long kdata = up ? b.g.n[2] : 1;                 // my boundary data plane
        const long long kguard = up ? 0 : b.g.n[2] + 1;            // the neighbour's guard plane that faces me
        for (int q = 0; q < fs.n; q++) {
            double* fld = fs.sel[q](b);
            double* peer = h->p2p_peer[up][p2p_index(h, fld)];
            PPS_CUDA_CHECK(cudaMemcpyAsync(peer + kguard * b.g.dims.plane, fld + kdata * b.g.dims.plane,
                                           sizeof(double) * static_cast<size_t>(b.g.dims.plane), cudaMemcpyDefault, h->halo_stream));
        }
        // my upper neighbour receives on ITS lower face (slot 0), my lower neighbour on its upper face (slot 1).  The flag is a
        // last DMA (4 bytes from a pinned host word) ordered after the planes by the stream: no SM takes part in the
        // transport, so receivers may wait inside a running kernel (PPS_OVERLAP=3) without starving the sender.
        unsigned int* word = h->epoch_ring + (h->epoch_ring_next++ % kEpochRing);
        *word = epoch;
        PPS_CUDA_CHECK(cudaMemcpyAsync(h->peer_flags[up] + 2 * slot + (up ? 0 : 1), word, sizeof(unsigned int), cudaMemcpyDefault,
                                       h->halo_stream));
    }
    check_launch("halo_push_p2p");
    return epoch;
}

// resetNeumanBCs<isMainLoop, fieldData> (iterativeSolverBase.hpp:62-169)
static void neumann_ghosts(pps_handle* h, Block& b, double* field, bool with_value, bool check_done) {
    bool any = false;
    for (int f = 0; f < 6; f++) any = any || (b.g.hb[f] && h->cfg.bcs_type[f] == 1);
    if (!any) return;
    LaunchScope ls(h, KC_GHOST);
    // orderNeumanBcs == 2: ghost = first interior plane -/+ 2 ds g; == 1: ghost = boundary plane -/+ ds g  (iterativeSolverBase.hpp:92-108)
    const int src_plane = h->cfg.order_neumann == 1 ? 1 : 2;
    const double step = h->cfg.order_neumann == 1 ? 1.0 : 2.0;
    if (h->batch_ghosts) {
        GhostBatch batch{};
        int blocks = 1;
        for (int f = 0; f < 6; f++) {
            if (!(b.g.hb[f] && h->cfg.bcs_type[f] == 1)) continue;
            if (with_value && b.dudn[f] == nullptr)
                throw std::runtime_error("Neumann face " + std::to_string(f) + " has no du/dn values: call pps_set_neumann_face first");
            const int q = batch.count++;
            batch.g[q] = face_geom(b.g, f, 0, src_plane);
            batch.dudn[q] = with_value ? b.dudn[f] : nullptr;
            batch.two_ds[q] = step * h->cfg.ds[f / 2];
            batch.upper[q] = f % 2;
            blocks = std::max(blocks, face_blocks(batch.g[q]));
        }
        neumann_ghost_batch_kernel<<<dim3(blocks, batch.count), 256, 0, h->stream>>>(field, batch, h->actl, check_done ? 0 : 1);
        ls.count(1);
        check_launch("neumann_ghost_batch");
        return;
    }
    for (int f = 0; f < 6; f++) {
        if (!(b.g.hb[f] && h->cfg.bcs_type[f] == 1)) continue;
        FaceGeom g = face_geom(b.g, f, 0, src_plane);
        const double two_ds = step * h->cfg.ds[f / 2];
        if (with_value && b.dudn[f] == nullptr)
            throw std::runtime_error("Neumann face " + std::to_string(f) + " of rank " + std::to_string(b.g.rank) +
                                     " has no du/dn values: call pps_set_neumann_face first");
        neumann_ghost_kernel<<<face_blocks(g), 256, 0, h->stream>>>(field, g, with_value ? b.dudn[f] : nullptr, two_ds,
                                                                    f % 2, h->actl, check_done ? 0 : 1);
        ls.count(1);
    }
    check_launch("neumann_ghost");
}

// adjustFieldBForDirichletNeumanBCs (iterativeSolverBase.hpp:429-534) on a copy of b
static void adjust_b(pps_handle* h, Block& b, double* bcopy) {
    LaunchScope ls(h, KC_SETUP);
    for (int f = 0; f < 6; f++) {
        if (!b.g.hb[f]) continue;
        const int neumann = h->cfg.bcs_type[f] == 1;
        FaceGeom g = neumann ? face_geom(b.g, f, 1, 1) : face_geom(b.g, f, 2, 1);
        if (neumann && b.dudn[f] == nullptr)
            throw std::runtime_error("Neumann face " + std::to_string(f) + " has no du/dn values");
        adjust_b_kernel<<<face_blocks(g), 256, 0, h->stream>>>(bcopy, b.x, g, b.dudn[f], h->cfg.ds[f / 2], neumann, f % 2,
                                                               h->cfg.order_neumann == 1 ? 1.0 : 2.0);
        ls.count(1);
    }
    check_launch("adjust_b");
}

static unsigned int total_ctas(pps_handle* h, bool stencil) {
    unsigned int n = 0;
    for (auto& b : h->blocks) n += make_tiling(h, b.g, b.g.solver_box(), stencil).ctas();
    return n;
}

// ------------------------------------------------------------------------------------------------
// preconditioners
// ------------------------------------------------------------------------------------------------
// X = M(B) per block.  NONE: X aliases B (the reference memcpy's, noneSolver.hpp:24-27).
// CHEBYSHEV: chebyshevIteration.hpp:48-140 with communicationOFF; the iterates y_{n-1}, y_n the reference
// computes and discards (its pointer swaps leave y_{n-2} in fieldW, :114-125) are not computed and the
// final X = -W is folded into the last live sweep -- bit-identical output, 35 instead of 45 vector passes.
// coefficients of the alpaka tree's folded kernels, evaluated in T_data_chebyshev = float with the kernels' own expressions
// (kernelsAlpakaChebyshev.hpp:145-151 first sweep, :240-246,252 later sweeps; chebyshevIterationAlpaka.hpp:123-124,158-159 for rho).
// The alpaka tree defines delta with the opposite sign of the CPU tree (chebyshevIterationAlpaka.hpp:29 vs chebyshevIteration.hpp:23),
// so sigma and every rho change sign as well and the iterates are the same numbers.
struct AlpakaCheb {
    float theta, delta, sigma, rho_old, rho;
};
static AlpakaCheb alpaka_cheb_start(const Block& b) {
    AlpakaCheb a;
    const double theta = b.theta, delta = -b.delta;        // b.delta follows the CPU tree (negative)
    a.theta = static_cast<float>(theta);
    a.delta = static_cast<float>(delta);
    a.sigma = a.theta / a.delta;                           // float quotient of the two float members (chebyshevIterationAlpaka.hpp:34-35,595-604)
    a.rho_old = 1 / a.sigma;
    a.rho = 1 / (2 * a.sigma - a.rho_old);
    return a;
}

static void chebyshev_blocked(pps_handle* h, Block& b, double* X, double* B, bool check_done) {
    const int last = h->cfg.cheb_max_iter - 2;   // X = -y_last (the two dead sweeps of the reference are not computed)
    if (last < 1) throw std::runtime_error("chebyshevMax < 3 is not supported");
    const int depth = std::max(1, std::min(h->cheb_block > 0 ? h->cheb_block : 1, kChebMaxLev));
    const int np = (last + depth - 1) / depth;
    const Box box = b.g.solver_box();
    int nm[6];
    for (int f = 0; f < 6; f++) nm[f] = (b.g.hb[f] && h->cfg.bcs_type[f] == 1) ? (h->cfg.order_neumann == 1 ? 1 : 2) : 0;
    ChebPassDesc p{};
    p.form = h->cheb_f32 ? (h->parity ? CHEB_FORM_ALPAKA_F32_PARITY : CHEB_FORM_ALPAKA_F32_FAST)
                         : (h->parity ? CHEB_FORM_CPU_PARITY : CHEB_FORM_CPU_FAST);
    p.cf = h->coef;
    p.theta = b.theta; p.two_sigma = 2 * b.sigma; p.two_over_delta = 2 / b.delta;
    double rho_old = 1 / b.sigma;
    double rho = 1 / (2 * b.sigma - rho_old);
    p.c1 = 2 * rho / b.delta;
    AlpakaCheb a = alpaka_cheb_start(b);
    const float r0 = static_cast<float>(h->cfg.ds[0] * h->cfg.ds[0]), r1 = static_cast<float>(h->cfg.ds[1] * h->cfg.ds[1]),
                r2 = static_cast<float>(h->cfg.ds[2] * h->cfg.ds[2]);
    p.a_theta = a.theta;
    p.a_f0_first = 2 * a.rho / a.delta * (1 / r0 / a.theta);
    p.a_f1_first = 2 * a.rho / a.delta * (1 / r1 / a.theta);
    p.a_f2_first = 2 * a.rho / a.delta * (1 / r2 / a.theta);
    p.a_fc0_first = 2 * a.rho / a.delta * (2 - 2 * (1 / r0 + 1 / r1 + 1 / r2) / a.theta);
    // iterate buffers: (Y, Z) ping-pong between (cy, cz) and (cw, c4)
    void* bufs[2][2] = {{b.cy, b.cz}, {b.cw, b.c4}};
    int c0 = 0;
    for (int q = 0; q < np; q++) {
        const int nlev = last / np + (q < last % np ? 1 : 0);
        p.nlev = nlev;
        p.first = q == 0;
        p.last = q == np - 1;
        for (int l = 1; l <= nlev; l++) {
            const int c = c0 + l;
            if (c >= 2) {
                rho_old = rho;
                rho = 1 / (2 * b.sigma - rho_old);
                a.rho_old = a.rho;
                a.rho = 1 / (2 * a.sigma - a.rho_old);
            }
            p.rho[l] = rho; p.rho_old[l] = rho_old;
            p.a_f0[l] = a.rho * 2 / a.delta * (1 / r0);
            p.a_f1[l] = a.rho * 2 / a.delta * (1 / r1);
            p.a_f2[l] = a.rho * 2 / a.delta * (1 / r2);
            p.a_fB[l] = a.rho * 2 / a.delta;
            p.a_fZ[l] = -a.rho * a.rho_old;
            p.a_fc0[l] = a.rho * (2 * a.sigma - 4 / a.delta * (1 / r0 + 1 / r1 + 1 / r2));
        }
        p.B = B;
        p.Yin = bufs[(q + 1) & 1][0]; p.Zin = bufs[(q + 1) & 1][1];
        p.Yout = bufs[q & 1][0]; p.Zout = bufs[q & 1][1];
        p.X = X;
        // z-chunks: about two waves of CTAs, chunks not shorter than 4 x the overlap
        const ChebTile t0 = cheb_make_tile(nlev, 1, nm);
        const long long tiles = static_cast<long long>((b.g.n[0] + t0.wx - 1) / t0.wx) * ((b.g.n[1] + t0.wy - 1) / t0.wy);
        const int nzb = std::max(1, box.k1 - box.k0);
        long long nch = std::max<long long>(1, (2 * 148 + tiles - 1) / tiles);
        int zc = static_cast<int>((nzb + nch - 1) / nch);
        zc = std::max(zc, std::min(nzb, 8 * nlev));
        if (h->zchunk_cheb > 0) zc = std::min(nzb, h->zchunk_cheb);
        const ChebTile tl = cheb_make_tile(nlev, zc, nm);
        LaunchScope ls(h, KC_CHEB_BLOCKED);
        cheb_blocked_launch(h->stream, b.g.dims, box, tl, p, check_done ? h->actl : nullptr);
        check_launch("cheb_blocked");
        ls.count(1);
        c0 += nlev;
    }
}

static void chebyshev_local(pps_handle* h, Block& b, double* X, double* B, bool check_done) {
    if ((h->cheb_block > 0 || h->cheb_f32) && h->cfg.dim == 3) return chebyshev_blocked(h, b, X, B, check_done);
    const int m = h->cfg.cheb_max_iter;
    const Box box = b.g.solver_box();
    const Tiling t = make_tiling(h, b.g, box, true);
    RedCtx red = make_red(h, 0, 1, 0, OP_NONE);
    double rho_old = 1 / b.sigma;
    double rho = 1 / (2 * b.sigma - rho_old);
    neumann_ghosts(h, b, B, false, check_done);
    const int last = m - 2;   // X = -y_last; validate() guarantees chebyshevMax >= 3, i.e. last >= 1
    if (last < 1) throw std::runtime_error("chebyshevMax < 3 is not supported");
    const double c1 = 2 * rho / b.delta;
    double *Y = b.cy, *Z = b.cz, *W = b.cw;
    if (h->parity) {
        EpiChebFirst<true> e{Z, last == 1 ? X : Y, b.theta, 1.0 / b.theta, c1, last == 1 ? -1.0 : 1.0};
        launch_stencil(h, KC_CHEB_FIRST, b, B, box, e, red, t, check_done);
    } else {
        EpiChebFirst<false> e{Z, last == 1 ? X : Y, b.theta, 1.0 / b.theta, c1, last == 1 ? -1.0 : 1.0};
        launch_stencil(h, KC_CHEB_FIRST, b, B, box, e, red, t, check_done);
    }
    for (int c = 2; c <= last; c++) {
        rho_old = rho;
        rho = 1 / (2 * b.sigma - rho_old);
        neumann_ghosts(h, b, Y, false, check_done);
        double* out = (c == last) ? X : W;
        const double sgn = (c == last) ? -1.0 : 1.0;
        if (h->parity) {
            EpiChebStep<true> e{out, B, Z, rho, rho_old, 2 * b.sigma, 2 / b.delta, sgn};
            launch_stencil(h, KC_CHEB_STEP, b, Y, box, e, red, t, check_done);
        } else {
            EpiChebStep<false> e{out, B, Z, rho, rho_old, 2 * b.sigma, 2 / b.delta, sgn};
            launch_stencil(h, KC_CHEB_STEP, b, Y, box, e, red, t, check_done);
        }
        // swap(Z, Y); swap(W, Y)  (chebyshevIteration.hpp:114-115)
        double* tmp = Z; Z = Y; Y = tmp;
        tmp = W; W = Y; Y = tmp;
    }
}

// ------------------------------------------------------------------------------------------------
// Nested Krylov preconditioners (inputParam.hpp:29,31): BiCGSTAB<.., isMainLoop = false, communicationOFF, NoneSolver> and
// BaseCG<.., isMainLoop = false, communicationOFF, ChebyshevIteration> in the preconditioner slot.  Everything is local to the
// block: no face exchange, no allreduce, Neumann ghosts are plain mirrors.  The nested solve has its own device control block
// (b.ictl): the operator / axpy kernels of the main loop are reused with it as the active control block, so its scalars,
// its `done` flag and its residual history live on the device exactly like the outer ones.  The host looks at the nested
// history `lag` iterations late to stop launching; when the OUTER solve has converged the nested one starts `done`.
// Reference quirks kept: X is zeroed and B is divided by its own norm and multiplied back at the end, so the caller's
// vector changes in its last bits (BiCGSTAB.hpp:96,310-314); when the start residual is already below the tolerance the
// solve returns WITHOUT multiplying back (:118-122).
// ------------------------------------------------------------------------------------------------
static void zero_field(pps_handle* h, const Block& b, double* f);
static void copy_field(pps_handle* h, const Block& b, double* dst, const double* src);

struct InnerScope {
    pps_handle* h;
    Ctl* saved;
    bool saved_local;
    InnerScope(pps_handle* h_, Ctl* c) : h(h_), saved(h_->actl), saved_local(h_->local_reduce) {
        h->actl = c;
        h->local_reduce = true;
    }
    ~InnerScope() {
        h->actl = saved;
        h->local_reduce = saved_local;
    }
};

// X = 0; B /= ||B||_solver; r = B - A X, ||r||.  Returns false when the nested solve must not iterate (start residual below the
// tolerance, or the outer solve is done).     normalizeProblemToFieldBNorm<false,false> + computeErrorOperatorA<false,false>
static bool nested_start(pps_handle* h, Block& b, double* X, double* B, double* r, bool check_done) {
    const Box box = b.g.solver_box();
    const Tiling tp = make_tiling(h, b.g, box, false);
    const Tiling ts = make_tiling(h, b.g, box, true);
    {
        LaunchScope ls(h, KC_SETUP);
        inner_begin_kernel<<<1, 1, 0, h->stream>>>(b.ictl, check_done ? h->ctl : nullptr, h->precond_tolerance, h->precond_max_iter);
        ls.count(1);
    }
    zero_field(h, b, X);                                                          // BiCGSTAB.hpp:96 / baseCG.hpp:79
    launch_pointwise(h, KC_DOT, b, box, OpDot{B, nullptr}, make_red(h, 2, tp.ctas(), 0, OP_NORM_B), tp, true);
    {
        LaunchScope ls(h, KC_SETUP);
        // X is all zeros: 0 / norm = 0, only B needs the division
        inner_scale_kernel<<<148 * 8, 256, 0, h->stream>>>(nullptr, B, b.g.dims.total, b.ictl, check_done ? h->ctl : nullptr, 0);
        ls.count(1);
        check_launch("inner_scale");
    }
    neumann_ghosts(h, b, X, false, true);
    if (h->parity) launch_stencil(h, KC_RESIDUAL, b, X, box, EpiResidual<true>{r, B}, make_red(h, 1, ts.ctas(), 0, OP_RESIDUAL0), ts, true);
    else           launch_stencil(h, KC_RESIDUAL, b, X, box, EpiResidual<false>{r, B}, make_red(h, 1, ts.ctas(), 0, OP_RESIDUAL0), ts, true);
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return !(b.ihist_host[0] < h->precond_tolerance);
}

// launch nested iterations until the (host-mapped) nested history shows convergence `lag` iterations ago
template <class EnqueueIteration>
static void nested_iterations(pps_handle* h, const double* ihist_host, EnqueueIteration&& enqueue) {
    const int lag = std::max(1, h->lag);
    const int nev = static_cast<int>(h->inner_events.size());
    int done_at = h->precond_max_iter;
    for (int it = 0; it < h->precond_max_iter; ++it) {
        enqueue();
        PPS_CUDA_CHECK(cudaEventRecord(h->inner_events[it % nev], h->stream));
        if (it >= lag) {
            PPS_CUDA_CHECK(cudaEventSynchronize(h->inner_events[(it - lag) % nev]));
            if (ihist_host[it - lag + 1] < h->precond_tolerance) { done_at = it - lag + 1; break; }
        }
    }
    if (done_at == h->precond_max_iter) {
        // the host looks `lag` iterations late: a solve that converged within the last `lag` iterations has not been seen yet.
        // The device stopped at the right iteration (every kernel honours the nested `done`); read the true count for the report.
        PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
        for (int it = std::max(0, h->precond_max_iter - lag); it < h->precond_max_iter; ++it)
            if (ihist_host[it + 1] < h->precond_tolerance) { done_at = it + 1; break; }
    }
    h->precond_iters += done_at;
}

// ghosts(X) (plain mirrors; BiCGSTAB.hpp:300, baseCG.hpp:237-238 for order 1) and X, B *= norm  (:310-314 / :250-254)
static void nested_finish(pps_handle* h, Block& b, double* X, double* B, bool reset_x_ghosts, bool check_done) {
    {
        // these run although the nested solve is `done` by now; only a converged OUTER solve switches them off
        InnerScope outer(h, h->ctl);
        if (reset_x_ghosts) neumann_ghosts(h, b, X, false, check_done);
    }
    LaunchScope ls(h, KC_SETUP);
    inner_scale_kernel<<<148 * 8, 256, 0, h->stream>>>(X, B, b.g.dims.total, b.ictl, check_done ? h->ctl : nullptr, 1);
    ls.count(1);
    check_launch("inner_scale");
}

static void nested_bicgstab(pps_handle* h, Block& b, double* X, double* B, bool check_done) {
    InnerScope scope(h, b.ictl);
    const Box box = b.g.solver_box();
    const Tiling tp = make_tiling(h, b.g, box, false);
    const Tiling ts = make_tiling(h, b.g, box, true);
    if (!nested_start(h, b, X, B, b.ir, check_done)) return;
    copy_field(h, b, b.ip, b.ir);                                                 // BiCGSTAB.hpp:125-126
    copy_field(h, b, b.ir0, b.ir);
    const bool parity = h->parity;
    nested_iterations(h, b.ihist_host, [&]() {
        // Mp = p (NoneSolver) ; ghosts ; v = A Mp ; r0.v ; alpha                                      :133-164
        neumann_ghosts(h, b, b.ip, false, true);
        launch_stencil(h, KC_APPLY_DOT, b, b.ip, box, EpiStoreDot{b.iv, b.ir0}, make_red(h, 1, ts.ctas(), 0, OP_BICG_ALPHA), ts, true);
        RedCtx none = make_red(h, 0, 1, 0, OP_NONE);
        if (parity) launch_pointwise(h, KC_S_UPDATE, b, box, OpSUpdate<true>{b.ir, b.iv, 0}, none, tp, true);          // :168-178
        else        launch_pointwise(h, KC_S_UPDATE, b, box, OpSUpdate<false>{b.ir, b.iv, 0}, none, tp, true);
        // z = r (NoneSolver) ; ghosts ; t = A z ; r.t, t.t ; omega                                    :181-225
        neumann_ghosts(h, b, b.ir, false, true);
        launch_stencil(h, KC_APPLY_DOT2, b, b.ir, box, EpiStoreDot2Self{b.it}, make_red(h, 2, ts.ctas(), 0, OP_BICG_OMEGA), ts, true);
        RedCtx rho = make_red(h, 2, tp.ctas(), 0, OP_BICG_RHO);                                         // :227-259
        if (parity) launch_pointwise(h, KC_XR_UPDATE, b, box, OpXRUpdate<true>{X, b.ir, b.ip, b.ir, b.it, b.ir0, 0, 0}, rho, tp, true);
        else        launch_pointwise(h, KC_XR_UPDATE, b, box, OpXRUpdate<false>{X, b.ir, b.ip, b.ir, b.it, b.ir0, 0, 0}, rho, tp, true);
        if (parity) launch_pointwise(h, KC_P_UPDATE, b, box, OpPUpdate<true>{b.ip, b.ir, b.iv, 0, 0}, none, tp, true);   // :262-272
        else        launch_pointwise(h, KC_P_UPDATE, b, box, OpPUpdate<false>{b.ip, b.ir, b.iv, 0, 0}, none, tp, true);
    });
    nested_finish(h, b, X, B, true, check_done);
}

static void chebyshev_local(pps_handle* h, Block& b, double* X, double* B, bool check_done);

static void nested_cg_chebyshev(pps_handle* h, Block& b, double* X, double* B, bool check_done) {
    InnerScope scope(h, b.ictl);
    const Box box = b.g.solver_box();
    const Tiling tp = make_tiling(h, b.g, box, false);
    const Tiling ts = make_tiling(h, b.g, box, true);
    if (!nested_start(h, b, X, B, b.ir, check_done)) return;
    chebyshev_local(h, b, b.iz, b.ir, true);                                      // baseCG.hpp:109
    copy_field(h, b, b.ip, b.iz);                                                 // :111
    const bool parity = h->parity;
    const bool order1 = h->cfg.order_neumann == 1;
    nested_iterations(h, b.ihist_host, [&]() {
        if (order1) neumann_ghosts(h, b, b.ip, false, true);                      // :123-124
        launch_stencil(h, KC_CG_APPLY, b, b.ip, box, EpiCgApply{b.iv, b.ir, b.iz}, make_red(h, 2, ts.ctas(), 0, OP_CG_ALPHA), ts, true);   // :126-151
        RedCtx none2 = make_red(h, 2, tp.ctas(), 0, OP_NONE);
        if (parity) launch_pointwise(h, KC_CG_XR, b, box, OpCgXR<true>{X, b.ir, b.ip, b.iv, 0}, none2, tp, true);        // :154-165
        else        launch_pointwise(h, KC_CG_XR, b, box, OpCgXR<false>{X, b.ir, b.ip, b.iv, 0}, none2, tp, true);
        chebyshev_local(h, b, b.iz, b.ir, true);                                  // :168
        launch_pointwise(h, KC_DOT, b, box, OpDot{b.ir, b.iz}, make_red(h, 2, tp.ctas(), 0, OP_CG_BETA), tp, true);      // :171-195
        RedCtx none = make_red(h, 0, 1, 0, OP_NONE);
        if (parity) launch_pointwise(h, KC_CG_P, b, box, OpCgP<true>{b.ip, b.iz, 0}, none, tp, true);                   // :197-208
        else        launch_pointwise(h, KC_CG_P, b, box, OpCgP<false>{b.ip, b.iz, 0}, none, tp, true);
    });
    nested_finish(h, b, X, B, order1, check_done);
}

static void precondition_all(pps_handle* h, FieldSel selX, FieldSel selB, bool check_done);   // all blocks (defined below the Krylov drivers)

// X = M(B) for one block: the preconditioner slot of the main solvers
static void precondition(pps_handle* h, Block& b, double* X, double* B, bool check_done) {
    switch (h->cfg.precond) {
        case PPS_PRECOND_NONE: return;
        case PPS_PRECOND_CHEBYSHEV: chebyshev_local(h, b, X, B, check_done); return;
        case PPS_PRECOND_BICGSTAB_LOCAL: nested_bicgstab(h, b, X, B, check_done); return;
        case PPS_PRECOND_CG_CHEB_LOCAL: nested_cg_chebyshev(h, b, X, B, check_done); return;
        default: throw std::runtime_error("unknown preconditioner");
    }
}

// ------------------------------------------------------------------------------------------------
// shared solver pieces
// ------------------------------------------------------------------------------------------------
static void upload_ctl(pps_handle* h) {
    PPS_CUDA_CHECK(cudaMemcpyAsync(h->ctl, &h->ctl_host, sizeof(Ctl), cudaMemcpyHostToDevice, h->stream));
}
static void download_ctl(pps_handle* h) {
    PPS_CUDA_CHECK(cudaMemcpyAsync(&h->ctl_host, h->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

static void zero_field(pps_handle* h, const Block& b, double* f) {
    PPS_CUDA_CHECK(cudaMemsetAsync(f, 0, sizeof(double) * static_cast<size_t>(b.g.dims.total), h->stream));
}
static void copy_field(pps_handle* h, const Block& b, double* dst, const double* src) {
    PPS_CUDA_CHECK(cudaMemcpyAsync(dst, src, sizeof(double) * static_cast<size_t>(b.g.dims.total), cudaMemcpyDeviceToDevice, h->stream));
}

// r = b - A x, ||r||  (computeErrorOperatorA, iterativeSolverBase.hpp:236-280)
static void residual(pps_handle* h, int op) {
    halo_exchange(h, sel_x, false);
    const unsigned int total = total_ctas(h, true);
    unsigned int off = 0;
    for (auto& b : h->blocks) {
        neumann_ghosts(h, b, b.x, true, false);
        const Box box = b.g.solver_box();
        const Tiling t = make_tiling(h, b.g, box, true);
        RedCtx red = make_red(h, 1, total, off, op);
        if (h->parity) launch_stencil(h, KC_RESIDUAL, b, b.x, box, EpiResidual<true>{b.r, b.b}, red, t, false);
        else           launch_stencil(h, KC_RESIDUAL, b, b.x, box, EpiResidual<false>{b.r, b.b}, red, t, false);
        off += t.ctas();
    }
    finish_reduction(h, 1, op, true);
}

// normalizeProblemToFieldBNorm<true, ON> (iterativeSolverBase.hpp:171-234)
static void normalize_problem(pps_handle* h) {
    const unsigned int total = total_ctas(h, false);
    unsigned int off = 0;
    for (auto& b : h->blocks) {
        copy_field(h, b, b.t, b.b);
        adjust_b(h, b, b.t);
        const Box box = b.g.solver_box();
        const Tiling t = make_tiling(h, b.g, box, false);
        RedCtx red = make_red(h, 2, total, off, OP_NORM_B);
        launch_pointwise(h, KC_DOT, b, box, OpDot{b.t, nullptr}, red, t, false);
        off += t.ctas();
    }
    finish_reduction(h, 1, OP_NORM_B, true);
    LaunchScope ls(h, KC_SETUP);
    for (auto& b : h->blocks) {
        scale_kernel<<<148 * 8, 256, 0, h->stream>>>(b.x, b.b, b.g.dims.total, h->ctl, 0, 1);
        ls.count(1);
        zero_field(h, b, b.t);
    }
    check_launch("scale");
}

static void denormalize(pps_handle* h) {
    LaunchScope ls(h, KC_SETUP);
    for (auto& b : h->blocks) {
        scale_kernel<<<148 * 8, 256, 0, h->stream>>>(b.x, b.b, b.g.dims.total, h->ctl, 1, 1);
        ls.count(1);
    }
    check_launch("scale");
}

static void begin_solve(pps_handle* h) {
    h->launch_count = 0;
    h->iter_in_solve = 0;
    h->precond_iters = 0;
    for (auto& s : h->stats) { s.ms = 0; s.launches = 0; }
    h->event_next = 0;
    Ctl& c = h->ctl_host;
    c.rho0 = 1; c.alpha = 1; c.omega = 1; c.beta = 1; c.err = -1; c.rz = 1; c.norm_b = 1;
    c.tol = h->cfg.tolerance;
    for (double& s : c.sums) s = 0;
    c.iter = 0; c.done = 0; c.max_iter = h->cfg.max_iter; c.pad = 0;
    for (int q = 0; q < 4; q++) {
        double* hh = h->hist_host[q];
        std::fill(hh, hh + h->hist_len, std::numeric_limits<double>::infinity());
    }
    c.hist_err = h->hist_dev[0]; c.hist_alpha = h->hist_dev[1]; c.hist_omega = h->hist_dev[2]; c.hist_rho = h->hist_dev[3];
    upload_ctl(h);
    PPS_CUDA_CHECK(cudaMemsetAsync(h->counter, 0, sizeof(unsigned int), h->stream));
    for (auto& b : h->blocks) {
        // std::fill(..., 0) of every work array at the top of operator() (BiCGSTAB.hpp:60-66)
        zero_field(h, b, b.r); zero_field(h, b, b.r0); zero_field(h, b, b.p); zero_field(h, b, b.v); zero_field(h, b, b.t);
        if (b.mp != b.p) zero_field(h, b, b.mp);
        if (b.z != b.r) zero_field(h, b, b.z);
        if (b.p2) { zero_field(h, b, b.p2); zero_field(h, b, b.v2); zero_field(h, b, b.s); }
    }
}

// host loop: launch iterations, look at the residual history `lag` iterations late
template <class EnqueueIteration>
static void run_iterations(pps_handle* h, EnqueueIteration&& enqueue) {
    const int lag = std::max(1, h->lag);
    const int nev = static_cast<int>(h->iter_events.size());
    const double tol = h->cfg.tolerance;
    const bool graph = h->use_graph && !h->profiling;
    for (int it = 0; it < h->cfg.max_iter; ++it) {
        if (graph) {
            if (h->iter_graph == nullptr) {
                // capture ONE iteration (launch-bound small grids: ~160 dependent launches per preconditioned iteration)
                const long long before = h->launch_count;
                cudaGraph_t g = nullptr;
                PPS_CUDA_CHECK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
                try {
                    enqueue();
                } catch (...) {
                    cudaStreamEndCapture(h->stream, &g);
                    if (g) cudaGraphDestroy(g);
                    throw;
                }
                PPS_CUDA_CHECK(cudaStreamEndCapture(h->stream, &g));
                cudaError_t e = cudaGraphInstantiate(&h->iter_graph, g, 0);
                cudaGraphDestroy(g);
                PPS_CUDA_CHECK(e);
                h->iter_graph_launches = h->launch_count - before;
                h->launch_count = before;
            }
            PPS_CUDA_CHECK(cudaGraphLaunch(h->iter_graph, h->stream));
            h->launch_count += h->iter_graph_launches;
        } else {
            enqueue();
        }
        PPS_CUDA_CHECK(cudaEventRecord(h->iter_events[it % nev], h->stream));
        if (it >= lag) {
            PPS_CUDA_CHECK(cudaEventSynchronize(h->iter_events[(it - lag) % nev]));
            // every rank reads the same (allreduced) value at the same point -> same decision, NCCL calls stay matched
            if (h->hist_host[0][it - lag + 1] < tol) break;
        }
    }
}

static void end_solve(pps_handle* h, bool reset_x_ghosts) {
    // BiCGSTAB.hpp:293-321 / baseCG.hpp:231-259
    halo_exchange(h, sel_x, false);
    if (reset_x_ghosts)
        for (auto& b : h->blocks) neumann_ghosts(h, b, b.x, true, false);
    PPS_CUDA_CHECK(cudaEventRecord(h->ev_loop1, h->stream));
    residual(h, OP_RESIDUAL_FINAL);
    denormalize(h);
    download_ctl(h);
    h->iters = h->ctl_host.iter;
    h->err_iter = h->ctl_host.err;
    h->err_op = h->ctl_host.sums[4];
    h->norm_b = h->ctl_host.norm_b;
    // normFieldB_ = 1 (BiCGSTAB.hpp:315), then the last halo exchange of x (:317-321)
    h->ctl_host.norm_b = 1;
    h->ctl_host.done = 0;
    upload_ctl(h);
    halo_exchange(h, sel_x, false);
}

// ------------------------------------------------------------------------------------------------
// halo exchange + Neumann ghosts + fused operator + scalar reduction: the block that BiCGSTAB.hpp:135-164 /
// :182-225 / baseCG.hpp:118-151 repeat.  With one block per GPU (world > 1) the faces travel on the halo
// stream while the operator runs on the interior box; the boundary shell (one cell thick on every face that
// has a neighbour) is launched after the exchange has landed.  All launches feed one ticket reduction.
// ------------------------------------------------------------------------------------------------
// `tz`: thickness of the z shells.  One plane would do, but a CTA that computes one plane still streams three; whole
// z-chunks keep the boundary launches as efficient as the interior one.
static void split_box(const BlockGeom& g, const Box& box, int tz, Box& inner, std::vector<Box>& shell) {
    inner = box;
    const int nzb = box.k1 - box.k0;
    tz = std::max(1, std::min(tz, nzb / 3));
    if (g.hc[4]) { shell.push_back(Box{box.i0, box.i1, box.j0, box.j1, inner.k0, inner.k0 + tz}); inner.k0 += tz; }
    if (g.hc[5]) { shell.push_back(Box{box.i0, box.i1, box.j0, box.j1, inner.k1 - tz, inner.k1}); inner.k1 -= tz; }
    if (g.hc[2]) { shell.push_back(Box{box.i0, box.i1, inner.j0, inner.j0 + 1, inner.k0, inner.k1}); inner.j0 += 1; }
    if (g.hc[3]) { shell.push_back(Box{box.i0, box.i1, inner.j1 - 1, inner.j1, inner.k0, inner.k1}); inner.j1 -= 1; }
    if (g.hc[0]) { shell.push_back(Box{inner.i0, inner.i0 + 1, inner.j0, inner.j1, inner.k0, inner.k1}); inner.i0 += 1; }
    if (g.hc[1]) { shell.push_back(Box{inner.i1 - 1, inner.i1, inner.j0, inner.j1, inner.k0, inner.k1}); inner.i1 -= 1; }
}
static bool box_empty(const Box& b) { return b.i0 >= b.i1 || b.j0 >= b.j1 || b.k0 >= b.k1; }

// `xchg`: fields whose faces are exchanged first, `ghost`: fields whose Neumann ghosts are rewritten (both are the operand
// itself for the plain operator kernels, and the INPUTS of the fused operand for stencil_tma_pre_kernel).
// `launch(block, box, tiling, red)` enqueues the operator kernel on h->launch_stream for a sub-box of the block.
// `plain_stencil`: the launch goes through launch_stencil, i.e. the in-kernel-wait / peer-transport schedules may be used.
template <class Launch>
static void overlapped_operator(pps_handle* h, const FieldSet& xchg, const FieldSet& ghost, int nacc, int op, bool plain_stencil, int slot,
                                Launch launch) {
    bool any_comm = false;
    for (int f = 0; f < 6; f++) any_comm = any_comm || h->blocks[0].g.hc[f];
    const bool overlap = h->world > 1 && h->overlap && !h->debug_no_halo && any_comm && xchg.n > 0;
    if (!overlap) {
        halo_exchange(h, xchg, true);
        const unsigned int total = total_ctas(h, true);
        unsigned int off = 0;
        for (auto& b : h->blocks) {
            for (int q = 0; q < ghost.n; q++) neumann_ghosts(h, b, ghost.sel[q](b), false, true);
            const Box box = b.g.solver_box();
            const Tiling t = make_tiling(h, b.g, box, true);
            RedCtx red = make_red(h, nacc, total, off, op);
            launch(b, box, t, red);
            off += t.ctas();
        }
    } else {
        Block& b = h->blocks[0];
        PPS_CUDA_CHECK(cudaEventRecord(h->ev_field_ready, h->stream));
        PPS_CUDA_CHECK(cudaStreamWaitEvent(h->halo_stream, h->ev_field_ready, 0));
        bool use_p2p = h->p2p && slot >= 0;
        for (int q = 0; q < xchg.n; q++) use_p2p = use_p2p && p2p_index(h, xchg.sel[q](b)) >= 0;
        unsigned int p2p_epoch = 0;
        const int p2p_field = slot;
        if (use_p2p) p2p_epoch = halo_push_p2p(h, slot, xchg);
        else halo_exchange(h, xchg, true, /*on_halo_stream=*/true);
        const bool z_only = !(b.g.hc[0] || b.g.hc[1] || b.g.hc[2] || b.g.hc[3]);
        for (int q = 0; q < ghost.n; q++) neumann_ghosts(h, b, ghost.sel[q](b), false, true);
        const Tiling t_all = make_tiling(h, b.g, b.g.solver_box(), true);
        if (use_p2p && plain_stencil && h->stencil_impl == 1 && h->overlap == 3 && t_all.grid.z >= 3) {
            // EXPERIMENTAL (PPS_OVERLAP=3, needs PPS_HALO_P2P=1): ONE launch; the TMA producers of the first / last z-chunk wait
            // in-kernel for the neighbours' epoch flags.  Unlike PPS_OVERLAP=2 the transport is pure DMA (halo_push_p2p), so the
            // waiting CTAs cannot keep a communication kernel off the SMs.
            const Box box = b.g.solver_box();
            const Tiling t = t_all;
            h->wait_next = HaloWait{h->recv_epoch + 2 * p2p_field, p2p_epoch, b.g.hc[4] ? 0 : -1, b.g.hc[5] ? b.g.n[2] + 1 : -1,
                                    (b.g.hc[4] && t.grid.z > 1) ? 1 : 0, h->recv_epoch + 2 * p2p_field + 1, 1};
            RedCtx red = make_red(h, nacc, t.ctas(), 0, op);
            launch(b, box, t, red);
        } else if (!use_p2p && plain_stencil && z_only && h->stencil_impl == 1 && h->overlap == 2 && t_all.grid.z >= 3) {
            // EXPERIMENTAL (PPS_OVERLAP=2): ONE launch; the first / last z-chunk run last and their TMA producer waits in-kernel
            // for the faces.  Fastest when it works, but it needs the NCCL kernel to become resident while waiting CTAs hold
            // the SMs -- CUDA gives no such forward-progress guarantee (it deadlocked on 8 GPUs), hence not the default.
            h->halo_epoch++;
            publish_halo_epoch_kernel<<<1, 1, 0, h->halo_stream>>>(h->halo_flag, h->halo_epoch);
            const Box box = b.g.solver_box();
            const Tiling t = t_all;
            h->wait_next = HaloWait{h->halo_flag, h->halo_epoch, b.g.hc[4] ? 0 : -1, b.g.hc[5] ? b.g.n[2] + 1 : -1,
                                    (b.g.hc[4] && t.grid.z > 1) ? 1 : 0};
            RedCtx red = make_red(h, nacc, t.ctas(), 0, op);
            launch(b, box, t, red);
        } else {
            // interior box on the compute stream now; the boundary shell on its own stream as soon as the faces have
            // landed -- both launches share the GPU (no serialisation, no in-kernel waiting) and feed one ticket reduction
            if (!use_p2p) PPS_CUDA_CHECK(cudaEventRecord(h->ev_halo_done, h->halo_stream));
            PPS_CUDA_CHECK(cudaEventRecord(h->ev_pre, h->stream));   // field updated, Neumann ghosts written
            Box inner;
            std::vector<Box> shell;
            split_box(b.g, b.g.solver_box(), t_all.zchunk, inner, shell);
            std::vector<Box> boxes;
            if (!box_empty(inner)) boxes.push_back(inner);
            const size_t n_inner = boxes.size();
            for (auto& sb : shell)
                if (!box_empty(sb)) boxes.push_back(sb);
            std::vector<Tiling> tl;
            unsigned int total = 0;
            for (auto& bx : boxes) { tl.push_back(make_tiling(h, b.g, bx, true)); total += tl.back().ctas(); }
            unsigned int off = 0;
            for (size_t q = 0; q < boxes.size(); q++) {
                if (q == n_inner) {
                    PPS_CUDA_CHECK(cudaStreamWaitEvent(h->bnd_stream, h->ev_pre, 0));
                    if (use_p2p) {
                        // the faces were pushed into my guard planes by the neighbours: wait for their epoch flags
                        for (int up = 0; up < 2; up++)
                            if (b.g.hc[4 + up])
                                await_epoch_sys_kernel<<<1, 1, 0, h->bnd_stream>>>(h->recv_epoch + 2 * p2p_field + up, p2p_epoch);
                    } else {
                        PPS_CUDA_CHECK(cudaStreamWaitEvent(h->bnd_stream, h->ev_halo_done, 0));
                    }
                    h->launch_stream = h->bnd_stream;
                }
                RedCtx red = make_red(h, nacc, total, off, op);
                launch(b, boxes[q], tl[q], red);
                off += tl[q].ctas();
            }
            h->launch_stream = h->stream;
            PPS_CUDA_CHECK(cudaEventRecord(h->ev_bnd_done, h->bnd_stream));
            PPS_CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->ev_bnd_done, 0));
            if (!use_p2p) PPS_CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->ev_halo_done, 0));
        }
    }
    if (nacc > 0) finish_reduction(h, nacc, op, false);
}

// the plain operator kernels: exchange + ghosts on the operand itself
template <class MakeEpi>
static void fused_operator(pps_handle* h, int kc, FieldSel sel, bool ghosts, int nacc, int op, MakeEpi make_epi) {
    // exchange slot of the peer transport: the operand of the second operator of a BiCGSTAB iteration is z (or r when fused
    // schedules fall back to the plain kernel); everything else is "first"; fields the peers do not map go over NCCL
    const int slot = (sel == sel_z || sel == sel_r) ? 1 : 0;
    overlapped_operator(h, FieldSet(sel), ghosts ? FieldSet(sel) : FieldSet(), nacc, op, true, slot,
                        [&](Block& b, const Box& box, const Tiling& t, const RedCtx& red) {
                            launch_stencil(h, kc, b, sel(b), box, make_epi(b), red, t, true);
                        });
}

// ------------------------------------------------------------------------------------------------
// BiCGSTAB (BiCGSTAB.hpp:55-322), isMainLoop = true, communicationON = true
// ------------------------------------------------------------------------------------------------
static void bicgstab_iteration(pps_handle* h) {
    const bool parity = h->parity;
    // Mp = M(p); halo(Mp); ghosts(Mp); v = A Mp; sum r0.v; alpha             :133-164
    precondition_all(h, sel_mp, sel_p, true);
    fused_operator(h, KC_APPLY_DOT, sel_mp, true, 1, OP_BICG_ALPHA, [](Block& b) { return EpiStoreDot{b.v, b.r0}; });
    for (auto& b : h->blocks) {                                               // :168-178
        const Box box = b.g.solver_box();
        const Tiling t = make_tiling(h, b.g, box, false);
        RedCtx red = make_red(h, 0, 1, 0, OP_NONE);
        if (parity) launch_pointwise(h, KC_S_UPDATE, b, box, OpSUpdate<true>{b.r, b.v, 0}, red, t, true);
        else        launch_pointwise(h, KC_S_UPDATE, b, box, OpSUpdate<false>{b.r, b.v, 0}, red, t, true);
    }
    // z = M(r); halo(z); ghosts(z); t = A z; sum r.t, t.t; omega             :181-225
    precondition_all(h, sel_z, sel_r, true);
    if (h->blocks[0].z == h->blocks[0].r) {
        fused_operator(h, KC_APPLY_DOT2, sel_z, true, 2, OP_BICG_OMEGA, [](Block& b) { return EpiStoreDot2Self{b.t}; });
    } else {
        fused_operator(h, KC_APPLY_DOT2, sel_z, true, 2, OP_BICG_OMEGA, [](Block& b) { return EpiStoreDot2{b.t, b.r}; });
    }
    {
        const unsigned int total = total_ctas(h, false);
        unsigned int off = 0;
        for (auto& b : h->blocks) {                                           // :227-246
            const Box box = b.g.solver_box();
            const Tiling t = make_tiling(h, b.g, box, false);
            RedCtx red = make_red(h, 2, total, off, OP_BICG_RHO);
            if (parity) launch_pointwise(h, KC_XR_UPDATE, b, box, OpXRUpdate<true>{b.x, b.r, b.mp, b.z, b.t, b.r0, 0, 0}, red, t, true);
            else        launch_pointwise(h, KC_XR_UPDATE, b, box, OpXRUpdate<false>{b.x, b.r, b.mp, b.z, b.t, b.r0, 0, 0}, red, t, true);
            off += t.ctas();
        }
        finish_reduction(h, 2, OP_BICG_RHO, false);                           // :247-259
    }
    for (auto& b : h->blocks) {                                               // :262-272
        const Box box = b.g.solver_box();
        const Tiling t = make_tiling(h, b.g, box, false);
        RedCtx red = make_red(h, 0, 1, 0, OP_NONE);
        if (parity) launch_pointwise(h, KC_P_UPDATE, b, box, OpPUpdate<true>{b.p, b.r, b.v, 0, 0}, red, t, true);
        else        launch_pointwise(h, KC_P_UPDATE, b, box, OpPUpdate<false>{b.p, b.r, b.v, 0, 0}, red, t, true);
    }
}

__global__ void compare_bits_kernel(const double* __restrict__ a, const double* __restrict__ b, long long n, unsigned long long* out);

// PPS_FUSE_CHECK=1 (diagnostics, needs PPS_FUSE_P=0 so that p2 / v2 are free): recompute s and t with the split kernels
// from the same r, v, alpha and compare bit for bit; host-synchronous, prints the first mismatches to stderr
static void fused_s_check(pps_handle* h, Block& b, const Box& box, const Tiling& ts, const Tiling& tp) {
    static unsigned long long* cnt = nullptr;
    if (!cnt) PPS_CUDA_CHECK(cudaMalloc(&cnt, 4 * sizeof(unsigned long long)));
    const unsigned long long init[4] = {0, ~0ull, 0, ~0ull};
    PPS_CUDA_CHECK(cudaMemcpyAsync(cnt, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    copy_field(h, b, b.p2, b.r);
    RedCtx none = make_red(h, 0, 1, 0, OP_NONE);
    launch_pointwise(h, KC_S_UPDATE, b, box, OpSUpdate<false>{b.p2, b.v, 0}, none, tp, true);
    launch_stencil(h, KC_APPLY, b, b.p2, box, EpiStore{b.v2}, none, ts, true);
    compare_bits_kernel<<<148 * 8, 256, 0, h->stream>>>(b.s, b.p2, b.g.dims.total, cnt);
    compare_bits_kernel<<<148 * 8, 256, 0, h->stream>>>(b.t, b.v2, b.g.dims.total, cnt + 2);
    unsigned long long got[4];
    PPS_CUDA_CHECK(cudaMemcpyAsync(got, cnt, sizeof(got), cudaMemcpyDeviceToHost, h->stream));
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if ((got[0] || got[2]) && h->fuse_check_events++ < 12) {
        auto dec = [&](unsigned long long i, char* buf) {
            const long long k = i / b.g.dims.plane, j = (i % b.g.dims.plane) / b.g.dims.pitch, c = i % b.g.dims.pitch;
            std::snprintf(buf, 64, "(i=%lld j=%lld k=%lld)", c - kOff, j, k);
        };
        char a[64] = "-", c[64] = "-";
        if (got[0]) dec(got[1], a);
        if (got[2]) dec(got[3], c);
        std::fprintf(stderr, "[fuse_check] enqueued iteration %d: s differs in %llu cells, first %s; t differs in %llu cells, first %s\n",
                     h->iter_in_solve, got[0], a, got[2], c);
    }
}

// PPS_FUSE_FULL: the same iteration in 3 kernels / 17 vector passes.  p and v are double-buffered (other CTAs still read
// the previous p and v on their halo while this CTA stores the new ones) and s gets its own array for the same reason.
// The fused operand u = f(inputs) is recomputed on the halo from the inputs' halo, so what travels between blocks / GPUs and
// what is mirrored on Neumann faces are the INPUTS (f is pointwise: f(mirror) = mirror(f), f(neighbour's) = neighbour's u):
//   fused_p reads r (new from the x/r update), p (stored by the previous fused_p, interior only), v (exchanged by the previous fused_s)
//   fused_s reads r (exchanged by this iteration's fused_p), v (new)
// i.e. three face exchanges per iteration (r + p in one message, then v) instead of the split schedule's two.
static void bicgstab_iteration_fused(pps_handle* h) {
    const bool parity = h->parity;
    const bool first = h->iter_in_solve == 0;
    if (!first && h->fuse_p) {
        // p' = r + beta (p - omega v) ; v' = A p' ; sum r0.v' ; alpha            :262-272 of the previous pass + :142-164
        const FieldSet in = h->fuse_s ? FieldSet(sel_r, sel_p) : FieldSet(sel_r, sel_p, sel_v);
        // longer chunks for fused_p where the block is deep enough to keep an interior box of several chunks
        h->zchunk_hint = h->zchunk_fused_p > 0 ? h->zchunk_fused_p : (h->blocks[0].g.n[2] >= 256 ? 85 : 0);
        overlapped_operator(h, in, in, 1, OP_BICG_ALPHA, false, 0, [&](Block& b, const Box& box, const Tiling& t, const RedCtx& red) {
            if (parity) launch_tma_pre<8, 4, true>(h, KC_FUSED_P, b, box, PrePUpdate<true>{b.p2, b.r, b.p, b.v, 0, 0}, EpiStoreDot{b.v2, b.r0}, red, t, true);
            else if (h->fuse_stages_p == 3) launch_tma_pre<8, 3, false>(h, KC_FUSED_P, b, box, PrePUpdate<false>{b.p2, b.r, b.p, b.v, 0, 0}, EpiStoreDot{b.v2, b.r0}, red, t, true);
            else        launch_tma_pre<8, 4, false>(h, KC_FUSED_P, b, box, PrePUpdate<false>{b.p2, b.r, b.p, b.v, 0, 0}, EpiStoreDot{b.v2, b.r0}, red, t, true);
        });
        h->zchunk_hint = 0;
        for (auto& b : h->blocks) {
            std::swap(b.p, b.p2);
            std::swap(b.v, b.v2);
            b.mp = b.p;
        }
    } else {
        // first iteration (p0 = r0 is already in place, BiCGSTAB.hpp:125) or PPS_FUSE_P=0: p-update in place, plain operator
        if (!first) {
            for (auto& b : h->blocks) {
                const Box box = b.g.solver_box();
                const Tiling tp = make_tiling(h, b.g, box, false);
                RedCtx none = make_red(h, 0, 1, 0, OP_NONE);
                if (parity) launch_pointwise(h, KC_P_UPDATE, b, box, OpPUpdate<true>{b.p, b.r, b.v, 0, 0}, none, tp, true);
                else        launch_pointwise(h, KC_P_UPDATE, b, box, OpPUpdate<false>{b.p, b.r, b.v, 0, 0}, none, tp, true);
            }
        }
        fused_operator(h, KC_APPLY_DOT, sel_p, true, 1, OP_BICG_ALPHA, [](Block& b) { return EpiStoreDot{b.v, b.r0}; });
    }
    if (h->fuse_s) {
        {   // s = r - alpha v ; t = A s ; sum s.t, t.t ; omega                        :168-225
            const FieldSet in = (first || !h->fuse_p) ? FieldSet(sel_v, sel_r) : FieldSet(sel_v);
            h->by_hint = (!parity && h->fuse_by_s == 16) ? 16 : 0;
            overlapped_operator(h, in, in, 2, OP_BICG_OMEGA, false, 1, [&](Block& b, const Box& box, const Tiling& t, const RedCtx& red) {
                if (parity) launch_tma_pre<8, 6, true>(h, KC_FUSED_S, b, box, PreSUpdate<true>{b.s, b.r, b.v, 0}, EpiStoreDot2Self{b.t}, red, t, true);
                else if (h->by_hint == 16 && h->fuse_stages_s == 3) launch_tma_pre<16, 3, false>(h, KC_FUSED_S, b, box, PreSUpdate<false>{b.s, b.r, b.v, 0}, EpiStoreDot2Self{b.t}, red, t, true);
                else if (h->by_hint == 16) launch_tma_pre<16, 4, false>(h, KC_FUSED_S, b, box, PreSUpdate<false>{b.s, b.r, b.v, 0}, EpiStoreDot2Self{b.t}, red, t, true);
                else if (h->fuse_stages_s == 3) launch_tma_pre<8, 3, false>(h, KC_FUSED_S, b, box, PreSUpdate<false>{b.s, b.r, b.v, 0}, EpiStoreDot2Self{b.t}, red, t, true);
                else if (h->fuse_stages_s == 4) launch_tma_pre<8, 4, false>(h, KC_FUSED_S, b, box, PreSUpdate<false>{b.s, b.r, b.v, 0}, EpiStoreDot2Self{b.t}, red, t, true);
                else        launch_tma_pre<8, 6, false>(h, KC_FUSED_S, b, box, PreSUpdate<false>{b.s, b.r, b.v, 0}, EpiStoreDot2Self{b.t}, red, t, true);
            });
        }
        h->by_hint = 0;
        if (h->fuse_check && !h->fuse_p) {
            Block& b = h->blocks[0];
            const Box box = b.g.solver_box();
            fused_s_check(h, b, box, make_tiling(h, b.g, box, true), make_tiling(h, b.g, box, false));
        }
        // x += alpha p + omega s ; r = s - omega t ; sum r0.r, r.r ; beta, rho     :227-259
        const unsigned int total = total_ctas(h, false);
        unsigned int off = 0;
        for (auto& b : h->blocks) {
            const Box box = b.g.solver_box();
            const Tiling tp = make_tiling(h, b.g, box, false);
            RedCtx red = make_red(h, 2, total, off, OP_BICG_RHO);
            if (parity) launch_pointwise(h, KC_XR_UPDATE, b, box, OpXRUpdateS<true>{b.x, b.r, b.p, b.s, b.t, b.r0, 0, 0}, red, tp, true);
            else        launch_pointwise(h, KC_XR_UPDATE, b, box, OpXRUpdateS<false>{b.x, b.r, b.p, b.s, b.t, b.r0, 0, 0}, red, tp, true);
            off += tp.ctas();
        }
        finish_reduction(h, 2, OP_BICG_RHO, false);
    } else {
        // PPS_FUSE_S=0: the split kernels for this half (s lives in r)
        for (auto& b : h->blocks) {
            const Box box = b.g.solver_box();
            const Tiling tp = make_tiling(h, b.g, box, false);
            RedCtx none = make_red(h, 0, 1, 0, OP_NONE);
            if (parity) launch_pointwise(h, KC_S_UPDATE, b, box, OpSUpdate<true>{b.r, b.v, 0}, none, tp, true);
            else        launch_pointwise(h, KC_S_UPDATE, b, box, OpSUpdate<false>{b.r, b.v, 0}, none, tp, true);
        }
        fused_operator(h, KC_APPLY_DOT2, sel_r, true, 2, OP_BICG_OMEGA, [](Block& b) { return EpiStoreDot2Self{b.t}; });
        const unsigned int total = total_ctas(h, false);
        unsigned int off = 0;
        for (auto& b : h->blocks) {
            const Box box = b.g.solver_box();
            const Tiling tp = make_tiling(h, b.g, box, false);
            RedCtx red3 = make_red(h, 2, total, off, OP_BICG_RHO);
            if (parity) launch_pointwise(h, KC_XR_UPDATE, b, box, OpXRUpdate<true>{b.x, b.r, b.p, b.r, b.t, b.r0, 0, 0}, red3, tp, true);
            else        launch_pointwise(h, KC_XR_UPDATE, b, box, OpXRUpdate<false>{b.x, b.r, b.p, b.r, b.t, b.r0, 0, 0}, red3, tp, true);
            off += tp.ctas();
        }
        finish_reduction(h, 2, OP_BICG_RHO, false);
    }
    h->iter_in_solve++;
}

static void cg_iteration(pps_handle* h) {
    const bool parity = h->parity;
    // halo(p); ghost reset only for orderNeumanBcs == 1 (baseCG.hpp:123-124); Ap = A p; sum r.z, p.Ap; alpha     :118-151
    const bool ghosts = h->cfg.order_neumann == 1;
    if (h->blocks[0].z == h->blocks[0].r)
        fused_operator(h, KC_CG_APPLY, sel_p, ghosts, 2, OP_CG_ALPHA, [](Block& b) { return EpiCgApplySelf{b.v, b.r}; });
    else
        fused_operator(h, KC_CG_APPLY, sel_p, ghosts, 2, OP_CG_ALPHA, [](Block& b) { return EpiCgApply{b.v, b.r, b.z}; });
    const bool none = h->cfg.precond == PPS_PRECOND_NONE;
    {
        const unsigned int total = total_ctas(h, false);
        unsigned int off = 0;
        for (auto& b : h->blocks) {                                           // :154-165 (+ :171-182 when z is r)
            const Box box = b.g.solver_box();
            const Tiling t = make_tiling(h, b.g, box, false);
            RedCtx red = make_red(h, 2, total, off, none ? OP_CG_BETA : OP_NONE);
            if (parity) launch_pointwise(h, KC_CG_XR, b, box, OpCgXR<true>{b.x, b.r, b.p, b.v, 0}, red, t, true);
            else        launch_pointwise(h, KC_CG_XR, b, box, OpCgXR<false>{b.x, b.r, b.p, b.v, 0}, red, t, true);
            off += t.ctas();
        }
        if (none) finish_reduction(h, 2, OP_CG_BETA, false);
    }
    if (!none) {
        precondition_all(h, sel_z, sel_r, true);                              // :168
        const unsigned int total = total_ctas(h, false);
        unsigned int off = 0;
        for (auto& b : h->blocks) {                                           // :171-182
            const Box box = b.g.solver_box();
            const Tiling t = make_tiling(h, b.g, box, false);
            RedCtx red = make_red(h, 2, total, off, OP_CG_BETA);
            launch_pointwise(h, KC_DOT, b, box, OpDot{b.r, b.z}, red, t, true);
            off += t.ctas();
        }
        finish_reduction(h, 2, OP_CG_BETA, false);
    }
    for (auto& b : h->blocks) {                                               // :197-208
        const Box box = b.g.solver_box();
        const Tiling t = make_tiling(h, b.g, box, false);
        RedCtx red = make_red(h, 0, 1, 0, OP_NONE);
        if (parity) launch_pointwise(h, KC_CG_P, b, box, OpCgP<true>{b.p, b.z, 0}, red, t, true);
        else        launch_pointwise(h, KC_CG_P, b, box, OpCgP<false>{b.p, b.z, 0}, red, t, true);
    }
}

// ------------------------------------------------------------------------------------------------
// Chebyshev iteration as the MAIN solver (chebyshevIteration.hpp:48-140 with isMainLoop = true, communicationON = true):
// chebyshevMax sweeps on the BC-adjusted copy of b, a face exchange + ghost reset before every sweep, no normalisation
// (normFieldB_ stays 1), x written on the solver range only, then ||b - A x||.  As in precondition(), the two iterates the
// reference computes and discards are skipped and X = -W is folded into the last live sweep (bit-identical x).
// ------------------------------------------------------------------------------------------------
// selB: the right-hand side (its faces are exchanged and its ghosts rewritten, as the reference does to the caller's array),
// selX: where X = -y_last goes.  Main solver: B = the BC-adjusted copy of b (in t), X = x.  Preconditioner slot with
// communicationON (a GLOBAL polynomial preconditioner instead of block-Jacobi): B = p or r, X = Mp or z.
template <bool PAR>
static void chebyshev_global_sweeps(pps_handle* h, FieldSel selB, FieldSel selX) {
    const int m = h->cfg.cheb_max_iter;
    const int last = m - 2;   // X = -y_last; validate() guarantees last >= 1
    const double theta = h->theta, inv_theta = 1.0 / h->theta;
    double rho_old = 1 / h->sigma;
    double rho = 1 / (2 * h->sigma - rho_old);
    const double c1 = 2 * rho / h->delta;
    // halo(B); ghosts(B); Z = B/theta; Y = c1 (2 B + A B / theta)                            :69-91
    fused_operator(h, KC_CHEB_FIRST, selB, true, 0, OP_NONE, [=](Block& b) {
        return EpiChebFirst<PAR>{b.cz, last == 1 ? selX(b) : b.cy, theta, inv_theta, c1, last == 1 ? -1.0 : 1.0};
    });
    for (int c = 2; c <= last; c++) {                                                        // :94-116
        rho_old = rho;
        rho = 1 / (2 * h->sigma - rho_old);
        const bool fin = c == last;
        const double r1 = rho, r0 = rho_old, s2 = 2 * h->sigma, d2 = 2 / h->delta;
        fused_operator(h, KC_CHEB_STEP, sel_cy, true, 0, OP_NONE, [=](Block& b) {
            return EpiChebStep<PAR>{fin ? selX(b) : b.cw, selB(b), b.cz, r1, r0, s2, d2, fin ? -1.0 : 1.0};
        });
        for (auto& b : h->blocks) {   // swap(Z, Y); swap(W, Y)  (:114-115)
            double* tmp = b.cz; b.cz = b.cy; b.cy = tmp;
            tmp = b.cw; b.cw = b.cy; b.cy = tmp;
        }
    }
}
template <bool PAR>
static void chebyshev_main_sweeps(pps_handle* h) {
    chebyshev_global_sweeps<PAR>(h, sel_t, sel_x);
}

// ------------------------------------------------------------------------------------------------
// GLOBAL nested BiCGSTAB preconditioner: BiCGSTAB<.., isMainLoop = false, communicationON = true, NoneSolver> in the preconditioner
// slot (BiCGSTAB.hpp:55-322; the alpaka tree names it T_PreconditionerBiCGStabGlobal, solverPoissonMPI_alpaka/include/inputParam.hpp:33).
// The nested solve spans ALL blocks / GPUs: face exchanges of its Mp (= p), z (= r) and X, allreduced scalars -- i.e. the main
// loop's `fused_operator` machinery (exchange + ghosts + overlapped operator + ticket reduction + allreduce) on the nested work
// vectors, with ONE nested control block as the active one (block 0's `ictl`: scalars, `done` flag and history on the device; after
// the allreduce every rank holds the same values, so all ranks take the same host decisions and NCCL calls stay matched).
// Same quirks as the local one: X zeroed (:96), B divided by its GLOBAL norm and multiplied back (:97,310-314), early return
// without multiplying back (:118-122), plain-mirror Neumann ghosts.
// ------------------------------------------------------------------------------------------------
static double* sel_ip(Block& b) { return b.ip; }
static double* sel_ir(Block& b) { return b.ir; }

struct GlobalInnerScope {   // the nested control block becomes the active one; reductions stay global
    pps_handle* h;
    Ctl* saved;
    GlobalInnerScope(pps_handle* h_, Ctl* c) : h(h_), saved(h_->actl) { h->actl = c; }
    ~GlobalInnerScope() { h->actl = saved; }
};

static void nested_bicgstab_global(pps_handle* h, FieldSel selX, FieldSel selB, bool check_done) {
    Ctl* gctl = h->blocks[0].ictl;
    const double* ghist = h->blocks[0].ihist_host;
    const Ctl* outer = check_done ? h->ctl : nullptr;
    const bool parity = h->parity;
    GlobalInnerScope scope(h, gctl);
    {
        LaunchScope ls(h, KC_SETUP);
        inner_begin_kernel<<<1, 1, 0, h->stream>>>(gctl, outer, h->precond_tolerance, h->precond_max_iter);
        ls.count(1);
    }
    for (auto& b : h->blocks) zero_field(h, b, selX(b));                                                    // BiCGSTAB.hpp:96
    {   // normalizeProblemToFieldBNorm<false, true> (iterativeSolverBase.hpp:171-234): global norm over the solver ranges
        const unsigned int total = total_ctas(h, false);
        unsigned int off = 0;
        for (auto& b : h->blocks) {
            const Box box = b.g.solver_box();
            const Tiling t = make_tiling(h, b.g, box, false);
            launch_pointwise(h, KC_DOT, b, box, OpDot{selB(b), nullptr}, make_red(h, 2, total, off, OP_NORM_B), t, true);
            off += t.ctas();
        }
        finish_reduction(h, 1, OP_NORM_B, false);
        LaunchScope ls(h, KC_SETUP);
        for (auto& b : h->blocks) {   // X is all zeros: only B needs the division
            inner_scale_kernel<<<148 * 8, 256, 0, h->stream>>>(nullptr, selB(b), b.g.dims.total, gctl, outer, 0);
            ls.count(1);
        }
        check_launch("inner_scale");
    }
    {   // computeErrorOperatorA<false, true> (:236-280): halo(X), plain-mirror ghosts, r = B - A X, global ||r||
        halo_exchange(h, FieldSet(selX), true);
        const unsigned int total = total_ctas(h, true);
        unsigned int off = 0;
        for (auto& b : h->blocks) {
            neumann_ghosts(h, b, selX(b), false, true);
            const Box box = b.g.solver_box();
            const Tiling t = make_tiling(h, b.g, box, true);
            RedCtx red = make_red(h, 1, total, off, OP_RESIDUAL0);
            if (parity) launch_stencil(h, KC_RESIDUAL, b, selX(b), box, EpiResidual<true>{b.ir, selB(b)}, red, t, true);
            else        launch_stencil(h, KC_RESIDUAL, b, selX(b), box, EpiResidual<false>{b.ir, selB(b)}, red, t, true);
            off += t.ctas();
        }
        finish_reduction(h, 1, OP_RESIDUAL0, false);
    }
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (ghist[0] < h->precond_tolerance) return;                  // :118-122 (also: the OUTER solve is done, history entry -1)
    for (auto& b : h->blocks) { copy_field(h, b, b.ip, b.ir); copy_field(h, b, b.ir0, b.ir); }              // :125-126
    nested_iterations(h, ghist, [&]() {
        // Mp = p (NoneSolver); halo(Mp); ghosts(Mp); v = A Mp; sum r0.v; alpha                              :133-164
        fused_operator(h, KC_APPLY_DOT, sel_ip, true, 1, OP_BICG_ALPHA, [](Block& b) { return EpiStoreDot{b.iv, b.ir0}; });
        for (auto& b : h->blocks) {                                                                          // :168-178
            const Box box = b.g.solver_box();
            const Tiling t = make_tiling(h, b.g, box, false);
            RedCtx none = make_red(h, 0, 1, 0, OP_NONE);
            if (parity) launch_pointwise(h, KC_S_UPDATE, b, box, OpSUpdate<true>{b.ir, b.iv, 0}, none, t, true);
            else        launch_pointwise(h, KC_S_UPDATE, b, box, OpSUpdate<false>{b.ir, b.iv, 0}, none, t, true);
        }
        // z = r (NoneSolver); halo(z); ghosts(z); t = A z; sum r.t, t.t; omega                               :181-225
        fused_operator(h, KC_APPLY_DOT2, sel_ir, true, 2, OP_BICG_OMEGA, [](Block& b) { return EpiStoreDot2Self{b.it}; });
        {
            const unsigned int total = total_ctas(h, false);
            unsigned int off = 0;
            for (auto& b : h->blocks) {                                                                      // :227-246
                const Box box = b.g.solver_box();
                const Tiling t = make_tiling(h, b.g, box, false);
                RedCtx red = make_red(h, 2, total, off, OP_BICG_RHO);
                if (parity) launch_pointwise(h, KC_XR_UPDATE, b, box, OpXRUpdate<true>{selX(b), b.ir, b.ip, b.ir, b.it, b.ir0, 0, 0}, red, t, true);
                else        launch_pointwise(h, KC_XR_UPDATE, b, box, OpXRUpdate<false>{selX(b), b.ir, b.ip, b.ir, b.it, b.ir0, 0, 0}, red, t, true);
                off += t.ctas();
            }
            finish_reduction(h, 2, OP_BICG_RHO, false);                                                      // :247-259
        }
        for (auto& b : h->blocks) {                                                                          // :262-272
            const Box box = b.g.solver_box();
            const Tiling t = make_tiling(h, b.g, box, false);
            RedCtx none = make_red(h, 0, 1, 0, OP_NONE);
            if (parity) launch_pointwise(h, KC_P_UPDATE, b, box, OpPUpdate<true>{b.ip, b.ir, b.iv, 0, 0}, none, t, true);
            else        launch_pointwise(h, KC_P_UPDATE, b, box, OpPUpdate<false>{b.ip, b.ir, b.iv, 0, 0}, none, t, true);
        }
    });
    {
        // halo(X); ghosts(X) (plain mirrors)                                                                :293-300
        // these run although the nested solve is `done` by now; only a converged OUTER solve switches them off
        GlobalInnerScope outer_scope(h, h->ctl);
        halo_exchange(h, FieldSet(selX), check_done);
        for (auto& b : h->blocks) neumann_ghosts(h, b, selX(b), false, check_done);
    }
    {
        LaunchScope ls(h, KC_SETUP);                                                                         // :310-314
        for (auto& b : h->blocks) {
            inner_scale_kernel<<<148 * 8, 256, 0, h->stream>>>(selX(b), selB(b), b.g.dims.total, gctl, outer, 1);
            ls.count(1);
        }
        check_launch("inner_scale");
    }
    halo_exchange(h, FieldSet(selX), check_done);                                                            // :317-321
}

// X = M(B) on every block: the preconditioner slot of the main solvers
static void precondition_all(pps_handle* h, FieldSel selX, FieldSel selB, bool check_done) {
    if (h->cfg.precond == PPS_PRECOND_NONE) return;
    if (h->cfg.precond == PPS_PRECOND_BICGSTAB_LOCAL && h->precond_comm) { nested_bicgstab_global(h, selX, selB, check_done); return; }
    if (h->cfg.precond == PPS_PRECOND_CHEBYSHEV && h->precond_comm) {
        // ChebyshevIteration<.., isMainLoop = false, communicationON, ..> (chebyshevIteration.hpp:69-73,97-101)
        if (h->parity) chebyshev_global_sweeps<true>(h, selB, selX);
        else           chebyshev_global_sweeps<false>(h, selB, selX);
        return;
    }
    for (auto& b : h->blocks) precondition(h, b, selX(b), selB(b), check_done);
}

static void solve_chebyshev_main(pps_handle* h) {
    begin_solve(h);
    for (auto& b : h->blocks) { copy_field(h, b, b.t, b.b); adjust_b(h, b, b.t); }          // :61-67
    PPS_CUDA_CHECK(cudaEventRecord(h->ev_loop0, h->stream));
    if (h->parity) chebyshev_main_sweeps<true>(h);
    else           chebyshev_main_sweeps<false>(h);
    PPS_CUDA_CHECK(cudaEventRecord(h->ev_loop1, h->stream));
    residual(h, OP_RESIDUAL_FINAL);                                                         // :134-135, with the caller's b
    download_ctl(h);
    h->iters = h->cfg.cheb_max_iter;                                                        // :137
    h->err_op = h->ctl_host.sums[4];
    h->err_iter = h->err_op;                                                                // :136
    h->norm_b = 1;
    h->hist_host[0][0] = h->err_op;
}

static void solve(pps_handle* h) {
    if (h->operator_only) throw std::runtime_error("this handle was created with PPS_FLAG_OPERATOR_ONLY: no solver vectors");
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    const auto wall0 = std::chrono::high_resolution_clock::now();
    PPS_CUDA_CHECK(cudaEventRecord(h->ev_start, h->stream));
    if (h->cfg.solver == PPS_SOLVER_CHEBYSHEV) {
        solve_chebyshev_main(h);
        PPS_CUDA_CHECK(cudaEventRecord(h->ev_end, h->stream));
        PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (h->halo_stream) PPS_CUDA_CHECK(cudaStreamSynchronize(h->halo_stream));
        if (h->bnd_stream) PPS_CUDA_CHECK(cudaStreamSynchronize(h->bnd_stream));
        float cms = 0;
        PPS_CUDA_CHECK(cudaEventElapsedTime(&cms, h->ev_loop0, h->ev_loop1));
        h->loop_seconds = cms * 1e-3;
        PPS_CUDA_CHECK(cudaEventElapsedTime(&cms, h->ev_start, h->ev_end));
        h->solver_seconds = cms * 1e-3;
        (void)wall0;
        if (h->profiling) collect_stats(h);
        return;
    }
    begin_solve(h);
    const bool cg = h->cfg.solver == PPS_SOLVER_CG;
    // halo(x); ghosts(x) with normFieldB_ = 1; normalise; r = b - A x      BiCGSTAB.hpp:86-112 / baseCG.hpp:69-103
    halo_exchange(h, sel_x, false);
    for (auto& b : h->blocks) neumann_ghosts(h, b, b.x, true, false);
    normalize_problem(h);
    residual(h, OP_RESIDUAL0);
    download_ctl(h);
    h->norm_b = h->ctl_host.norm_b;
    h->err_op = h->ctl_host.err;
    if (h->ctl_host.err < h->cfg.tolerance) {
        // the reference returns here without de-normalising (BiCGSTAB.hpp:118-122)
        h->iters = 0;
        h->err_iter = -1;
        PPS_CUDA_CHECK(cudaEventRecord(h->ev_end, h->stream));
        PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
        h->loop_seconds = 0;
        h->solver_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - wall0).count();
        return;
    }
    if (cg) {
        precondition_all(h, sel_z, sel_r, false);                             // baseCG.hpp:109
        for (auto& b : h->blocks) copy_field(h, b, b.p, b.z);                 // :111
    } else {
        for (auto& b : h->blocks) { copy_field(h, b, b.p, b.r); copy_field(h, b, b.r0, b.r); }   // BiCGSTAB.hpp:125-126
    }
    PPS_CUDA_CHECK(cudaEventRecord(h->ev_loop0, h->stream));
    if (cg) run_iterations(h, [&]() { cg_iteration(h); });
    else if (h->fuse_full) run_iterations(h, [&]() { bicgstab_iteration_fused(h); });
    else    run_iterations(h, [&]() { bicgstab_iteration(h); });
    end_solve(h, /*reset_x_ghosts=*/!cg || h->cfg.order_neumann == 1);   // BiCGSTAB.hpp:300 always; baseCG.hpp:237-238 for order 1
    PPS_CUDA_CHECK(cudaEventRecord(h->ev_end, h->stream));
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (h->halo_stream) PPS_CUDA_CHECK(cudaStreamSynchronize(h->halo_stream));   // exchanges of iterations launched past convergence
    if (h->bnd_stream) PPS_CUDA_CHECK(cudaStreamSynchronize(h->bnd_stream));
    if (h->p2p) {
        // peer-memory path: nobody may re-use or free its arrays while a neighbour still pushes into them -- my pushes are
        // complete (halo stream synchronised above); this allreduce returns once every rank can say the same
        PPS_NCCL_CHECK(nccl().AllReduce(h->ctl->sums + 6, h->ctl->sums + 6, 1, ncclDouble, ncclSum, h->comm, h->stream));
        PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    }
    float ms = 0;
    PPS_CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev_loop0, h->ev_loop1));
    h->loop_seconds = ms * 1e-3;
    PPS_CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev_start, h->ev_end));
    h->solver_seconds = ms * 1e-3;
    if (h->profiling) collect_stats(h);
}

// ------------------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------------------
static void validate(const pps_config& c, int rank, int world) {
    if (c.abi_version != PPS_ABI_VERSION) throw std::runtime_error("pps_config.abi_version mismatch");
    if (c.dim < 1 || c.dim > 3) throw std::runtime_error("DIM must be 1, 2 or 3");
    const int nr = c.nranks[0] * c.nranks[1] * c.nranks[2];
    for (int d = 0; d < 3; d++) {
        if (c.nranks[d] < 1 || c.npglobal[d] < 1) throw std::runtime_error("bad npglobal / nranks");
        if (d >= c.dim) {
            // main.cpp:40-48 reads only DIM rank counts; the unused axes hold one point (blockGrid.hpp:166-167)
            if (c.nranks[d] != 1) throw std::runtime_error("nranks must be 1 along the axes >= DIM");
            continue;
        }
        if (c.guards[d] != 1) throw std::runtime_error("guards must be 1 (second-order centred stencil)");
        if (c.npglobal[d] / c.nranks[d] < 3) throw std::runtime_error("blocks need at least 3 points per axis");
        if (!(c.ds[d] > 0)) throw std::runtime_error("ds must be positive");
    }
    for (int f = 0; f < 6; f++)
        if (c.bcs_type[f] != 0 && c.bcs_type[f] != 1) throw std::runtime_error("bcs_type must be 0 (Dirichlet) or 1 (Neumann)");
    if (world != 1 && world != nr)
        // same message class as main.cpp:51-55
        throw std::runtime_error("configuration of ranks not coherent: world_size " + std::to_string(world) + " ranks " +
                                 std::to_string(c.nranks[0]) + " " + std::to_string(c.nranks[1]) + " " + std::to_string(c.nranks[2]));
    if (rank < 0 || rank >= world) throw std::runtime_error("rank out of range");
    if (c.order_neumann != 1 && c.order_neumann != 2) throw std::runtime_error("orderNeumanBcs must be 1 or 2");
    if (c.solver != PPS_SOLVER_BICGSTAB && c.solver != PPS_SOLVER_CG && c.solver != PPS_SOLVER_CHEBYSHEV) throw std::runtime_error("unknown solver");
    if (c.precond < PPS_PRECOND_NONE || c.precond > PPS_PRECOND_CG_CHEB_LOCAL) throw std::runtime_error("unknown preconditioner");
    if (c.precond_max_iter < 0 || c.precond_tolerance < 0) throw std::runtime_error("precond_max_iter / precond_tolerance must be >= 0");
    if (c.solver == PPS_SOLVER_CHEBYSHEV && c.precond != PPS_PRECOND_NONE)
        throw std::runtime_error("Chebyshev as main solver takes no preconditioner (its T_Preconditioner slot is unused, chebyshevIteration.hpp:143)");
    if ((c.precond == PPS_PRECOND_CHEBYSHEV || c.precond == PPS_PRECOND_CG_CHEB_LOCAL || c.solver == PPS_SOLVER_CHEBYSHEV) && c.cheb_max_iter < 3)
        throw std::runtime_error("chebyshevMax must be >= 3");
    if (c.max_iter < 0) throw std::runtime_error("max_iter must be >= 0");
    if (c.precond_communication != 0 && c.precond_communication != 1) throw std::runtime_error("precond_communication must be 0 or 1");
    if (c.precond_communication == 1 && c.precond != PPS_PRECOND_BICGSTAB_LOCAL &&
        (c.precond != PPS_PRECOND_CHEBYSHEV || c.cheb_precision != PPS_CHEB_FP64 || c.cheb_eigenvalues != PPS_CHEB_EIG_GLOBAL))
        throw std::runtime_error("precond_communication = 1 is implemented for the fp64 Chebyshev preconditioner with global eigenvalue bounds "
                                 "and for the nested BiCGSTAB (BiCGSTAB<.., false, communicationON, NoneSolver>)");
    if (c.cheb_precision == PPS_CHEB_FP32 && c.dim != 3 && (c.precond == PPS_PRECOND_CHEBYSHEV || c.precond == PPS_PRECOND_CG_CHEB_LOCAL))
        throw std::runtime_error("cheb_precision = PPS_CHEB_FP32 (T_data_chebyshev = float) is implemented for DIM = 3 only");
    if (c.precond_communication == 1 && c.precond == PPS_PRECOND_BICGSTAB_LOCAL && c.solver != PPS_SOLVER_BICGSTAB)
        throw std::runtime_error("the global nested BiCGSTAB preconditioner is implemented inside the BiCGSTAB main solver");
}

static void destroy(pps_handle* h);

static pps_handle* create(const pps_config& cfg, int rank, int world, const unsigned char* uid) {
    validate(cfg, rank, world);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw std::runtime_error(std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
    // any throw below (out of memory at large grids, NCCL bootstrap, peer mapping) releases what has been acquired so far
    std::unique_ptr<pps_handle, void (*)(pps_handle*)> h(new pps_handle(), &destroy);
    h->cfg = cfg;
    h->rank = rank;
    h->world = world;
    if (cfg.device >= 0) h->device = cfg.device;
    else PPS_CUDA_CHECK(cudaGetDevice(&h->device));
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    h->parity = cfg.arithmetic == PPS_ARITH_PARITY;
    h->by = env_int("PPS_TILE_ROWS", 8) == 4 ? 4 : 8;
    h->stencil_impl = env_int("PPS_STENCIL_TMA", 1) ? 1 : 0;
    h->by_tma = env_int("PPS_TMA_ROWS", 8) == 16 ? 16 : 8;
    h->zchunk_stencil = env_int("PPS_ZCHUNK_STENCIL", 0);
    h->zchunk_point = env_int("PPS_ZCHUNK_POINT", 0);
    h->tma_l2_promo = env_int("PPS_TMA_L2PROMO", 3);
    h->zchunk_fused_p = env_int("PPS_ZCHUNK_FUSED_P", 0);
    h->lag = env_int("PPS_LAG", 3);
    h->overlap = env_int("PPS_OVERLAP", 1);
    h->debug_no_halo = env_int("PPS_DEBUG_NO_HALO", 0);
    h->batch_ghosts = env_int("PPS_BATCH_GHOSTS", 1);   // all Neumann faces of a block in one launch (bitwise identical; validated on B200 in round 2)
    h->zchunk_cheb = env_int("PPS_ZCHUNK_CHEB", 0);
    // alpaka-only configuration surface: mixed-precision and local-eigenvalue Chebyshev preconditioner
    h->cheb_eig_local = env_int("PPS_CHEB_EIG_LOCAL", cfg.cheb_eigenvalues) == PPS_CHEB_EIG_LOCAL;
    h->cheb_f32 = env_int("PPS_CHEB_F32", cfg.cheb_precision) == PPS_CHEB_FP32;
    h->cheb_block = std::max(0, std::min(env_int("PPS_CHEB_BLOCK", cfg.cheb_block), kChebMaxLev));
    h->precond_comm = cfg.precond_communication != 0;
    // graphs: one block-set on one GPU, iteration-invariant launches only (no ping-pong schedule, no host-synchronising nested solves)
    // Default: on for launch-bound problems (<= 2^25 cells on one GPU; the shipped 128x128x256 + Chebyshev default issues ~80 dependent
    // launches per iteration: 0.097 s per solve with stream launches, 0.070 s with graph replay + batched ghosts, B200, round 2).
    const long long cells_total = static_cast<long long>(cfg.npglobal[0]) * cfg.npglobal[1] * cfg.npglobal[2];
    h->use_graph = env_int("PPS_GRAPH", cells_total <= (1LL << 25) ? 1 : 0) != 0 && world == 1 && cfg.precond != PPS_PRECOND_BICGSTAB_LOCAL &&
                   cfg.precond != PPS_PRECOND_CG_CHEB_LOCAL && cfg.solver != PPS_SOLVER_CHEBYSHEV;
    PPS_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->launch_stream = h->stream;
    for (int d = 0; d < 3; d++) {
        h->coef.ds[d] = cfg.ds[d];
        h->coef.ds2[d] = cfg.ds[d] * cfg.ds[d];
        h->coef.inv[d] = 1.0 / (cfg.ds[d] * cfg.ds[d]);
        if (d >= cfg.dim) {
            // DIM < 3 (matrixFreeOperatorA.hpp:24-32): no term along this axis.  Its guard planes are zero, so the 3-D kernels
            // add (0 - 2u + 0) / inf = -0 (PARITY) or (..) * 0 (FAST)
            h->coef.ds2[d] = std::numeric_limits<double>::infinity();
            h->coef.inv[d] = 0.0;
        }
    }
    const int nr = cfg.nranks[0] * cfg.nranks[1] * cfg.nranks[2];
    if (world == 1) for (int r = 0; r < nr; r++) { Block b; b.g = make_block(cfg, r); h->blocks.push_back(std::move(b)); }
    else { Block b; b.g = make_block(cfg, rank); h->blocks.push_back(std::move(b)); }
    const bool has_precond = cfg.precond != PPS_PRECOND_NONE;   // any preconditioner: Mp and z are vectors of their own
    const bool cheb_vectors = cfg.precond == PPS_PRECOND_CHEBYSHEV || cfg.precond == PPS_PRECOND_CG_CHEB_LOCAL;
    const bool nested = cfg.precond == PPS_PRECOND_BICGSTAB_LOCAL || cfg.precond == PPS_PRECOND_CG_CHEB_LOCAL;
    if (cfg.precond_max_iter > 0) h->precond_max_iter = cfg.precond_max_iter;
    if (cfg.precond_tolerance > 0) h->precond_tolerance = cfg.precond_tolerance;
    {
        // 17-pass schedule (PPS_FUSE_FULL; what PPS_FUSE_AUTO picks since round 2, when the ring-release hazard that made it
        // irreproducible was root-caused, stencil_tma.cuh) wherever every halo value of the fused operand can be recomputed
        // locally: one block, no Neumann face, no preconditioner, BiCGSTAB, TMA operator kernels.  PPS_FUSE=1 forces the split schedule.
        const int want = env_int("PPS_FUSE", cfg.fusion);
        bool neumann = false;
        for (int f = 0; f < 6; f++) neumann = neumann || cfg.bcs_type[f] == 1;
        (void)neumann;
        h->fuse_full = want != PPS_FUSE_SPLIT && cfg.dim == 3 && !has_precond && cfg.solver == PPS_SOLVER_BICGSTAB &&
                       h->stencil_impl == 1 && h->by_tma == 8;
        h->fuse_p = env_int("PPS_FUSE_P", 1) != 0;
        h->fuse_s = env_int("PPS_FUSE_S", 1) != 0;
        h->fuse_check = env_int("PPS_FUSE_CHECK", 0);
        h->fuse_by_s = env_int("PPS_FUSE_BY_S", 16);   // 64 x 16 tiles for fused_s (halo rows 18/16 instead of 10/8): 0.753 ms against 0.771 ms at 512^3
        h->fuse_stages_p = env_int("PPS_FUSE_STAGES_P", 3);   // 3 stages = 64 KB of ring = 3 CTAs per SM: 1.062 ms against 1.106 ms with 4 stages at 512^3
        h->fuse_stages_s = env_int("PPS_FUSE_STAGES_S", 4);   // 46 KB of ring = 4 CTAs per SM: 0.789 ms against 0.799 ms with 6 stages at 512^3
    }
    unsigned long long max_ctas = 0;
    h->operator_only = (cfg.flags & PPS_FLAG_OPERATOR_ONLY) != 0;
    if (h->operator_only) h->fuse_full = false;
    if (h->fuse_full) h->use_graph = false;
    for (auto& b : h->blocks) {
        const long long n = b.g.dims.total;
        if (h->operator_only) {
            b.p = dalloc(b, n, h->stream); b.v = dalloc(b, n, h->stream);
            if (!(cfg.flags & PPS_FLAG_NO_DOT_VECTOR)) b.r0 = dalloc(b, n, h->stream);
            b.mp = b.p;
            max_ctas += static_cast<unsigned long long>((b.g.n[0] + 63) / 64) * ((b.g.n[1] + 3) / 4) * b.g.n[2];
            continue;
        }
        b.x = dalloc(b, n, h->stream); b.b = dalloc(b, n, h->stream); b.r = dalloc(b, n, h->stream);
        b.r0 = dalloc(b, n, h->stream); b.p = dalloc(b, n, h->stream); b.v = dalloc(b, n, h->stream);
        b.t = dalloc(b, n, h->stream);
        if (has_precond) {
            b.mp = dalloc(b, n, h->stream); b.z = dalloc(b, n, h->stream);
            if (cheb_vectors) {
                b.cy = dalloc(b, n, h->stream); b.cz = dalloc(b, n, h->stream); b.cw = dalloc(b, n, h->stream);
                if (h->cheb_block > 0 || h->cheb_f32) b.c4 = dalloc(b, n, h->stream);
            }
            if (nested) {
                b.ip = dalloc(b, n, h->stream); b.ir = dalloc(b, n, h->stream); b.iv = dalloc(b, n, h->stream);
                if (cfg.precond == PPS_PRECOND_BICGSTAB_LOCAL) { b.ir0 = dalloc(b, n, h->stream); b.it = dalloc(b, n, h->stream); }
                else b.iz = dalloc(b, n, h->stream);
                // control block + host-mapped histories of the nested solver
                const size_t hl = static_cast<size_t>(h->precond_max_iter) + 2;
                PPS_CUDA_CHECK(cudaHostAlloc(&b.ihist_host, sizeof(double) * 4 * hl, cudaHostAllocMapped));
                double* hd = nullptr;
                PPS_CUDA_CHECK(cudaHostGetDevicePointer(&hd, b.ihist_host, 0));
                Ctl ic{};
                ic.norm_b = 1;
                ic.hist_err = hd; ic.hist_alpha = hd + hl; ic.hist_omega = hd + 2 * hl; ic.hist_rho = hd + 3 * hl;
                PPS_CUDA_CHECK(cudaMalloc(&b.ictl, sizeof(Ctl)));
                PPS_CUDA_CHECK(cudaMemcpy(b.ictl, &ic, sizeof(Ctl), cudaMemcpyHostToDevice));
            }
        } else {
            b.mp = b.p;
            b.z = b.r;
        }
        if (cfg.solver == PPS_SOLVER_CG && !has_precond) b.z = b.r;
        if (cfg.solver == PPS_SOLVER_CHEBYSHEV) { b.cy = dalloc(b, n, h->stream); b.cz = dalloc(b, n, h->stream); b.cw = dalloc(b, n, h->stream); }
        if (h->fuse_full) { b.p2 = dalloc(b, n, h->stream); b.v2 = dalloc(b, n, h->stream); b.s = dalloc(b, n, h->stream); }
        if (world > 1) {
            for (int f = 0; f < 4; f++) {
                if (!b.g.hc[f]) continue;
                const long long cnt = static_cast<long long>(b.g.n[f / 2 == 0 ? 1 : 0]) * b.g.n[2];
                b.sendbuf[f] = dalloc(b, cnt * kMaxExchangeFields, h->stream);
                b.recvbuf[f] = dalloc(b, cnt * kMaxExchangeFields, h->stream);
            }
        }
        // upper bound of CTAs any tiling of this block can produce (zchunk >= 1)
        max_ctas += static_cast<unsigned long long>((b.g.n[0] + 63) / 64) * ((b.g.n[1] + 3) / 4) * b.g.n[2];
    }
    // partial sums: enough for the finest tiling we ever launch (capped; make_red checks)
    h->partial_capacity = static_cast<long long>(std::min<unsigned long long>(max_ctas, 1ull << 22));
    PPS_CUDA_CHECK(cudaMalloc(&h->partials, sizeof(double) * kMaxAcc * static_cast<size_t>(h->partial_capacity)));
    PPS_CUDA_CHECK(cudaMalloc(&h->counter, sizeof(unsigned int)));
    PPS_CUDA_CHECK(cudaMemsetAsync(h->counter, 0, sizeof(unsigned int), h->stream));
    PPS_CUDA_CHECK(cudaMalloc(&h->ctl, sizeof(Ctl)));
    h->actl = h->ctl;
    h->hist_len = cfg.max_iter + 2;
    for (int q = 0; q < 4; q++) {
        PPS_CUDA_CHECK(cudaHostAlloc(&h->hist_host[q], sizeof(double) * h->hist_len, cudaHostAllocMapped));
        PPS_CUDA_CHECK(cudaHostGetDevicePointer(&h->hist_dev[q], h->hist_host[q], 0));
    }
    // chebyshevIteration.hpp:22-26 (global eigenvalues; delta < 0)
    const double* eg = h->blocks[0].g.eig_global;
    h->theta = (eg[0] * cfg.cheb_rescale_min + eg[1] * cfg.cheb_rescale_max) * 0.5 * (1.0 + cfg.cheb_epsilon);
    h->delta = (eg[0] * cfg.cheb_rescale_min - eg[1] * cfg.cheb_rescale_max) * 0.5;
    h->sigma = h->theta / h->delta;
    for (auto& b : h->blocks) {
        b.theta = h->theta; b.delta = h->delta;
        if (h->cheb_eig_local) {
            // alpaka tree, `local` (chebyshevIterationAlpaka.hpp:30-31,71-76): every rank uses the bounds of ITS block, not rescaled,
            // no epsilon (sign of delta as in the CPU tree; the alpaka sign is applied where its kernels are evaluated)
            b.theta = (b.g.eig_local[0] + b.g.eig_local[1]) * 0.5;
            b.delta = (b.g.eig_local[0] - b.g.eig_local[1]) * 0.5;
        }
        b.sigma = b.theta / b.delta;
    }
    h->iter_events.resize(std::max(8, h->lag + 2));
    for (auto& ev : h->iter_events) PPS_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (nested) {
        h->inner_events.resize(std::max(8, h->lag + 2));
        for (auto& ev : h->inner_events) PPS_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    PPS_CUDA_CHECK(cudaEventCreate(&h->ev_start));
    PPS_CUDA_CHECK(cudaEventCreate(&h->ev_loop0));
    PPS_CUDA_CHECK(cudaEventCreate(&h->ev_loop1));
    PPS_CUDA_CHECK(cudaEventCreate(&h->ev_end));
    if (world > 1) {
        if (uid == nullptr) throw std::runtime_error("world_size > 1 needs the NCCL unique id of rank 0");
        ncclUniqueId id;
        static_assert(sizeof(ncclUniqueId) <= PPS_UNIQUE_ID_BYTES, "unique id size");
        std::memcpy(&id, uid, sizeof(id));
        PPS_NCCL_CHECK(nccl().CommInitRank(&h->comm, world, id, rank));
        // a second communicator for the faces, so that exchanges on the halo stream never interleave with the
        // scalar allreduces on the compute stream
        PPS_NCCL_CHECK(nccl().CommSplit(h->comm, 0, rank, &h->comm_halo, nullptr));
        int lo = 0, hi = 0;
        PPS_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        PPS_CUDA_CHECK(cudaStreamCreateWithPriority(&h->halo_stream, cudaStreamNonBlocking, hi));
        PPS_CUDA_CHECK(cudaStreamCreateWithPriority(&h->bnd_stream, cudaStreamNonBlocking, hi));
        PPS_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_pre, cudaEventDisableTiming));
        PPS_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_bnd_done, cudaEventDisableTiming));
        PPS_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_field_ready, cudaEventDisableTiming));
        PPS_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_halo_done, cudaEventDisableTiming));
        PPS_CUDA_CHECK(cudaMalloc(&h->halo_flag, sizeof(unsigned int)));
        PPS_CUDA_CHECK(cudaMemsetAsync(h->halo_flag, 0, sizeof(unsigned int), h->stream));
    }
    // peer-memory face transport for z-slabs (copy engines + DMA-written epoch flags, no SM, no NCCL kernel): the default since
    // round 2 (2 GPUs, 1024x1024x256: 90 % of the exchange hidden against 76 % over NCCL); PPS_HALO_P2P=0 forces NCCL send/recv
    if (world > 1 && env_int("PPS_HALO_P2P", 1) && h->overlap) setup_p2p(h.get());
    // In-kernel allreduce over peer memory (PeerReduce): EXPERIMENTAL.  Parity-green on 2 GPUs (tests/test_gpu_multi.py), worth
    // +0.2 % there; on 8 GPUs it produced wrong sums (bench.py's parity gate against the reference's first 20 iterations failed with
    // 1.5e-3 at iteration 10; not root-caused) -- so it is refused beyond 2 ranks unless forced with PPS_ALLREDUCE_P2P=2 for debugging.
    const int ar_mode = env_int("PPS_ALLREDUCE_P2P", 0);
    if (world > 1 && world <= 256 && (ar_mode == 2 || (ar_mode == 1 && world <= 2))) setup_allreduce_p2p(h.get());
    else if (ar_mode == 1 && world > 2 && rank == 0)
        std::fprintf(stderr, "[pps] PPS_ALLREDUCE_P2P=1 ignored on %d ranks (validated on 2 only): NCCL allreduce\n", world);
    h->ctl_host = Ctl{};
    h->ctl_host.norm_b = 1;
    upload_ctl(h.get());
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return h.release();
}

static void destroy(pps_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->halo_stream) cudaStreamSynchronize(h->halo_stream);
    if (h->comm_halo) nccl().CommDestroy(h->comm_halo);
    if (h->comm) nccl().CommDestroy(h->comm);
    if (h->halo_stream) cudaStreamDestroy(h->halo_stream);
    if (h->bnd_stream) cudaStreamDestroy(h->bnd_stream);
    if (h->ev_pre) cudaEventDestroy(h->ev_pre);
    if (h->ev_bnd_done) cudaEventDestroy(h->ev_bnd_done);
    if (h->ev_field_ready) cudaEventDestroy(h->ev_field_ready);
    if (h->ev_halo_done) cudaEventDestroy(h->ev_halo_done);
    if (h->halo_flag) cudaFree(h->halo_flag);
    if (h->iter_graph) cudaGraphExecDestroy(h->iter_graph);
    for (void* q : h->ipc_opened) cudaIpcCloseMemHandle(q);
    if (h->recv_epoch) cudaFree(h->recv_epoch);
    if (h->epoch_ring) cudaFreeHost(h->epoch_ring);
    if (h->ar_mail) cudaFree(h->ar_mail);
    if (h->ar_peer_table) cudaFree(h->ar_peer_table);
    for (auto& b : h->blocks) {
        for (double* p : b.owned) cudaFree(p);
        if (b.ictl) cudaFree(b.ictl);
        if (b.check_buf) cudaFree(b.check_buf);
        if (b.ihist_host) cudaFreeHost(b.ihist_host);
    }
    for (auto e : h->inner_events)
        if (e) cudaEventDestroy(e);
    if (h->partials) cudaFree(h->partials);
    if (h->counter) cudaFree(h->counter);
    if (h->ctl) cudaFree(h->ctl);
    for (int q = 0; q < 4; q++)
        if (h->hist_host[q]) cudaFreeHost(h->hist_host[q]);
    for (auto e : h->event_pool) cudaEventDestroy(e);
    for (auto e : h->iter_events)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {h->ev_start, h->ev_loop0, h->ev_loop1, h->ev_end})
        if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
}

// host (reference layout) <-> device (pitched layout)
// (DIM < 3: the host array has no guards along the unused axes, blockGrid.hpp:172-182; on the device it is row j = 1 /
//  plane k = 1 of the 3-D layout, whose other rows / planes stay zero)
static long long host_origin(const Block& b) {
    return kOff + (b.g.dim < 2 ? b.g.dims.pitch : 0) + (b.g.dim < 3 ? b.g.dims.plane : 0);
}
static void upload_field(pps_handle* h, const Block& b, double* dev, const double* host) {
    const size_t w = sizeof(double) * (b.g.n[0] + 2);
    PPS_CUDA_CHECK(cudaMemcpy2DAsync(dev + host_origin(b), sizeof(double) * b.g.dims.pitch, host, w, w,
                                     static_cast<size_t>(b.g.ref_extent(1)) * b.g.ref_extent(2), cudaMemcpyHostToDevice, h->stream));
}
static void download_field(pps_handle* h, const Block& b, double* host, const double* dev) {
    const size_t w = sizeof(double) * (b.g.n[0] + 2);
    PPS_CUDA_CHECK(cudaMemcpy2DAsync(host, w, dev + host_origin(b), sizeof(double) * b.g.dims.pitch, w,
                                     static_cast<size_t>(b.g.ref_extent(1)) * b.g.ref_extent(2), cudaMemcpyDeviceToHost, h->stream));
}


// ------------------------------------------------------------------------------------------------
// diagnostics of the fused operator kernels (pps_debug_fused): repeat ONE fused launch on frozen inputs and compare
// its outputs bit for bit with the split kernels' result
// ------------------------------------------------------------------------------------------------
struct OpFillHash {
    static constexpr int NACC = 0;
    double* out;
    unsigned long long seed;
    __device__ __forceinline__ void begin(const Ctl*) {}
    __device__ __forceinline__ double val(long long i) const {
        unsigned long long z = static_cast<unsigned long long>(i) * 0x9E3779B97F4A7C15ull + seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return static_cast<double>(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    }
    __device__ __forceinline__ void operator()(long long idx, bool m0, bool m1, double*) const {
        st2(out + idx, make_double2(val(idx), val(idx + 1)), m0, m1);
    }
};

// per-CTA sum and max of |x - u| over the data range (rows are grid-strided, columns block-strided); max starts at -1 like the
// reference's (iterativeSolverBase.hpp:298)
__global__ void check_solution_kernel(const double* __restrict__ x, const double* __restrict__ u, Dims d, Box bx, double* __restrict__ psum,
                                      double* __restrict__ pmax) {
    const int nj = bx.j1 - bx.j0;
    const long long rows = static_cast<long long>(nj) * (bx.k1 - bx.k0);
    double s = 0.0, m = -1.0;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int j = bx.j0 + static_cast<int>(row % nj), k = bx.k0 + static_cast<int>(row / nj);
        const long long base = kOff + d.pitch * (j + static_cast<long long>(d.ny + 2) * k);
        for (int i = bx.i0 + threadIdx.x; i < bx.i1; i += blockDim.x) {
            const double e = fabs(x[base + i] - u[base + i]);
            s += e;
            m = fmax(m, e);
        }
    }
    __shared__ double ss[8], sm[8];
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(kFullMask, s, o);
        m = fmax(m, __shfl_down_sync(kFullMask, m, o));
    }
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sm[threadIdx.x >> 5] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { s += ss[w]; m = fmax(m, sm[w]); }
        psum[blockIdx.x] = s;
        pmax[blockIdx.x] = m;
    }
}

// out[0] += number of differing entries, out[1] = min index of a differing entry
__global__ void compare_bits_kernel(const double* __restrict__ a, const double* __restrict__ b, long long n, unsigned long long* out) {
    unsigned long long cnt = 0, first = ~0ull;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        if (__double_as_longlong(a[i]) != __double_as_longlong(b[i])) {
            cnt++;
            if (static_cast<unsigned long long>(i) < first) first = static_cast<unsigned long long>(i);
        }
    if (cnt) {
        atomicAdd(out, cnt);
        atomicMin(out + 1, first);
    }
}

static void debug_fused(pps_handle* h, int which, int variant, int reps, long long* out, int nout) {
    if (!h->fuse_full) throw std::runtime_error("pps_debug_fused needs a handle created with PPS_FUSE_FULL");
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    for (int q = 0; q < nout; q++) out[q] = 0;
    Block& b = h->blocks[0];
    const Box box = b.g.solver_box();
    const Tiling ts = make_tiling(h, b.g, box, true);
    const Tiling tp = make_tiling(h, b.g, box, false);
    const long long n = b.g.dims.total;
    RedCtx none = make_red(h, 0, 1, 0, OP_NONE);
    h->ctl_host.done = 0; h->ctl_host.alpha = 0.37; h->ctl_host.beta = 0.81; h->ctl_host.omega = 0.53;
    upload_ctl(h);
    PPS_CUDA_CHECK(cudaMemsetAsync(h->counter, 0, sizeof(unsigned int), h->stream));
    launch_pointwise(h, KC_SETUP, b, box, OpFillHash{b.r, 1}, none, tp, false);
    launch_pointwise(h, KC_SETUP, b, box, OpFillHash{b.v, 2}, none, tp, false);
    launch_pointwise(h, KC_SETUP, b, box, OpFillHash{b.p, 3}, none, tp, false);
    launch_pointwise(h, KC_SETUP, b, box, OpFillHash{b.r0, 4}, none, tp, false);
    // reference with the split kernels: operand into p2, A*operand into v2
    double *u_ref = b.p2, *au_ref = b.v2, *u_out = b.s, *au_out = b.t;
    zero_field(h, b, u_ref); zero_field(h, b, au_ref); zero_field(h, b, u_out); zero_field(h, b, au_out);
    const int nacc = which == 0 ? 2 : 1;
    if (which == 0) {
        copy_field(h, b, u_ref, b.r);
        launch_pointwise(h, KC_S_UPDATE, b, box, OpSUpdate<false>{u_ref, b.v, 0}, none, tp, false);
        launch_stencil(h, KC_APPLY_DOT2, b, u_ref, box, EpiStoreDot2Self{au_ref}, make_red(h, 2, ts.ctas(), 0, OP_NONE), ts, false);
    } else {
        copy_field(h, b, u_ref, b.p);
        launch_pointwise(h, KC_P_UPDATE, b, box, OpPUpdate<false>{u_ref, b.r, b.v, 0, 0}, none, tp, false);
        launch_stencil(h, KC_APPLY_DOT, b, u_ref, box, EpiStoreDot{au_ref, b.r0}, make_red(h, 1, ts.ctas(), 0, OP_NONE), ts, false);
    }
    download_ctl(h);
    double sums_ref[2] = {h->ctl_host.sums[0], h->ctl_host.sums[1]};
    unsigned long long* cnt = nullptr;
    PPS_CUDA_CHECK(cudaMalloc(&cnt, 4 * sizeof(unsigned long long)));
    int events = 0;
    const bool stop_on_bad = reps < 0;   // negative count: stop at the first bad launch and leave every array as it is (pps_debug_peek)
    if (reps < 0) reps = -reps;
    for (int rep = 0; rep < reps; rep++) {
        const unsigned long long init[4] = {0, ~0ull, 0, ~0ull};
        PPS_CUDA_CHECK(cudaMemcpyAsync(cnt, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
        RedCtx red = make_red(h, nacc, ts.ctas(), 0, OP_NONE);
        if (which == 0) {
            const PreSUpdate<false> pre{u_out, b.r, b.v, 0};
            const EpiStoreDot2Self epi{au_out};
            switch (variant) {
                case 3: launch_tma_pre<8, 3, false>(h, KC_FUSED_S, b, box, pre, epi, red, ts, true); break;
                case 4: launch_tma_pre<8, 4, false>(h, KC_FUSED_S, b, box, pre, epi, red, ts, true); break;
                case 6: launch_tma_pre<8, 6, false>(h, KC_FUSED_S, b, box, pre, epi, red, ts, true); break;
                case 16: launch_tma_pre<8, 6, false, 1>(h, KC_FUSED_S, b, box, pre, epi, red, ts, true); break;
                default: throw std::runtime_error("pps_debug_fused: variant must be 3, 4, 6 (ring stages) or 16 (6 stages + proxy fence)");
            }
        } else {
            const PrePUpdate<false> pre{u_out, b.r, b.p, b.v, 0, 0};
            const EpiStoreDot epi{au_out, b.r0};
            switch (variant) {
                case 3: launch_tma_pre<8, 3, false>(h, KC_FUSED_P, b, box, pre, epi, red, ts, true); break;
                case 4: launch_tma_pre<8, 4, false>(h, KC_FUSED_P, b, box, pre, epi, red, ts, true); break;
                default: throw std::runtime_error("pps_debug_fused: variant must be 3 or 4 for the p kernel");
            }
        }
        compare_bits_kernel<<<148 * 8, 256, 0, h->stream>>>(u_out, u_ref, n, cnt);
        compare_bits_kernel<<<148 * 8, 256, 0, h->stream>>>(au_out, au_ref, n, cnt + 2);
        unsigned long long got[4];
        PPS_CUDA_CHECK(cudaMemcpyAsync(got, cnt, sizeof(got), cudaMemcpyDeviceToHost, h->stream));
        download_ctl(h);   // synchronises
        const bool bad_sum = h->ctl_host.sums[0] != sums_ref[0] || (nacc > 1 && h->ctl_host.sums[1] != sums_ref[1]);
        out[0] += static_cast<long long>(got[0]);
        out[1] += static_cast<long long>(got[2]);
        out[2] += bad_sum ? 1 : 0;
        if (got[0] || got[2] || bad_sum) {
            out[3]++;
            if (events < 4 && 8 + 5 * events + 4 < nout) {
                long long* e = out + 8 + 5 * events++;
                e[0] = rep; e[1] = static_cast<long long>(got[0]); e[2] = got[0] ? static_cast<long long>(got[1]) : -1;
                e[3] = static_cast<long long>(got[2]); e[4] = got[2] ? static_cast<long long>(got[3]) : -1;
            }
            if (stop_on_bad) break;
        }
    }
    out[4] = b.g.dims.pitch; out[5] = b.g.dims.plane; out[6] = ts.zchunk; out[7] = ts.ctas();
    cudaFree(cnt);
}

}  // namespace pps

// ================================================================================================
// C ABI
// ================================================================================================
#define PPS_API_BEGIN try {
#define PPS_API_END                              \
    return 0;                                    \
    }                                            \
    catch (const std::exception& e) {            \
        pps::g_last_error = e.what();            \
        return 1;                                \
    }                                            \
    catch (...) {                                \
        pps::g_last_error = "unknown error";     \
        return 1;                                \
    }

extern "C" {

const char* pps_last_error(void) { return pps::g_last_error.c_str(); }
int pps_version(void) { return PPS_ABI_VERSION; }

void pps_default_config(pps_config* c) {
    std::memset(c, 0, sizeof(*c));
    c->abi_version = PPS_ABI_VERSION;
    c->dim = 3;
    const int np[3] = {128, 128, 256};
    const int bcs[6] = {0, 1, 0, 1, 0, 1};
    for (int d = 0; d < 3; d++) { c->npglobal[d] = np[d]; c->nranks[d] = 1; c->ds[d] = 0.1; c->origin[d] = 0; c->guards[d] = 1; }
    for (int f = 0; f < 6; f++) c->bcs_type[f] = bcs[f];
    c->solver = PPS_SOLVER_BICGSTAB;
    c->precond = PPS_PRECOND_CHEBYSHEV;
    c->tolerance = 1e2 * 1e-10;
    c->max_iter = 1700;
    c->cheb_max_iter = 11;
    c->cheb_epsilon = 1e-4;
    c->cheb_rescale_min = 500;
    c->cheb_rescale_max = 1 - 1e-4;
    c->order_neumann = 2;
    c->arithmetic = PPS_ARITH_FAST;
    c->fusion = PPS_FUSE_AUTO;
    c->device = -1;
}

int pps_get_unique_id(unsigned char id[PPS_UNIQUE_ID_BYTES]) {
    PPS_API_BEGIN
    ncclUniqueId nid;
    PPS_NCCL_CHECK(nccl().GetUniqueId(&nid));
    std::memset(id, 0, PPS_UNIQUE_ID_BYTES);
    std::memcpy(id, &nid, sizeof(nid));
    PPS_API_END
}

int pps_create(const pps_config* cfg, int rank, int world_size, const unsigned char* unique_id, pps_handle** out) {
    PPS_API_BEGIN
    if (!cfg || !out) throw std::runtime_error("null argument");
    *out = pps::create(*cfg, rank, world_size, unique_id);
    PPS_API_END
}

int pps_destroy(pps_handle* h) {
    PPS_API_BEGIN
    pps::destroy(h);
    PPS_API_END
}

int pps_num_local_blocks(const pps_handle* h) { return h ? static_cast<int>(h->blocks.size()) : 0; }

int pps_block_info_get(const pps_handle* h, int rank, pps_block_info* out) {
    PPS_API_BEGIN
    if (!h || !out) throw std::runtime_error("null argument");
    // geometry of ANY rank of the decomposition can be queried, hosted here or not
    const int nr = h->cfg.nranks[0] * h->cfg.nranks[1] * h->cfg.nranks[2];
    if (rank < 0 || rank >= nr) throw std::runtime_error("rank out of range");
    const BlockGeom g = make_block(h->cfg, rank);
    out->rank = rank;
    for (int d = 0; d < 3; d++) {
        out->global_location[d] = g.loc[d];
        out->nlocal_noguards[d] = g.n[d];
        out->nlocal_guards[d] = g.ref_extent(d);
    }
    for (int f = 0; f < 6; f++) {
        // reference numbering: an unused axis (>= DIM) has no guards, its single point is index 0 (blockGrid.hpp:193-204)
        const int shift = (f / 2 >= g.dim) ? 1 : 0;
        out->limits_data[f] = g.ld[f] - shift;
        out->limits_solver[f] = g.ls[f] - shift;
        out->has_boundary[f] = g.hb[f];
        out->has_communication[f] = g.hc[f];
    }
    out->ntot_guards = g.ref_total();
    PPS_API_END
}

int pps_eigenvalues(const pps_handle* h, int rank, double g[2], double l[2]) {
    PPS_API_BEGIN
    if (!h || !g || !l) throw std::runtime_error("null argument");
    if (rank < 0 || rank >= h->cfg.nranks[0] * h->cfg.nranks[1] * h->cfg.nranks[2]) throw std::runtime_error("rank out of range");
    const BlockGeom b = make_block(h->cfg, rank);
    g[0] = b.eig_global[0]; g[1] = b.eig_global[1];
    l[0] = b.eig_local[0]; l[1] = b.eig_local[1];
    PPS_API_END
}

int pps_set_fields(pps_handle* h, int rank, const double* x_host, const double* b_host) {
    PPS_API_BEGIN
    if (h->operator_only) throw std::runtime_error("pps_set_fields: handle is operator-only (PPS_FLAG_OPERATOR_ONLY)");
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    Block* b = find_block(h, rank);
    if (x_host) upload_field(h, *b, b->x, x_host);
    if (b_host) upload_field(h, *b, b->b, b_host);
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    PPS_API_END
}

int pps_set_neumann_face(pps_handle* h, int rank, int face, const double* dudn_host, size_t count) {
    PPS_API_BEGIN
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    Block* b = find_block(h, rank);
    if (face < 0 || face >= 2 * b->g.dim) throw std::runtime_error("face out of range");
    int u, v;
    b->g.tangential(face, u, v);
    const size_t want = static_cast<size_t>(b->g.n[u]) * b->g.n[v];
    if (count != want) throw std::runtime_error("pps_set_neumann_face: expected " + std::to_string(want) + " values");
    if (!b->dudn[face]) b->dudn[face] = dalloc(*b, static_cast<long long>(want), h->stream);
    PPS_CUDA_CHECK(cudaMemcpyAsync(b->dudn[face], dudn_host, sizeof(double) * want, cudaMemcpyHostToDevice, h->stream));
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    PPS_API_END
}

int pps_solve(pps_handle* h) {
    PPS_API_BEGIN
    pps::solve(h);
    PPS_API_END
}

int pps_save_fields(pps_handle* h) {
    PPS_API_BEGIN
    if (h->operator_only) throw std::runtime_error("pps_save_fields: handle is operator-only (PPS_FLAG_OPERATOR_ONLY)");
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    for (auto& b : h->blocks) {
        if (!b.x_saved) { b.x_saved = dalloc(b, b.g.dims.total, h->stream); b.b_saved = dalloc(b, b.g.dims.total, h->stream); }
        copy_field(h, b, b.x_saved, b.x);
        copy_field(h, b, b.b_saved, b.b);
    }
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    PPS_API_END
}

int pps_restore_fields(pps_handle* h) {
    PPS_API_BEGIN
    if (h->operator_only) throw std::runtime_error("pps_restore_fields: handle is operator-only (PPS_FLAG_OPERATOR_ONLY)");
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    for (auto& b : h->blocks) {
        if (!b.x_saved) throw std::runtime_error("pps_restore_fields without pps_save_fields");
        copy_field(h, b, b.x, b.x_saved);
        copy_field(h, b, b.b, b.b_saved);
    }
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    PPS_API_END
}

int pps_get_solution(pps_handle* h, int rank, double* x_host) {
    PPS_API_BEGIN
    if (h->operator_only) throw std::runtime_error("pps_get_solution: handle is operator-only (PPS_FLAG_OPERATOR_ONLY)");
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    Block* b = find_block(h, rank);
    download_field(h, *b, x_host, b->x);
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    PPS_API_END
}

int pps_get_rhs(pps_handle* h, int rank, double* b_host) {
    PPS_API_BEGIN
    if (h->operator_only) throw std::runtime_error("pps_get_rhs: handle is operator-only (PPS_FLAG_OPERATOR_ONLY)");
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    Block* b = find_block(h, rank);
    download_field(h, *b, b_host, b->b);
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    PPS_API_END
}

int pps_get_iterations(const pps_handle* h) { return h->iters; }
long long pps_get_preconditioner_iterations(const pps_handle* h) { return h->precond_iters; }
double pps_get_error_iteration(const pps_handle* h) { return h->err_iter; }
double pps_get_error_operator(const pps_handle* h) { return h->err_op; }
double pps_get_norm_b(const pps_handle* h) { return h->norm_b; }
double pps_get_solver_seconds(const pps_handle* h) { return h->solver_seconds; }
double pps_get_loop_seconds(const pps_handle* h) { return h->loop_seconds; }

int pps_get_history(const pps_handle* h, int which, double* out, int capacity) {
    PPS_API_BEGIN
    if (which < 0 || which > 3) throw std::runtime_error("history selector out of range");
    // Chebyshev as main solver keeps no history (chebyshevIteration.hpp:132-139): entry 0 is the final residual
    const int n = h->cfg.solver == PPS_SOLVER_CHEBYSHEV ? (which == 0 ? 1 : 0) : (which == 0 ? h->iters + 1 : h->iters);
    if (capacity < n) throw std::runtime_error("history buffer too small: need " + std::to_string(n));
    std::memcpy(out, h->hist_host[which], sizeof(double) * n);
    PPS_API_END
}

int pps_check_solution(pps_handle* h, int rank, const double* u_exact_host, double* sum_abs, double* max_abs) {
    PPS_API_BEGIN
    if (h->operator_only) throw std::runtime_error("pps_check_solution: handle is operator-only (PPS_FLAG_OPERATOR_ONLY)");
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    Block* b = find_block(h, rank);
    // checkSolutionLocalGlobal (iterativeSolverBase.hpp:300-317): sum and max of |x - u| over the data range, reporting only.
    // u goes up (t is a free work vector between solves), per-CTA partial sums / maxima come back and are combined in index order.
    constexpr int kBlocks = 148 * 8;
    // (the buffer is per BLOCK: with virtual ranks the rank-threads of the C++ driver call this concurrently on one handle, each for its own block)
    if (!b->check_buf) PPS_CUDA_CHECK(cudaMalloc(&b->check_buf, sizeof(double) * 2 * kBlocks));
    upload_field(h, *b, b->t, u_exact_host);
    pps::check_solution_kernel<<<kBlocks, 256, 0, h->stream>>>(b->x, b->t, b->g.dims, b->g.data_box(), b->check_buf, b->check_buf + kBlocks);
    check_launch("check_solution");
    std::vector<double> part(2 * kBlocks);
    PPS_CUDA_CHECK(cudaMemcpyAsync(part.data(), b->check_buf, sizeof(double) * 2 * kBlocks, cudaMemcpyDeviceToHost, h->stream));
    zero_field(h, *b, b->t);
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    double s = 0, m = -1;
    for (int q = 0; q < kBlocks; q++) {
        s += part[q];
        if (part[kBlocks + q] > m) m = part[kBlocks + q];
    }
    *sum_abs = s;
    *max_abs = m;
    PPS_API_END
}

int pps_apply_operator(pps_handle* h, int rank, const double* in_host, double* out_host) {
    PPS_API_BEGIN
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    Block* b = find_block(h, rank);
    upload_field(h, *b, b->p, in_host);
    zero_field(h, *b, b->v);
    const Box box = b->g.solver_box();
    const Tiling t = make_tiling(h, b->g, box, true);
    RedCtx red = make_red(h, 0, 1, 0, OP_NONE);
    launch_stencil(h, KC_APPLY, *b, b->p, box, EpiStore{b->v}, red, t, false);
    download_field(h, *b, out_host, b->v);
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    PPS_API_END
}

int pps_apply_preconditioner(pps_handle* h, int rank, const double* b_host, double* x_host) {
    PPS_API_BEGIN
    if (h->operator_only) throw std::runtime_error("pps_apply_preconditioner: handle is operator-only (PPS_FLAG_OPERATOR_ONLY)");
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    if (h->precond_comm && h->cfg.precond == PPS_PRECOND_BICGSTAB_LOCAL)
        throw std::runtime_error("pps_apply_preconditioner applies a block-local preconditioner; the global nested BiCGSTAB is collective");
    Block* b = find_block(h, rank);
    upload_field(h, *b, b->p, b_host);
    if (b->mp != b->p) zero_field(h, *b, b->mp);
    precondition(h, *b, b->mp, b->p, false);
    download_field(h, *b, x_host, b->mp);
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    PPS_API_END
}

int pps_bench_operator(pps_handle* h, int reps, int with_dot, double* avg_ms) {
    PPS_API_BEGIN
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    Block& b = h->blocks[0];
    const Box box = b.g.solver_box();
    const Tiling t = make_tiling(h, b.g, box, true);
    cudaEvent_t e0, e1;
    PPS_CUDA_CHECK(cudaEventCreate(&e0));
    PPS_CUDA_CHECK(cudaEventCreate(&e1));
    h->ctl_host.done = 0;
    upload_ctl(h);
    PPS_CUDA_CHECK(cudaMemsetAsync(h->counter, 0, sizeof(unsigned int), h->stream));
    const bool with_halo = (with_dot & 2) != 0;
    with_dot &= 1;
    if (with_dot && b.r0 == nullptr) throw std::runtime_error("pps_bench_operator: the handle has no dot-product vector (PPS_FLAG_NO_DOT_VECTOR)");
    auto once = [&]() {
        if (with_halo) halo_exchange(h, sel_p, false);
        if (with_dot) {
            RedCtx red = make_red(h, 1, t.ctas(), 0, OP_NONE);
            launch_stencil(h, KC_APPLY_DOT, b, b.p, box, EpiStoreDot{b.v, b.r0}, red, t, false);
        } else {
            RedCtx red = make_red(h, 0, 1, 0, OP_NONE);
            launch_stencil(h, KC_APPLY, b, b.p, box, EpiStore{b.v}, red, t, false);
        }
    };
    for (int i = 0; i < 3; i++) once();
    PPS_CUDA_CHECK(cudaEventRecord(e0, h->stream));
    for (int i = 0; i < reps; i++) once();
    PPS_CUDA_CHECK(cudaEventRecord(e1, h->stream));
    PPS_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    PPS_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *avg_ms = ms / std::max(1, reps);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    PPS_API_END
}

int pps_set_profiling(pps_handle* h, int enabled) {
    // 0 off, 1 every kernel class, 2 + c only kernel class c (two event records per launch of that class)
    h->profiling = enabled != 0;
    h->profile_only = enabled >= 2 ? enabled - 2 : -1;
    return 0;
}

int pps_get_kernel_stats(const pps_handle* h, int which, double* avg_ms, long long* launches, const char** name) {
    PPS_API_BEGIN
    if (which < 0 || which >= KC_COUNT) throw std::runtime_error("kernel class out of range");
    const KernelStat& s = h->stats[which];
    if (launches) *launches = s.launches;
    if (avg_ms) *avg_ms = s.launches ? s.ms / static_cast<double>(s.launches) : 0.0;
    if (name) *name = kKernelNames[which];
    PPS_API_END
}

long long pps_get_launch_count(const pps_handle* h) { return h->launch_count; }

int pps_set_max_iterations(pps_handle* h, int max_iter) {
    PPS_API_BEGIN
    if (max_iter < 0 || max_iter + 2 > h->hist_len) throw std::runtime_error("max_iter exceeds the value the handle was created with");
    h->cfg.max_iter = max_iter;
    PPS_API_END
}

int pps_allgather(pps_handle* h, const double* in_host, int n, double* out_host) {
    PPS_API_BEGIN
    if (h->world == 1) {
        std::memcpy(out_host, in_host, sizeof(double) * n);
    } else {
        PPS_CUDA_CHECK(cudaSetDevice(h->device));
        if (static_cast<long long>(n) * h->world > h->partial_capacity) throw std::runtime_error("pps_allgather: too many values");
        double* dev = h->partials;   // scratch: idle between solves
        PPS_CUDA_CHECK(cudaMemcpyAsync(dev + static_cast<size_t>(h->rank) * n, in_host, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
        PPS_NCCL_CHECK(nccl().AllGather(dev + static_cast<size_t>(h->rank) * n, dev, n, ncclDouble, h->comm, h->stream));
        PPS_CUDA_CHECK(cudaMemcpyAsync(out_host, dev, sizeof(double) * n * h->world, cudaMemcpyDeviceToHost, h->stream));
        PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    }
    PPS_API_END
}

int pps_debug_fused(pps_handle* h, int which, int variant, int reps, long long* out, int nout) {
    PPS_API_BEGIN
    if (nout < 8) throw std::runtime_error("pps_debug_fused: out needs at least 8 entries");
    pps::debug_fused(h, which, variant, reps, out, nout);
    PPS_API_END
}

int pps_debug_peek(pps_handle* h, int array, long long offset, int n, double* out) {
    PPS_API_BEGIN
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    pps::Block& b = h->blocks[0];
    const double* src[] = {b.s, b.t, b.p2, b.v2, b.r, b.v, b.p, b.r0, b.x};
    if (array < 0 || array > 8 || src[array] == nullptr) throw std::runtime_error("pps_debug_peek: no such array");
    if (offset < 0 || offset + n > b.g.dims.total) throw std::runtime_error("pps_debug_peek: out of range");
    PPS_CUDA_CHECK(cudaMemcpy(out, src[array] + offset, sizeof(double) * n, cudaMemcpyDeviceToHost));
    PPS_API_END
}

int pps_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int pps_synchronize(pps_handle* h) {
    PPS_API_BEGIN
    PPS_CUDA_CHECK(cudaSetDevice(h->device));
    PPS_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    PPS_API_END
}

}  // extern "C"

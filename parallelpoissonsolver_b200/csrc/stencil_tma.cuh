// stencil_tma.cuh -- operator family with TMA-staged planes (sm_100a: cp.async.bulk.tensor + mbarrier).
//
// Why: the plain-load kernel (kernels.cuh) keeps only one 16-B load per thread in flight towards DRAM, and
// measured ~4.0 TB/s on y = A x at 512^3.  Here a dedicated producer warp streams the halo'd xy-tile of
// every z-plane of the CTA's chunk into a shared-memory ring with 3-D tensor-map bulk copies, STAGES planes
// ahead of the consumers, so the bytes in flight per SM are set by the ring depth and not by registers.
//
//   tile            64 x BY cells  (one consumer warp per row, one double2 per lane)
//   box             72 x (BY+2) x 1 doubles, starting 4 columns left of the tile: a 32-B aligned start, the
//                   same 18 sectors per row that the 66 wide halo'd row touches anyway; TMA zero-fills rows /
//                   columns outside the array, those cells are masked
//   ring            STAGES boxes, full[s] (TMA complete_tx) / empty[s] (one arrive per consumer warp) mbarriers
//   consumers       keep k-1 / k / k+1 of their own column in registers (z-march), read y+-1 from the ring,
//                   x+-1 by warp shuffle (edge lanes read the ring), fuse the epilogue functor, store with
//                   128-B aligned double2 stores
#pragma once

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "kernels.cuh"

namespace pps {

constexpr int kTmaBoxX = 72;
constexpr int kTmaLead = 4;   // box starts kTmaLead columns left of the tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// spin until the halo stream has published `epoch` (monotonic counter), then order the following bulk-tensor read
// (async proxy) after the acquire
__device__ __forceinline__ void wait_halo_flag(const unsigned int* flag, unsigned int epoch, int sys_scope = 0) {
    unsigned int v;
    do {
        if (sys_scope) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        else           asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (static_cast<int>(v - epoch) >= 0) break;
        __nanosleep(200);
    } while (true);
    asm volatile("fence.proxy.async;" ::: "memory");
}

__global__ void publish_halo_epoch_kernel(unsigned int* flag, unsigned int epoch) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}

template <int BY, int STAGES, int NAUX>
struct TmaSmem {
    static constexpr int kMainElems = kTmaBoxX * (BY + 2);
    static constexpr int kAuxElems = 64 * BY;                      // aux streams need no halo
    static constexpr int kStageElems = kMainElems + NAUX * kAuxElems;
    static constexpr int kMainBytes = kMainElems * 8;
    static constexpr int kStageBytes = kStageElems * 8;
    static constexpr int kBytes = STAGES * kStageBytes + 2 * STAGES * 8;
};

template <int BY, int STAGES, bool PARITY, class Epi>
__global__ void __launch_bounds__(32 * (BY + 1)) stencil_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                   const __grid_constant__ CUtensorMap tmap_a0,
                                                                   const __grid_constant__ CUtensorMap tmap_a1, Dims d, Box rg,
                                                                   Coef cf, int zchunk, TileOrigin org, HaloWait hw, Epi epi,
                                                                   RedCtx red, const Ctl* ctl) {
    if (ctl != nullptr && ctl->done) return;
    constexpr int NACC = Epi::NACC;
    constexpr int NAUX = Epi::NAUX;
    constexpr int SX = kTmaBoxX;
    using SM = TmaSmem<BY, STAGES, NAUX>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * SM::kStageBytes);
    uint64_t* empty = full + STAGES;

    const int lane = threadIdx.x, warp = threadIdx.y;
    const int col0 = kFirstDataCol + 64 * (blockIdx.x + org.bx0);
    const int y0 = 1 + BY * (blockIdx.y + org.by0);
    const int zidx = hw.shift ? static_cast<int>((blockIdx.z + hw.shift) % gridDim.z) : static_cast<int>(blockIdx.z);
    const int kchunk = org.kfirst + zidx * zchunk;
    const int kb = max(rg.k0, kchunk), ke = min(rg.k1, kchunk + zchunk);
    const int nplanes = ke - kb + 2;   // planes kb-1 .. ke

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], BY);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double acc[NACC > 0 ? NACC : 1];
#pragma unroll
    for (int a = 0; a < (NACC > 0 ? NACC : 1); a++) acc[a] = 0.0;

    if (kb < ke) {
        if (warp == BY) {
            // ---------------- producer: one lane streams the planes of this chunk through the ring
            if (lane == 0) {
                for (int p = 0; p < nplanes; p++) {
                    const int s = p % STAGES;
                    if (p >= STAGES) mbar_wait(&empty[s], ((p / STAGES) - 1) & 1);
                    double* dst = ring + s * SM::kStageElems;
                    // aux streams are only read where the operator is evaluated: planes kb .. ke-1
                    const bool with_aux = NAUX > 0 && p >= 1 && p <= nplanes - 2;
                    const int kk = kb - 1 + p;
                    if (hw.flag != nullptr && (kk == hw.lo_plane || kk == hw.hi_plane))
                        wait_halo_flag((kk == hw.hi_plane && hw.flag_hi != nullptr) ? hw.flag_hi : hw.flag, hw.epoch, hw.sys_scope);
                    mbar_arrive_expect_tx(&full[s], with_aux ? SM::kStageBytes : SM::kMainBytes);
                    tma_load_3d(dst, &tmap, &full[s], col0 - kTmaLead, y0 - 1, kk);
                    if (with_aux) {
                        if (NAUX > 0) tma_load_3d(dst + SM::kMainElems, &tmap_a0, &full[s], col0, y0, kb - 1 + p);
                        if (NAUX > 1) tma_load_3d(dst + SM::kMainElems + SM::kAuxElems, &tmap_a1, &full[s], col0, y0, kb - 1 + p);
                    }
                }
            }
        } else {
            // ---------------- consumers
            const int col = col0 + 2 * lane;
            const int j = y0 + warp;
            const int i0 = col - kOff;
            const bool jr = j >= rg.j0 && j < rg.j1;
            const bool m0 = jr && i0 >= rg.i0 && i0 < rg.i1;
            const bool m1 = jr && i0 + 1 >= rg.i0 && i0 + 1 < rg.i1;
            const bool warp_any = __any_sync(kFullMask, m0 || m1);
            const long long colc = col < d.pitch ? col : d.pitch - 2;
            const long long rowoff = colc + d.pitch * min(j, d.ny);
            const int so = (warp + 1) * SX + kTmaLead + 2 * lane;
            const int sa = SM::kMainElems + warp * 64 + 2 * lane;

            // Stage 0 (plane kb-1) is released at the END of the first loop iteration, together with stage 1: a ring stage may
            // only be handed back to the producer after the values read from it have reached registers.  An LDS.128 of a warp
            // is four 8-lane wavefronts, and a wavefront can still be queued when a later mbarrier.arrive of the same warp has
            // already been performed (nothing orders the two unless an instruction CONSUMES the loaded registers first).
            // Releasing stage 0 right after the load let the refill (plane kb-1+STAGES) overtake such a wavefront about once
            // per 1e6 CTAs at 512^3 (root cause of the round-1 fused-schedule anomaly; tools/fused_debug3.py found the
            // refilled plane's values in the k-1 neighbour).  In the loop every value read from stage s feeds the stores
            // that precede the release of s.
            mbar_wait(&full[0], 0);
            double2 cm = *reinterpret_cast<const double2*>(ring + so);
            mbar_wait(&full[1 % STAGES], (1 / STAGES) & 1);
            double2 cc = *reinterpret_cast<const double2*>(ring + (1 % STAGES) * SM::kStageElems + so);
            for (int k = kb; k < ke; ++k) {
                const int p = k - kb + 1;
                const int s = p % STAGES, sn = (p + 1) % STAGES;
                mbar_wait(&full[sn], ((p + 1) / STAGES) & 1);
                const double* stg = ring + s * SM::kStageElems;
                const double* cur = stg + so;
                const double2 cp = *reinterpret_cast<const double2*>(ring + sn * SM::kStageElems + so);
                if (warp_any) {
                    const double2 ym = *reinterpret_cast<const double2*>(cur - SX);
                    const double2 yp = *reinterpret_cast<const double2*>(cur + SX);
                    double2 ax[NAUX > 0 ? NAUX : 1];
#pragma unroll
                    for (int a = 0; a < NAUX; a++) ax[a] = *reinterpret_cast<const double2*>(stg + sa + a * SM::kAuxElems);
                    double xl = __shfl_up_sync(kFullMask, cc.y, 1);
                    double xr = __shfl_down_sync(kFullMask, cc.x, 1);
                    if (lane == 0) xl = cur[-1];
                    if (lane == 31) xr = cur[2];
                    double2 au;
                    au.x = laplacian<PARITY>(cf, xl, cc.x, cc.y, ym.x, yp.x, cm.x, cp.x);
                    au.y = laplacian<PARITY>(cf, cc.x, cc.y, xr, ym.y, yp.y, cm.y, cp.y);
                    epi(rowoff + k * d.plane, au, cc, ax, m0, m1, acc);
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&empty[s]);
                    if (k == kb) mbar_arrive(&empty[0]);   // plane kb-1: cm has been consumed by now (see the prologue)
                }
                cm = cc;
                cc = cp;
            }
        }
    }
    if (NACC > 0) grid_reduce_finish<(NACC > 0 ? NACC : 1)>(acc, red);
}

}  // namespace pps

// ================================================================================================
// Fused operand: u = f(in0, in1, in2) is computed on the fly from NIN halo'd input streams, stored once, and
// A u is evaluated in the same pass (PPS_FUSE_FULL: the s- and p-updates of BiCGSTAB.hpp:168-178,262-272 move into
// the operator kernels, 19 -> 17 vector passes per iteration).  f is pointwise, so the value a thread needs at
// its y+-1 / x-edge neighbours is recomputed from the neighbours' inputs in the ring -- bit-identical to what
// the owner stores, no shared-memory exchange, no extra barrier.  Cells outside the solver box evaluate to 0
// because every input is 0 there (work vectors are only ever written inside the box).
// ================================================================================================
namespace pps {

// s = r - alpha v                                                  (BiCGSTAB.hpp:168-178)
template <bool PARITY>
struct PreSUpdate {
    static constexpr int NIN = 2;
    double* out;
    const double* in0;   // r
    const double* in1;   // v
    double alpha;
    __host__ __device__ __forceinline__ const double* input(int i) const { return i == 0 ? in0 : in1; }
    __device__ __forceinline__ void begin(const Ctl* c) { alpha = c->alpha; }
    __device__ __forceinline__ double f(double r, double v, double) const { return submul<PARITY>(r, alpha, v); }
};
// p = r + beta (p - omega v)                                        (BiCGSTAB.hpp:262-272)
template <bool PARITY>
struct PrePUpdate {
    static constexpr int NIN = 3;
    double* out;
    const double* in0;   // r
    const double* in1;   // p (previous)
    const double* in2;   // v (previous)
    double beta, omega;
    __host__ __device__ __forceinline__ const double* input(int i) const { return i == 0 ? in0 : (i == 1 ? in1 : in2); }
    __device__ __forceinline__ void begin(const Ctl* c) { beta = c->beta; omega = c->omega; }
    __device__ __forceinline__ double f(double r, double p, double v) const { return muladd<PARITY>(beta, submul<PARITY>(p, omega, v), r); }
};

template <int BY, int STAGES, int NIN, int NAUX>
struct TmaPreSmem {
    static constexpr int kMainElems = kTmaBoxX * (BY + 2);
    static constexpr int kAuxElems = 64 * BY;
    static constexpr int kStageElems = NIN * kMainElems + NAUX * kAuxElems;
    static constexpr int kMainBytes = NIN * kMainElems * 8;
    static constexpr int kStageBytes = kStageElems * 8;
    static constexpr int kBytes = STAGES * kStageBytes + 2 * STAGES * 8;
};

// The tensor maps are individual __grid_constant__ parameters, exactly like in stencil_tma_kernel (whose maps also change
// from launch to launch, e.g. the rotating Chebyshev buffers), not members of a wrapper struct.
// DBG (diagnostics, pps_debug_fused): bit 0 = cross-proxy fence before a consumer releases a ring stage
template <int BY, int STAGES, bool PARITY, class Pre, class Epi, int DBG = 0>
__global__ void __launch_bounds__(32 * (BY + 1)) stencil_tma_pre_kernel(const __grid_constant__ CUtensorMap map_in0,
                                                                       const __grid_constant__ CUtensorMap map_in1,
                                                                       const __grid_constant__ CUtensorMap map_in2,
                                                                       const __grid_constant__ CUtensorMap map_aux0,
                                                                       const __grid_constant__ CUtensorMap map_aux1, Dims d, Box rg,
                                                                       Coef cf, int zchunk, TileOrigin org, Pre pre, Epi epi,
                                                                       RedCtx red, const Ctl* ctl) {
    if (ctl != nullptr && ctl->done) return;
    constexpr int NACC = Epi::NACC;
    constexpr int NAUX = Epi::NAUX;
    constexpr int NIN = Pre::NIN;
    constexpr int SX = kTmaBoxX;
    using SM = TmaPreSmem<BY, STAGES, NIN, NAUX>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * SM::kStageBytes);
    uint64_t* empty = full + STAGES;

    const int lane = threadIdx.x, warp = threadIdx.y;
    const int col0 = kFirstDataCol + 64 * (blockIdx.x + org.bx0);
    const int y0 = 1 + BY * (blockIdx.y + org.by0);
    const int kchunk = org.kfirst + blockIdx.z * zchunk;
    const int kb = max(rg.k0, kchunk), ke = min(rg.k1, kchunk + zchunk);
    const int nplanes = ke - kb + 2;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], BY);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double acc[NACC > 0 ? NACC : 1];
#pragma unroll
    for (int a = 0; a < (NACC > 0 ? NACC : 1); a++) acc[a] = 0.0;

    if (kb < ke) {
        if (warp == BY) {
            if (lane == 0) {
                for (int p = 0; p < nplanes; p++) {
                    const int s = p % STAGES;
                    if (p >= STAGES) mbar_wait(&empty[s], ((p / STAGES) - 1) & 1);
                    double* dst = ring + s * SM::kStageElems;
                    const bool with_aux = NAUX > 0 && p >= 1 && p <= nplanes - 2;
                    const int kk = kb - 1 + p;
                    mbar_arrive_expect_tx(&full[s], with_aux ? SM::kStageBytes : SM::kMainBytes);
                    tma_load_3d(dst, &map_in0, &full[s], col0 - kTmaLead, y0 - 1, kk);
                    if (NIN > 1) tma_load_3d(dst + SM::kMainElems, &map_in1, &full[s], col0 - kTmaLead, y0 - 1, kk);
                    if (NIN > 2) tma_load_3d(dst + 2 * SM::kMainElems, &map_in2, &full[s], col0 - kTmaLead, y0 - 1, kk);
                    if (with_aux) {
                        if (NAUX > 0) tma_load_3d(dst + NIN * SM::kMainElems, &map_aux0, &full[s], col0, y0, kk);
                        if (NAUX > 1) tma_load_3d(dst + NIN * SM::kMainElems + SM::kAuxElems, &map_aux1, &full[s], col0, y0, kk);
                    }
                }
            }
        } else {
            pre.begin(ctl);
            const int col = col0 + 2 * lane;
            const int j = y0 + warp;
            const int i0 = col - kOff;
            const bool jr = j >= rg.j0 && j < rg.j1;
            const bool m0 = jr && i0 >= rg.i0 && i0 < rg.i1;
            const bool m1 = jr && i0 + 1 >= rg.i0 && i0 + 1 < rg.i1;
            const bool warp_any = __any_sync(kFullMask, m0 || m1);
            const long long colc = col < d.pitch ? col : d.pitch - 2;
            const long long rowoff = colc + d.pitch * min(j, d.ny);
            const int so = (warp + 1) * SX + kTmaLead + 2 * lane;
            const int sa = NIN * SM::kMainElems + warp * 64 + 2 * lane;

            // u at two adjacent cells / one cell, from the inputs of ring stage `stg`
            auto eval2 = [&](const double* stg, int off) {
                const double2 a = *reinterpret_cast<const double2*>(stg + off);
                double2 b = make_double2(0.0, 0.0), c = make_double2(0.0, 0.0);
                if (NIN > 1) b = *reinterpret_cast<const double2*>(stg + SM::kMainElems + off);
                if (NIN > 2) c = *reinterpret_cast<const double2*>(stg + 2 * SM::kMainElems + off);
                return make_double2(pre.f(a.x, b.x, c.x), pre.f(a.y, b.y, c.y));
            };
            auto eval1 = [&](const double* stg, int off) {
                const double a = stg[off];
                const double b = NIN > 1 ? stg[SM::kMainElems + off] : 0.0;
                const double c = NIN > 2 ? stg[2 * SM::kMainElems + off] : 0.0;
                return pre.f(a, b, c);
            };

            // stage 0 is released at the end of the first iteration (see stencil_tma_kernel: loads must be consumed first)
            mbar_wait(&full[0], 0);
            double2 cm = eval2(ring, so);
            mbar_wait(&full[1 % STAGES], (1 / STAGES) & 1);
            double2 cc = eval2(ring + (1 % STAGES) * SM::kStageElems, so);
            for (int k = kb; k < ke; ++k) {
                const int p = k - kb + 1;
                const int s = p % STAGES, sn = (p + 1) % STAGES;
                mbar_wait(&full[sn], ((p + 1) / STAGES) & 1);
                const double* stg = ring + s * SM::kStageElems;
                const double2 cp = eval2(ring + sn * SM::kStageElems, so);
                if (warp_any) {
                    const double2 ym = eval2(stg, so - SX);
                    const double2 yp = eval2(stg, so + SX);
                    double2 ax[NAUX > 0 ? NAUX : 1];
#pragma unroll
                    for (int a = 0; a < NAUX; a++) ax[a] = *reinterpret_cast<const double2*>(stg + sa + a * SM::kAuxElems);
                    double xl = __shfl_up_sync(kFullMask, cc.y, 1);
                    double xr = __shfl_down_sync(kFullMask, cc.x, 1);
                    if (lane == 0) xl = eval1(stg, so - 1);
                    if (lane == 31) xr = eval1(stg, so + 2);
                    double2 au;
                    au.x = laplacian<PARITY>(cf, xl, cc.x, cc.y, ym.x, yp.x, cm.x, cp.x);
                    au.y = laplacian<PARITY>(cf, cc.x, cc.y, xr, ym.y, yp.y, cm.y, cp.y);
                    const long long idx = rowoff + k * d.plane;
                    st2(pre.out + idx, cc, m0, m1);
                    epi(idx, au, cc, ax, m0, m1, acc);
                }
                if (DBG & 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&empty[s]);
                    if (k == kb) mbar_arrive(&empty[0]);
                }
                cm = cc;
                cc = cp;
            }
        }
    }
    if (NACC > 0) grid_reduce_finish<(NACC > 0 ? NACC : 1)>(acc, red);
}

}  // namespace pps

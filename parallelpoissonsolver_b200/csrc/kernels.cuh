// kernels.cuh -- hand-written sm_100a kernels of the Poisson hot path.
//
// Every kernel works on the pitched guard-padded layout of common.cuh.  Thread mapping shared by the
// z-marching kernels: a warp owns 64 consecutive columns of one row (one double2 = 16 B per lane, rows are
// 128-B aligned so a warp request is exactly four 128-B lines), a CTA owns BY rows (tile 64 x BY) and
// marches through its z-chunk keeping the k-1 / k / k+1 values of its own column in registers; x-neighbours
// come from warp shuffles (two extra scalar loads per warp at the tile edge), y-neighbours from L1/L2.
// Reductions: per-thread partials over the z-march -> warp shuffle tree -> shared-memory block tree ->
// one store per CTA -> the last CTA to arrive (atomic ticket) sums the partials in index order, so a
// reduction is bit-reproducible for a fixed launch shape (the reference's alpaka kernels use racy atomics).
#pragma once

#include "common.cuh"

namespace pps {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v, bool m0, bool m1) {
    if (m0 && m1) *reinterpret_cast<double2*>(p) = v;
    else if (m0) p[0] = v.x;
    else if (m1) p[1] = v.y;
}

// a*b + c with the reference's two roundings (PARITY) or one FMA (fast)
template <bool PARITY>
__device__ __forceinline__ double muladd(double a, double b, double c) {
    if (PARITY) return __dadd_rn(__dmul_rn(a, b), c);
    return fma(a, b, c);
}
// c - a*b
template <bool PARITY>
__device__ __forceinline__ double submul(double c, double a, double b) {
    if (PARITY) return __dsub_rn(c, __dmul_rn(a, b));
    return fma(-a, b, c);
}

// matrixFreeOperatorA.hpp:33-38: (u[-1] - 2u + u[+1])/(dx*dx) + (...)/(dy*dy) + (...)/(dz*dz), in this order
template <bool PARITY>
__device__ __forceinline__ double laplacian(const Coef& cf, double xm, double c, double xp, double ym, double yp,
                                            double zm, double zp) {
    if (PARITY) {
        const double c2 = __dmul_rn(2.0, c);
        const double tx = __ddiv_rn(__dadd_rn(__dsub_rn(xm, c2), xp), cf.ds2[0]);
        const double ty = __ddiv_rn(__dadd_rn(__dsub_rn(ym, c2), yp), cf.ds2[1]);
        const double tz = __ddiv_rn(__dadd_rn(__dsub_rn(zm, c2), zp), cf.ds2[2]);
        return __dadd_rn(__dadd_rn(tx, ty), tz);
    } else {
        const double dx = (fma(-2.0, c, xm) + xp);
        const double dy = (fma(-2.0, c, ym) + yp);
        const double dz = (fma(-2.0, c, zm) + zp);
        return fma(dz, cf.inv[2], fma(dy, cf.inv[1], dx * cf.inv[0]));
    }
}

// ------------------------------------------------------------------------------------------------
// scalar updates after a fused reduction (see ScalarOp)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void apply_scalar_op(int op, Ctl* c) {
    const double* s = c->sums;
    switch (op) {
        case OP_BICG_ALPHA:
            c->alpha = c->rho0 / s[0];
            break;
        case OP_BICG_OMEGA:
            c->omega = s[0] / s[1];
            break;
        case OP_BICG_RHO: {
            const double rho1 = s[0];
            const double err = sqrt(s[1]);
            c->beta = rho1 / c->rho0 * c->alpha / c->omega;   // BiCGSTAB.hpp:258
            c->rho0 = rho1;
            c->err = err;
            int it = c->iter;
            c->hist_alpha[it] = c->alpha;
            c->hist_omega[it] = c->omega;
            c->hist_rho[it] = rho1;
            it++;
            c->iter = it;
            c->hist_err[it] = err;
            if (err < c->tol || it >= c->max_iter) c->done = 1;
            break;
        }
        case OP_NORM_B:
            c->norm_b = sqrt(s[0]);
            break;
        case OP_RESIDUAL0:
            c->err = sqrt(s[0]);
            c->hist_err[0] = c->err;
            if (c->err < c->tol || c->max_iter <= 0) c->done = 1;
            break;
        case OP_RESIDUAL_FINAL:
            c->sums[4] = sqrt(s[0]);
            break;
        case OP_CG_ALPHA:
            c->rz = s[0];
            c->alpha = s[0] / s[1];
            break;
        case OP_CG_BETA: {
            const double err = sqrt(s[1]);
            c->beta = s[0] / c->rz;                            // baseCG.hpp:187
            c->err = err;
            int it = c->iter;
            c->hist_alpha[it] = c->alpha;
            c->hist_omega[it] = c->beta;
            c->hist_rho[it] = s[0];
            it++;
            c->iter = it;
            c->hist_err[it] = err;
            if (err < c->tol || it >= c->max_iter) c->done = 1;
            break;
        }
        default:
            break;
    }
}

static __global__ void scalar_op_kernel(int op, Ctl* c, int ignore_done) {
    if (!ignore_done && c->done) return;
    apply_scalar_op(op, c);
}

// ------------------------------------------------------------------------------------------------
// block -> grid reduction with an atomic ticket; deterministic for a fixed launch shape
// ------------------------------------------------------------------------------------------------
template <int NACC>
__device__ __forceinline__ void block_sum(double (&acc)[NACC], double (*sm)[32], int tid, int nwarps) {
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int a = 0; a < NACC; a++) {
        double v = acc[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFullMask, v, o);
        if (lane == 0) sm[a][wid] = v;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int a = 0; a < NACC; a++) {
            double v = lane < nwarps ? sm[a][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFullMask, v, o);
            acc[a] = v;   // valid in thread 0
        }
    }
}

template <int NACC>
__device__ __forceinline__ void grid_reduce_finish(double (&acc)[NACC], const RedCtx& red) {
    __shared__ double sm[NACC][32];
    __shared__ unsigned int s_ticket;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;
    const int nwarps = nthreads >> 5;
    block_sum<NACC>(acc, sm, tid, nwarps);
    const unsigned int cta = red.cta_offset + blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (tid == 0) {
#pragma unroll
        for (int a = 0; a < NACC; a++) red.partials[a * red.capacity + cta] = acc[a];
        __threadfence();
        s_ticket = atomicAdd(red.counter, 1u);
    }
    __syncthreads();
    if (s_ticket != red.total_ctas - 1) return;
    // last CTA: every partial is visible; sum them in index order
    __threadfence();
    double tot[NACC];
#pragma unroll
    for (int a = 0; a < NACC; a++) {
        double v = 0.0;
        for (unsigned int i = tid; i < red.total_ctas; i += nthreads) v += __ldcg(&red.partials[a * red.capacity + i]);
        tot[a] = v;
    }
    __syncthreads();
    block_sum<NACC>(tot, sm, tid, nwarps);
    if (red.pr.epoch != 0) {
        // ---- allreduce over peer memory, fused into this kernel (see PeerReduce)
        __shared__ double s_tot[NACC];
        if (tid == 0) {
#pragma unroll
            for (int a = 0; a < NACC; a++) s_tot[a] = tot[a];
        }
        __syncthreads();
        const int W = red.pr.world, R = red.pr.rank;
        const unsigned long long e = red.pr.epoch;
        const int par = static_cast<int>(e & 1ull);
        if (tid < W) {
            double* dst = red.pr.peer_mail[tid] + (static_cast<long long>(par) * W + R) * 4;
#pragma unroll
            for (int a = 0; a < NACC; a++) asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(dst + a), "d"(s_tot[a]) : "memory");
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst + 3), "l"(e) : "memory");
            const double* src = red.pr.my_mail + (static_cast<long long>(par) * W + tid) * 4;
            unsigned long long seen;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(src + 3) : "memory");
                if (seen == e) break;
                __nanosleep(50);
            } while (true);
        }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int a = 0; a < NACC; a++) {
                double v = 0.0;
                for (int r = 0; r < W; r++) {
                    double x;
                    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(x) : "l"(red.pr.my_mail + (static_cast<long long>(par) * W + r) * 4 + a) : "memory");
                    v += x;
                }
                tot[a] = v;
            }
        }
    }
    if (tid == 0) {
#pragma unroll
        for (int a = 0; a < NACC; a++) red.ctl->sums[a] = tot[a];
        *red.counter = 0;
        if (red.op != OP_NONE) apply_scalar_op(red.op, red.ctl);
    }
}

// ------------------------------------------------------------------------------------------------
// thread -> cell mapping shared by the z-marching kernels
// ------------------------------------------------------------------------------------------------
struct CellMap {
    long long rowoff;   // offset of (col, j) inside a plane (clamped to valid memory)
    int kb, ke;         // z range of this CTA intersected with the box
    bool m0, m1;        // the two cells of this thread are inside the box
    bool warp_active;
};

template <int BY>
__device__ __forceinline__ CellMap make_cellmap(const Dims& d, const Box& rg, int zchunk, const TileOrigin& org) {
    CellMap m;
    const int col = kFirstDataCol + 2 * ((blockIdx.x + org.bx0) * 32 + threadIdx.x);
    const int j = 1 + (blockIdx.y + org.by0) * BY + threadIdx.y;
    const int i0 = col - kOff;
    const bool jr = j >= rg.j0 && j < rg.j1;
    m.m0 = jr && i0 >= rg.i0 && i0 < rg.i1;
    m.m1 = jr && i0 + 1 >= rg.i0 && i0 + 1 < rg.i1;
    const int kb = org.kfirst + blockIdx.z * zchunk;
    m.kb = max(rg.k0, kb);
    m.ke = min(rg.k1, kb + zchunk);
    // lanes right of the row keep valid addresses; the first column right of the data range (the x+ guard)
    // is still loaded truthfully because its left neighbour needs it through the shuffle
    const long long colc = col < d.pitch ? col : d.pitch - 2;
    const int jc = min(j, d.ny);
    m.rowoff = colc + d.pitch * jc;
    m.warp_active = __any_sync(kFullMask, m.m0 || m.m1) && m.kb < m.ke;
    return m;
}

// ------------------------------------------------------------------------------------------------
// operator family:  Au = A*u on the box, then an epilogue functor fuses what follows in the algorithm
// (store, dot products, residual, Chebyshev recurrences).  Epi::operator()(idx, Au, u, m0, m1, acc).
// ------------------------------------------------------------------------------------------------
template <int BY, bool PARITY, class Epi>
__global__ void __launch_bounds__(32 * BY) stencil_kernel(const double* __restrict__ u, Dims d, Box rg, Coef cf,
                                                         int zchunk, TileOrigin org, Epi epi, RedCtx red, const Ctl* ctl) {
    if (ctl != nullptr && ctl->done) return;
    constexpr int NACC = Epi::NACC;
    double acc[NACC > 0 ? NACC : 1];
#pragma unroll
    for (int a = 0; a < (NACC > 0 ? NACC : 1); a++) acc[a] = 0.0;
    const CellMap m = make_cellmap<BY>(d, rg, zchunk, org);
    if (m.warp_active) {
        const int lane = threadIdx.x;
        const double* up = u + m.rowoff;
        double2 cm = ldg2(up + (m.kb - 1) * d.plane);
        double2 cc = ldg2(up + m.kb * d.plane);
#pragma unroll 2
        for (int k = m.kb; k < m.ke; ++k) {
            const double* pk = up + k * d.plane;
            const double2 cp = ldg2(pk + d.plane);
            const double2 ym = ldg2(pk - d.pitch);
            const double2 yp = ldg2(pk + d.pitch);
            double xl = __shfl_up_sync(kFullMask, cc.y, 1);
            double xr = __shfl_down_sync(kFullMask, cc.x, 1);
            if (lane == 0) xl = __ldg(pk - 1);
            if (lane == 31) xr = __ldg(pk + 2);
            double2 au;
            au.x = laplacian<PARITY>(cf, xl, cc.x, cc.y, ym.x, yp.x, cm.x, cp.x);
            au.y = laplacian<PARITY>(cf, cc.x, cc.y, xr, ym.y, yp.y, cm.y, cp.y);
            const long long idx = m.rowoff + k * d.plane;
            double2 ax[Epi::NAUX > 0 ? Epi::NAUX : 1];
#pragma unroll
            for (int a = 0; a < Epi::NAUX; a++) ax[a] = ldg2(epi.aux(a) + idx);
            epi(idx, au, cc, ax, m.m0, m.m1, acc);
            cm = cc;
            cc = cp;
        }
    }
    if (NACC > 0) grid_reduce_finish<(NACC > 0 ? NACC : 1)>(acc, red);
}

// Epilogue functors.  NACC = fused reductions, NAUX = extra input streams read at the same cell (aux(i) names
// them: the plain-load kernel fetches them with ldg, the TMA kernel stages them through its ring).
// operator()(idx, Au, u, ax, m0, m1, acc): idx = element offset of the thread's first cell, ax[i] = aux values.

// out = A u                                                      (BiCGSTAB.hpp:189-199 loop nest)
struct EpiStore {
    static constexpr int NACC = 0, NAUX = 0;
    double* out;
    __host__ __device__ __forceinline__ const double* aux(int) const { return nullptr; }
    __device__ __forceinline__ void operator()(long long idx, double2 au, double2, const double2*, bool m0, bool m1, double*) const {
        st2(out + idx, au, m0, m1);
    }
};
// v = A p ; acc0 += w.v                                           (BiCGSTAB.hpp:142-155)
struct EpiStoreDot {
    static constexpr int NACC = 1, NAUX = 1;
    double* out;
    const double* w;
    __host__ __device__ __forceinline__ const double* aux(int) const { return w; }
    __device__ __forceinline__ void operator()(long long idx, double2 au, double2, const double2* ax, bool m0, bool m1, double* acc) const {
        st2(out + idx, au, m0, m1);
        acc[0] += (m0 ? ax[0].x * au.x : 0.0) + (m1 ? ax[0].y * au.y : 0.0);
    }
};
// t = A z ; acc0 += r.t ; acc1 += t.t                              (BiCGSTAB.hpp:189-214)
struct EpiStoreDot2 {
    static constexpr int NACC = 2, NAUX = 1;
    double* out;
    const double* w;
    __host__ __device__ __forceinline__ const double* aux(int) const { return w; }
    __device__ __forceinline__ void operator()(long long idx, double2 au, double2, const double2* ax, bool m0, bool m1, double* acc) const {
        st2(out + idx, au, m0, m1);
        acc[0] += (m0 ? ax[0].x * au.x : 0.0) + (m1 ? ax[0].y * au.y : 0.0);
        acc[1] += (m0 ? au.x * au.x : 0.0) + (m1 ? au.y * au.y : 0.0);
    }
};
// same without a preconditioner: z is r itself, so the operand doubles as the dot-product partner
struct EpiStoreDot2Self {
    static constexpr int NACC = 2, NAUX = 0;
    double* out;
    __host__ __device__ __forceinline__ const double* aux(int) const { return nullptr; }
    __device__ __forceinline__ void operator()(long long idx, double2 au, double2 u, const double2*, bool m0, bool m1, double* acc) const {
        st2(out + idx, au, m0, m1);
        acc[0] += (m0 ? u.x * au.x : 0.0) + (m1 ? u.y * au.y : 0.0);
        acc[1] += (m0 ? au.x * au.x : 0.0) + (m1 ? au.y * au.y : 0.0);
    }
};
// r = b - A x ; acc0 += r.r                                       (iterativeSolverBase.hpp:257-268)
template <bool PARITY>
struct EpiResidual {
    static constexpr int NACC = 1, NAUX = 1;
    double* r;
    const double* b;
    __host__ __device__ __forceinline__ const double* aux(int) const { return b; }
    __device__ __forceinline__ void operator()(long long idx, double2 au, double2, const double2* ax, bool m0, bool m1, double* acc) const {
        double2 rv;
        rv.x = PARITY ? __dsub_rn(ax[0].x, au.x) : ax[0].x - au.x;
        rv.y = PARITY ? __dsub_rn(ax[0].y, au.y) : ax[0].y - au.y;
        st2(r + idx, rv, m0, m1);
        acc[0] += (m0 ? rv.x * rv.x : 0.0) + (m1 ? rv.y * rv.y : 0.0);
    }
};
// Ap = A p ; acc0 += r.z ; acc1 += p.Ap                           (baseCG.hpp:126-140)
struct EpiCgApply {
    static constexpr int NACC = 2, NAUX = 2;
    double* out;
    const double* r;
    const double* z;
    __host__ __device__ __forceinline__ const double* aux(int i) const { return i == 0 ? r : z; }
    __device__ __forceinline__ void operator()(long long idx, double2 au, double2 u, const double2* ax, bool m0, bool m1, double* acc) const {
        st2(out + idx, au, m0, m1);
        acc[0] += (m0 ? ax[0].x * ax[1].x : 0.0) + (m1 ? ax[0].y * ax[1].y : 0.0);
        acc[1] += (m0 ? u.x * au.x : 0.0) + (m1 ? u.y * au.y : 0.0);
    }
};
// same without a preconditioner (z is r): acc0 += r.r
struct EpiCgApplySelf {
    static constexpr int NACC = 2, NAUX = 1;
    double* out;
    const double* r;
    __host__ __device__ __forceinline__ const double* aux(int) const { return r; }
    __device__ __forceinline__ void operator()(long long idx, double2 au, double2 u, const double2* ax, bool m0, bool m1, double* acc) const {
        st2(out + idx, au, m0, m1);
        acc[0] += (m0 ? ax[0].x * ax[0].x : 0.0) + (m1 ? ax[0].y * ax[0].y : 0.0);
        acc[1] += (m0 ? u.x * au.x : 0.0) + (m1 ? u.y * au.y : 0.0);
    }
};
// Chebyshev start: Z = B/theta ; Y = (2 rho1/delta) (2 B + A B/theta)       (chebyshevIteration.hpp:79-90)
// Y may be the final output with the sign flipped (chebyshevMax == 3 -> X = -y1)
template <bool PARITY>
struct EpiChebFirst {
    static constexpr int NACC = 0, NAUX = 0;
    double* Z;
    double* Y;
    double theta, inv_theta, c1, ysign;
    __host__ __device__ __forceinline__ const double* aux(int) const { return nullptr; }
    __device__ __forceinline__ double y(double b, double ab) const {
        if (PARITY) return __dmul_rn(c1, __dadd_rn(__dmul_rn(2.0, b), __ddiv_rn(ab, theta)));
        return c1 * fma(ab, inv_theta, 2.0 * b);
    }
    __device__ __forceinline__ void operator()(long long idx, double2 au, double2 u, const double2*, bool m0, bool m1, double*) const {
        double2 zv, yv;
        zv.x = PARITY ? __ddiv_rn(u.x, theta) : u.x * inv_theta;
        zv.y = PARITY ? __ddiv_rn(u.y, theta) : u.y * inv_theta;
        yv.x = ysign * y(u.x, au.x);
        yv.y = ysign * y(u.y, au.y);
        if (Z != nullptr) st2(Z + idx, zv, m0, m1);
        st2(Y + idx, yv, m0, m1);
    }
};
// Chebyshev step: W = rho (2 sigma Y + (2/delta)(B + A Y) - rho_old Z)       (chebyshevIteration.hpp:103-113)
// wsign = -1 folds the final X = -W (:118-128) into the last live sweep
template <bool PARITY>
struct EpiChebStep {
    static constexpr int NACC = 0, NAUX = 2;
    double* W;
    const double* B;
    const double* Z;
    double rho, rho_old, two_sigma, two_over_delta, wsign;
    __host__ __device__ __forceinline__ const double* aux(int i) const { return i == 0 ? B : Z; }
    __device__ __forceinline__ double w(double y, double ay, double b, double z) const {
        if (PARITY)
            return __dmul_rn(rho, __dsub_rn(__dadd_rn(__dmul_rn(two_sigma, y), __dmul_rn(two_over_delta, __dadd_rn(b, ay))),
                                            __dmul_rn(rho_old, z)));
        return rho * fma(-rho_old, z, fma(two_sigma, y, two_over_delta * (b + ay)));
    }
    __device__ __forceinline__ void operator()(long long idx, double2 au, double2 u, const double2* ax, bool m0, bool m1, double*) const {
        double2 wv;
        wv.x = wsign * w(u.x, au.x, ax[0].x, ax[1].x);
        wv.y = wsign * w(u.y, au.y, ax[0].y, ax[1].y);
        st2(W + idx, wv, m0, m1);
    }
};

// ------------------------------------------------------------------------------------------------
// pointwise family (same tiling, no stencil): axpy updates fused with their reductions
// ------------------------------------------------------------------------------------------------
template <int BY, class Op>
__global__ void __launch_bounds__(32 * BY) pointwise_kernel(Dims d, Box rg, int zchunk, TileOrigin org, Op op, RedCtx red,
                                                           const Ctl* ctl) {
    if (ctl != nullptr && ctl->done) return;
    constexpr int NACC = Op::NACC;
    double acc[NACC > 0 ? NACC : 1];
#pragma unroll
    for (int a = 0; a < (NACC > 0 ? NACC : 1); a++) acc[a] = 0.0;
    const CellMap m = make_cellmap<BY>(d, rg, zchunk, org);
    if (m.warp_active && (m.m0 || m.m1)) {
        op.begin(ctl);
#pragma unroll 4
        for (int k = m.kb; k < m.ke; ++k) op(m.rowoff + k * d.plane, m.m0, m.m1, acc);
    }
    if (NACC > 0) grid_reduce_finish<(NACC > 0 ? NACC : 1)>(acc, red);
}

// r <- r - alpha v                                                 (BiCGSTAB.hpp:168-178)
template <bool PARITY>
struct OpSUpdate {
    static constexpr int NACC = 0;
    double* r;
    const double* v;
    double alpha;
    __device__ __forceinline__ void begin(const Ctl* c) { alpha = c->alpha; }
    __device__ __forceinline__ void operator()(long long idx, bool m0, bool m1, double*) const {
        double2 rv = ld2(r + idx);
        const double2 vv = ldg2(v + idx);
        rv.x = submul<PARITY>(rv.x, alpha, vv.x);
        rv.y = submul<PARITY>(rv.y, alpha, vv.y);
        st2(r + idx, rv, m0, m1);
    }
};
// x += alpha Mp + omega z ; r <- r - omega t ; acc0 += r0.r ; acc1 += r.r        (BiCGSTAB.hpp:227-246)
template <bool PARITY>
struct OpXRUpdate {
    static constexpr int NACC = 2;
    double* x;
    double* r;
    const double* mp;
    const double* z;    // may alias r (no preconditioner: z is r before this update)
    const double* t;
    const double* r0;
    double alpha, omega;
    __device__ __forceinline__ void begin(const Ctl* c) { alpha = c->alpha; omega = c->omega; }
    __device__ __forceinline__ void operator()(long long idx, bool m0, bool m1, double* acc) const {
        double2 xv = ld2(x + idx);
        double2 rv = ld2(r + idx);
        const double2 pv = ldg2(mp + idx);
        double2 zv = rv;
        if (z != r) zv = ldg2(z + idx);
        const double2 tv = ldg2(t + idx);
        const double2 qv = ldg2(r0 + idx);
        xv.x = muladd<PARITY>(omega, zv.x, muladd<PARITY>(alpha, pv.x, xv.x));
        xv.y = muladd<PARITY>(omega, zv.y, muladd<PARITY>(alpha, pv.y, xv.y));
        rv.x = submul<PARITY>(rv.x, omega, tv.x);
        rv.y = submul<PARITY>(rv.y, omega, tv.y);
        st2(x + idx, xv, m0, m1);
        st2(r + idx, rv, m0, m1);
        acc[0] += (m0 ? qv.x * rv.x : 0.0) + (m1 ? qv.y * rv.y : 0.0);
        acc[1] += (m0 ? rv.x * rv.x : 0.0) + (m1 ? rv.y * rv.y : 0.0);
    }
};
// fused schedule (PPS_FUSE_FULL), s lives in its own array:
// x += alpha p + omega s ; r <- s - omega t ; acc0 += r0.r ; acc1 += r.r            (BiCGSTAB.hpp:227-246)
template <bool PARITY>
struct OpXRUpdateS {
    static constexpr int NACC = 2;
    double* x;
    double* r;
    const double* p;
    const double* s;
    const double* t;
    const double* r0;
    double alpha, omega;
    __device__ __forceinline__ void begin(const Ctl* c) { alpha = c->alpha; omega = c->omega; }
    __device__ __forceinline__ void operator()(long long idx, bool m0, bool m1, double* acc) const {
        double2 xv = ld2(x + idx);
        const double2 pv = ldg2(p + idx);
        const double2 sv = ldg2(s + idx);
        const double2 tv = ldg2(t + idx);
        const double2 qv = ldg2(r0 + idx);
        double2 rv;
        xv.x = muladd<PARITY>(omega, sv.x, muladd<PARITY>(alpha, pv.x, xv.x));
        xv.y = muladd<PARITY>(omega, sv.y, muladd<PARITY>(alpha, pv.y, xv.y));
        rv.x = submul<PARITY>(sv.x, omega, tv.x);
        rv.y = submul<PARITY>(sv.y, omega, tv.y);
        st2(x + idx, xv, m0, m1);
        st2(r + idx, rv, m0, m1);
        acc[0] += (m0 ? qv.x * rv.x : 0.0) + (m1 ? qv.y * rv.y : 0.0);
        acc[1] += (m0 ? rv.x * rv.x : 0.0) + (m1 ? rv.y * rv.y : 0.0);
    }
};
// p <- r + beta (p - omega v)                                       (BiCGSTAB.hpp:262-272)
template <bool PARITY>
struct OpPUpdate {
    static constexpr int NACC = 0;
    double* p;
    const double* r;
    const double* v;
    double beta, omega;
    __device__ __forceinline__ void begin(const Ctl* c) { beta = c->beta; omega = c->omega; }
    __device__ __forceinline__ void operator()(long long idx, bool m0, bool m1, double*) const {
        double2 pv = ld2(p + idx);
        const double2 rv = ldg2(r + idx);
        const double2 vv = ldg2(v + idx);
        pv.x = muladd<PARITY>(beta, submul<PARITY>(pv.x, omega, vv.x), rv.x);
        pv.y = muladd<PARITY>(beta, submul<PARITY>(pv.y, omega, vv.y), rv.y);
        st2(p + idx, pv, m0, m1);
    }
};
// acc0 += a.b ; acc1 += a.a   (b == nullptr: only a.a in acc0)      (iterativeSolverBase.hpp:182-192, baseCG.hpp:171-182)
struct OpDot {
    static constexpr int NACC = 2;
    const double* a;
    const double* b;
    __device__ __forceinline__ void begin(const Ctl*) {}
    __device__ __forceinline__ void operator()(long long idx, bool m0, bool m1, double* acc) const {
        const double2 av = ldg2(a + idx);
        if (b != nullptr) {
            const double2 bv = ldg2(b + idx);
            acc[0] += (m0 ? av.x * bv.x : 0.0) + (m1 ? av.y * bv.y : 0.0);
            acc[1] += (m0 ? av.x * av.x : 0.0) + (m1 ? av.y * av.y : 0.0);
        } else {
            acc[0] += (m0 ? av.x * av.x : 0.0) + (m1 ? av.y * av.y : 0.0);
        }
    }
};
// x += alpha p ; r -= alpha Ap ; acc0 += r.r ; acc1 += r.r           (baseCG.hpp:154-165, :171-182 when z == r)
template <bool PARITY>
struct OpCgXR {
    static constexpr int NACC = 2;
    double* x;
    double* r;
    const double* p;
    const double* ap;
    double alpha;
    __device__ __forceinline__ void begin(const Ctl* c) { alpha = c->alpha; }
    __device__ __forceinline__ void operator()(long long idx, bool m0, bool m1, double* acc) const {
        double2 xv = ld2(x + idx);
        double2 rv = ld2(r + idx);
        const double2 pv = ldg2(p + idx);
        const double2 av = ldg2(ap + idx);
        xv.x = muladd<PARITY>(alpha, pv.x, xv.x);
        xv.y = muladd<PARITY>(alpha, pv.y, xv.y);
        rv.x = submul<PARITY>(rv.x, alpha, av.x);
        rv.y = submul<PARITY>(rv.y, alpha, av.y);
        st2(x + idx, xv, m0, m1);
        st2(r + idx, rv, m0, m1);
        const double s = (m0 ? rv.x * rv.x : 0.0) + (m1 ? rv.y * rv.y : 0.0);
        acc[0] += s;
        acc[1] += s;
    }
};
// p <- z + beta p                                                   (baseCG.hpp:197-208)
template <bool PARITY>
struct OpCgP {
    static constexpr int NACC = 0;
    double* p;
    const double* z;
    double beta;
    __device__ __forceinline__ void begin(const Ctl* c) { beta = c->beta; }
    __device__ __forceinline__ void operator()(long long idx, bool m0, bool m1, double*) const {
        double2 pv = ld2(p + idx);
        const double2 zv = ldg2(z + idx);
        pv.x = muladd<PARITY>(beta, pv.x, zv.x);
        pv.y = muladd<PARITY>(beta, pv.y, zv.y);
        st2(p + idx, pv, m0, m1);
    }
};

// ------------------------------------------------------------------------------------------------
// whole-array kernels (guards and padding included, like the reference's 0..ntot loops)
// ------------------------------------------------------------------------------------------------
// mode 0: a /= norm_b (iterativeSolverBase.hpp:227-231)   mode 1: a *= norm_b (BiCGSTAB.hpp:310-314)
static __global__ void scale_kernel(double* __restrict__ a, double* __restrict__ b, long long n, const Ctl* ctl, int mode,
                             int ignore_done) {
    if (!ignore_done && ctl->done) return;
    const double s = ctl->norm_b;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (mode == 0) {
            a[i] = __ddiv_rn(a[i], s);
            b[i] = __ddiv_rn(b[i], s);
        } else {
            a[i] = __dmul_rn(a[i], s);
            b[i] = __dmul_rn(b[i], s);
        }
    }
}

// ---- nested (block-local) Krylov preconditioners: BiCGSTAB / BaseCG with isMainLoop = false, communicationOFF ----
// start values of a nested solve (BiCGSTAB.hpp:60-66,127; baseCG.hpp:49-66).  When the OUTER solve has already converged the
// nested solve starts `done`, so every kernel of it returns at once, and entry 0 of its history tells the host to stop.
static __global__ void inner_begin_kernel(Ctl* c, const Ctl* outer, double tol, int max_iter) {
    c->rho0 = 1; c->alpha = 1; c->omega = 1; c->beta = 1; c->err = -1; c->rz = 1; c->norm_b = 1; c->tol = tol;
    for (int q = 0; q < 8; q++) c->sums[q] = 0;
    c->iter = 0; c->max_iter = max_iter; c->pad = 0;
    c->done = (outer != nullptr && outer->done) ? 1 : 0;
    c->hist_err[0] = c->done ? -1.0 : 1e300;
}
// a, b /= (mode 0) or *= (mode 1) the norm the NESTED solve normalised with, over every entry (iterativeSolverBase.hpp:227-231,
// BiCGSTAB.hpp:310-314); skipped when the OUTER solve is done.  a may be null.
static __global__ void inner_scale_kernel(double* __restrict__ a, double* __restrict__ b, long long n, const Ctl* inner, const Ctl* outer,
                                   int mode) {
    if (outer != nullptr && outer->done) return;
    const double s = inner->norm_b;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (mode == 0) {
            if (a != nullptr) a[i] = __ddiv_rn(a[i], s);
            b[i] = __ddiv_rn(b[i], s);
        } else {
            if (a != nullptr) a[i] = __dmul_rn(a[i], s);
            b[i] = __dmul_rn(b[i], s);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// face kernels: O(N^2) work on one face of a block
// ------------------------------------------------------------------------------------------------
struct FaceGeom {
    long long base_a;     // first cell of plane A (meaning depends on the kernel)
    long long base_b;     // first cell of plane B
    long long stride_u;   // element stride of the fast tangential axis
    long long stride_v;   // element stride of the slow tangential axis
    int nu, nv;           // tangential extent (data range)
};

// resetNeumanBCs (iterativeSolverBase.hpp:62-169), orderNeumanBcs = 2:
//   ghost(A) = mirror(B)                                  helper fields
//   ghost(A) = mirror(B) -/+ 2 ds dudn / norm_b           solution field in the main loop (:100, :148)
// orderNeumanBcs == 1 (:92-95, :105-108, :140-143, :153-156): the host passes the boundary plane as B and ds as `two_ds`
static __global__ void neumann_ghost_kernel(double* __restrict__ f, FaceGeom g, const double* __restrict__ dudn, double two_ds,
                                     int upper, const Ctl* ctl, int ignore_done) {
    if (!ignore_done && ctl != nullptr && ctl->done) return;
    const long long n = static_cast<long long>(g.nu) * g.nv;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long off = (t % g.nu) * g.stride_u + (t / g.nu) * g.stride_v;
        double v = f[g.base_b + off];
        if (dudn != nullptr) {
            const double corr = __ddiv_rn(__dmul_rn(two_ds, dudn[t]), ctl->norm_b);
            v = upper ? __dadd_rn(v, corr) : __dsub_rn(v, corr);
        }
        f[g.base_a + off] = v;
    }
}

// all Neumann faces of a block in ONE launch (blockIdx.y = face slot): same arithmetic as neumann_ghost_kernel.  The faces
// write disjoint ghost cells and read data cells only, so their order does not matter (iterativeSolverBase.hpp:69-168).
struct GhostBatch {
    FaceGeom g[6];
    const double* dudn[6];
    double two_ds[6];
    int upper[6];
    int count;
};
static __global__ void neumann_ghost_batch_kernel(double* __restrict__ f, GhostBatch batch, const Ctl* ctl, int ignore_done) {
    if (!ignore_done && ctl != nullptr && ctl->done) return;
    const int q = blockIdx.y;
    if (q >= batch.count) return;
    const FaceGeom g = batch.g[q];
    const double* dudn = batch.dudn[q];
    const long long n = static_cast<long long>(g.nu) * g.nv;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long off = (t % g.nu) * g.stride_u + (t / g.nu) * g.stride_v;
        double v = f[g.base_b + off];
        if (dudn != nullptr) {
            const double corr = __ddiv_rn(__dmul_rn(batch.two_ds[q], dudn[t]), ctl->norm_b);
            v = batch.upper[q] ? __dadd_rn(v, corr) : __dsub_rn(v, corr);
        }
        f[g.base_a + off] = v;
    }
}

// adjustFieldBForDirichletNeumanBCs (iterativeSolverBase.hpp:429-534), one face:
//   Dirichlet: b(A = first interior plane) -= x(B = boundary plane) / ds^2        (:454, :502)
//   Neumann:   b(A = boundary plane)      +/-= nfac dudn / ds                      (:475-480, :522-527)
//              nfac = 2 for orderNeumanBcs == 2, 1 for orderNeumanBcs == 1 (1 * dudn is exact, so both are the reference's bits)
static __global__ void adjust_b_kernel(double* __restrict__ b, const double* __restrict__ x, FaceGeom g,
                                const double* __restrict__ dudn, double ds, int neumann, int upper, double nfac) {
    const long long n = static_cast<long long>(g.nu) * g.nv;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long off = (t % g.nu) * g.stride_u + (t / g.nu) * g.stride_v;
        if (!neumann) {
            b[g.base_a + off] = __dsub_rn(b[g.base_a + off], __ddiv_rn(x[g.base_b + off], __dmul_rn(ds, ds)));
        } else {
            const double c = __ddiv_rn(__dmul_rn(nfac, dudn[t]), ds);
            b[g.base_a + off] = upper ? __dsub_rn(b[g.base_a + off], c) : __dadd_rn(b[g.base_a + off], c);
        }
    }
}

// halo between two blocks that live on the same GPU: guard plane (A of dst) <- boundary data plane (B of src)
// (what CommunicatorMPI::operator() moves per face, communicationMPI.hpp:51-292)
static __global__ void face_copy_kernel(double* __restrict__ dst, const double* __restrict__ src, FaceGeom g, const Ctl* ctl,
                                 int ignore_done) {
    if (!ignore_done && ctl != nullptr && ctl->done) return;
    const long long n = static_cast<long long>(g.nu) * g.nv;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long off = (t % g.nu) * g.stride_u + (t / g.nu) * g.stride_v;
        dst[g.base_a + off] = src[g.base_b + off];
    }
}

// peer-memory halo path: await the epoch a neighbour's copy engine writes into my flag after its plane has landed
static __global__ void await_epoch_sys_kernel(const unsigned int* flag, unsigned int epoch) {
    unsigned int v;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (static_cast<int>(v - epoch) >= 0) break;
        __nanosleep(100);
    } while (true);
}

// pack / unpack a face into / from a contiguous buffer (NCCL send/recv of x- and y-faces)
static __global__ void face_pack_kernel(double* __restrict__ buf, const double* __restrict__ f, FaceGeom g, const Ctl* ctl,
                                 int ignore_done) {
    if (!ignore_done && ctl != nullptr && ctl->done) return;
    const long long n = static_cast<long long>(g.nu) * g.nv;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n;
         t += static_cast<long long>(gridDim.x) * blockDim.x)
        buf[t] = f[g.base_b + (t % g.nu) * g.stride_u + (t / g.nu) * g.stride_v];
}
static __global__ void face_unpack_kernel(double* __restrict__ f, const double* __restrict__ buf, FaceGeom g, const Ctl* ctl,
                                   int ignore_done) {
    if (!ignore_done && ctl != nullptr && ctl->done) return;
    const long long n = static_cast<long long>(g.nu) * g.nv;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n;
         t += static_cast<long long>(gridDim.x) * blockDim.x)
        f[g.base_a + (t % g.nu) * g.stride_u + (t / g.nu) * g.stride_v] = buf[t];
}

}  // namespace pps

// cheb_blocked.cuh -- temporally blocked Chebyshev sweeps (chebyshevIteration.hpp:93-116 with communicationOFF; the
// alpaka tree's shared-memory variants are kernelsAlpakaChebyshev.hpp:192-273).
//
// The per-sweep schedule (solver.cu chebyshev_local) moves 4 vectors per sweep through HBM: 3 + 4 (n - 3) passes = 280 B per
// cell at chebyshevMax = 11, plus one ghost launch per Neumann face per sweep.  Here ONE pass over the block advances the
// recurrence by NLEV sweeps: a CTA owns a 64 x FY footprint of columns and marches in z; sweep l of the pass trails sweep
// l - 1 by one plane, so at march step s level l computes plane s - l from
//     z-neighbours  the thread's own register window of level l - 1 (planes s-l-1, s-l, s-l+1)
//     x-neighbours  warp shuffles (a warp owns one row, one value pair per lane)
//     y-neighbours  the shared-memory copy of plane s - l of level l - 1, written one step earlier (double buffered:
//                   ONE __syncthreads per plane for all levels)
//     B, y_{c-2}    registers (B travels down a shift register, y_{c-2} is the oldest entry of level l - 2's window)
// The footprint shrinks by one cell per level (halo of NLEV cells, recomputed by the neighbouring CTAs), z-chunks overlap by
// NLEV planes on each side.  HBM traffic per pass: read Y, Z, B, write Y', Z' (5 vectors) for NLEV sweeps instead of 4 NLEV.
//
// Neumann faces are folded in as index remaps: a boundary cell reads, in place of the ghost, the value the ghost would hold
// (its opposite neighbour for orderNeumanBcs = 2, itself for 1; iterativeSolverBase.hpp:92-108) -- the ghost planes of the
// intermediate iterates are never materialised and no neumann_ghost launch is needed.  Cells outside the solver box hold 0
// in every iterate (Dirichlet planes, inter-rank guards of the block-Jacobi preconditioner), as in the per-sweep schedule.
//
// Arithmetic is a policy (`Form`): ChebFormCpu<PARITY> evaluates exactly the expressions of EpiChebFirst / EpiChebStep
// (kernels.cuh), so the blocked schedule is bit-identical to the per-sweep one in both arithmetic modes;
// ChebFormAlpaka<T, PARITY> is the folded 7-point form of the alpaka tree (kernelsAlpakaChebyshev.hpp:136-183,233-270) in
// T = float or double -- the mixed-precision preconditioner (T_data_chebyshev = float, solverSetup.hpp:14 of the alpaka tree).
#pragma once

#include "kernels.cuh"

namespace pps {

constexpr int kChebMaxLev = 4;
#ifndef PPS_CHEB_OCC2_MAXLEV
#define PPS_CHEB_OCC2_MAXLEV 0   // fp64: depths up to this value are compiled for two CTAs per SM (64 registers spill the fp64 windows: 0 = none)
#endif

template <typename T> struct ChebVec;
template <> struct ChebVec<double> { using type = double2; };
template <> struct ChebVec<float> { using type = float2; };

template <typename T>
__device__ __forceinline__ typename ChebVec<T>::type cheb_make2(T a, T b) {
    typename ChebVec<T>::type v;
    v.x = a; v.y = b;
    return v;
}

// geometry of one blocked pass
struct ChebTile {
    int le;            // halo cells of the footprint in x on each side: NLEV rounded up to even (keeps the value pairs 16-byte aligned)
    int wx, wy;        // output tile = 64 - 2 le columns, FY - 2 NLEV rows
    int zchunk;        // output planes per CTA
    int nm[6];         // Neumann remap per face: 0 none, 2 mirror (orderNeumanBcs = 2), 1 boundary value (orderNeumanBcs = 1)
};

// ------------------------------------------------------------------------------------------------
// arithmetic policies
// ------------------------------------------------------------------------------------------------
// the CPU reference's expressions (chebyshevIteration.hpp:79-90,103-113): same device code as EpiChebFirst::y / EpiChebStep::w
template <bool PARITY>
struct ChebFormCpu {
    using T = double;
    Coef cf;
    double theta, inv_theta, c1, two_sigma, two_over_delta;
    double rho[kChebMaxLev + 1], rho_old[kChebMaxLev + 1];   // per level of this pass (index 1..NLEV)
    __device__ __forceinline__ double cast_in(double b) const { return b; }
    __device__ __forceinline__ double y0(double b) const { return PARITY ? __ddiv_rn(b, theta) : b * inv_theta; }
    __device__ __forceinline__ double first(double b, double xm, double xp, double ym, double yp, double zm, double zp) const {
        const double ab = laplacian<PARITY>(cf, xm, b, xp, ym, yp, zm, zp);
        if (PARITY) return __dmul_rn(c1, __dadd_rn(__dmul_rn(2.0, b), __ddiv_rn(ab, theta)));
        return c1 * fma(ab, inv_theta, 2.0 * b);
    }
    __device__ __forceinline__ double step(int lev, double y, double xm, double xp, double ym, double yp, double zm, double zp,
                                           double b, double z) const {
        const double ay = laplacian<PARITY>(cf, xm, y, xp, ym, yp, zm, zp);
        const double r = rho[lev], ro = rho_old[lev];
        if (PARITY)
            return __dmul_rn(r, __dsub_rn(__dadd_rn(__dmul_rn(two_sigma, y), __dmul_rn(two_over_delta, __dadd_rn(b, ay))), __dmul_rn(ro, z)));
        return r * fma(-ro, z, fma(two_sigma, y, two_over_delta * (b + ay)));
    }
    __device__ __forceinline__ double out_x(double w) const { return -w; }
};

template <typename T> __device__ __forceinline__ T rn_mul(T a, T b);
template <> __device__ __forceinline__ double rn_mul<double>(double a, double b) { return __dmul_rn(a, b); }
template <> __device__ __forceinline__ float rn_mul<float>(float a, float b) { return __fmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T rn_add(T a, T b);
template <> __device__ __forceinline__ double rn_add<double>(double a, double b) { return __dadd_rn(a, b); }
template <> __device__ __forceinline__ float rn_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <typename T> __device__ __forceinline__ T rn_div(T a, T b);
template <> __device__ __forceinline__ double rn_div<double>(double a, double b) { return __ddiv_rn(a, b); }
template <> __device__ __forceinline__ float rn_div<float>(float a, float b) { return __fdiv_rn(a, b); }

// the alpaka tree's folded form in T_data_chebyshev = T (kernelsAlpakaChebyshev.hpp):
//   first  tmp = b fc0 + (xm + xp) f0;  tmp += (ym + yp) f1 + (zm + zp) f2                                   :158-162
//   step   tmp = y fc0 + (xm + xp) f0;  tmp += (ym + yp) f1 + (zm + zp) f2;  tmp += b fB;  tmp += z fZ        :255-270
// coefficients are evaluated on the host in T with the kernels' own expressions (solver.cu cheb_alpaka_coefficients).
// PARITY: one rounding per operation in this order (what a non-contracting build of the alpaka kernels computes; the oracle's
// restatement); FAST: the compiler may contract into FMAs.
template <typename TT, bool PARITY>
struct ChebFormAlpaka {
    using T = TT;
    T theta;
    T fc0[kChebMaxLev + 1], f0[kChebMaxLev + 1], f1[kChebMaxLev + 1], f2[kChebMaxLev + 1], fB[kChebMaxLev + 1], fZ[kChebMaxLev + 1];
    T fc0_first, f0_first, f1_first, f2_first;
    __device__ __forceinline__ T cast_in(double b) const { return static_cast<T>(b); }                        // CastPrecisionFieldKernel :8-18
    __device__ __forceinline__ T y0(T b) const { return PARITY ? rn_div<T>(b, theta) : b / theta; }           // :150
    __device__ __forceinline__ T stencil(T c, T xm, T xp, T ym, T yp, T zm, T zp, T kc, T k0, T k1, T k2) const {
        if (PARITY) {
            T tmp = rn_add<T>(rn_mul<T>(c, kc), rn_mul<T>(rn_add<T>(xm, xp), k0));
            return rn_add<T>(tmp, rn_add<T>(rn_mul<T>(rn_add<T>(ym, yp), k1), rn_mul<T>(rn_add<T>(zm, zp), k2)));
        }
        T tmp = c * kc + (xm + xp) * k0;
        tmp += (ym + yp) * k1 + (zm + zp) * k2;
        return tmp;
    }
    __device__ __forceinline__ T first(T b, T xm, T xp, T ym, T yp, T zm, T zp) const {
        return stencil(b, xm, xp, ym, yp, zm, zp, fc0_first, f0_first, f1_first, f2_first);
    }
    __device__ __forceinline__ T step(int lev, T y, T xm, T xp, T ym, T yp, T zm, T zp, T b, T z) const {
        T tmp = stencil(y, xm, xp, ym, yp, zm, zp, fc0[lev], f0[lev], f1[lev], f2[lev]);
        if (PARITY) return rn_add<T>(rn_add<T>(tmp, rn_mul<T>(b, fB[lev])), rn_mul<T>(z, fZ[lev]));
        tmp += b * fB[lev];
        tmp += z * fZ[lev];
        return tmp;
    }
    // AssignFieldWith1FieldKernel with constB = -1 in T, result widened to the main solver's double (:24-45)
    __device__ __forceinline__ double out_x(T w) const { return static_cast<double>(static_cast<T>(-1.0) * w); }
};

// arrays of one pass.  Iterates live in T (fp64 or fp32) arrays with the pitched layout of common.cuh (element counts, not bytes)
template <typename T>
struct ChebIO {
    const double* B;   // right-hand side of the preconditioner call (fp64)
    const T* Yin;      // y_{c0}, read with its halo      (unused in the first pass: level 0 is B itself)
    const T* Zin;      // y_{c0-1}, centre only
    T* Yout;           // y_{c0+NLEV}                      (unused in the last pass)
    T* Zout;           // y_{c0+NLEV-1}
    double* X;         // last pass: X = -y_last (chebyshevIteration.hpp:118-128)
};

// ------------------------------------------------------------------------------------------------
// the kernel: grid = x tiles * y tiles * z chunks, block = 32 x FY
//
// The kernel is instruction-issue bound, not bandwidth bound (ncu / SASS of the first version: ~700 instructions per march
// step for 84 fp64 operations), so the march is written for a short instruction stream:
//   * the three-plane register windows ROTATE instead of shifting: the march is unrolled by three and step R writes the new
//     plane into slot R (new = w[R], mid = w[R+2 mod 3], old = w[R+1 mod 3]) -- no register moves;
//   * loads are two predicated 8-byte loads per field (no divergent branches, no zero-filling), the plane test is CTA-uniform;
//   * the Neumann remaps in x / y exist only in the instantiation that boundary CTAs take (CTA-uniform branch); the z remap is
//     a step-uniform branch.
// ------------------------------------------------------------------------------------------------
template <int N> struct ChebInt { static constexpr int value = N; };

// two CTAs per SM (<= 64 registers) where the register windows allow it: the kernel is latency bound at 16 warps per SM
template <int NLEV, class Form>
struct ChebOcc { static constexpr int kMinBlocks = (sizeof(typename Form::T) == 4 || NLEV <= PPS_CHEB_OCC2_MAXLEV) ? 2 : 1; };

template <int NLEV, int FY, bool FIRST, bool LAST, class Form>
__global__ void __launch_bounds__(32 * FY, ChebOcc<NLEV, Form>::kMinBlocks) cheb_blocked_kernel(Dims d, Box rg, ChebTile tl, Form fm, ChebIO<typename Form::T> io,
                                                              const Ctl* ctl) {
    if (ctl != nullptr && ctl->done) return;
    using T = typename Form::T;
    using V = typename ChebVec<T>::type;
    static_assert(NLEV >= 1 && NLEV <= kChebMaxLev, "levels per pass");
    extern __shared__ __align__(16) unsigned char cheb_smem_raw[];
    T (*plane)[2][FY][64] = reinterpret_cast<T (*)[2][FY][64]>(cheb_smem_raw);   // [NLEV][2][FY][64]

    const int lane = threadIdx.x, row = threadIdx.y;
    const int ox = 1 + blockIdx.x * tl.wx, oy = 1 + blockIdx.y * tl.wy;   // first output cell of this tile
    const int i0 = ox - tl.le + 2 * lane;                                  // my two columns: i0, i0 + 1
    const int j = oy - NLEV + row;
    const int ka = rg.k0 + blockIdx.z * tl.zchunk, kb = min(rg.k1, ka + tl.zchunk);   // output planes [ka, kb)
    const bool jr = j >= rg.j0 && j < rg.j1;
    const bool a0 = jr && i0 >= rg.i0 && i0 < rg.i1;          // inside the solver box (xy)
    const bool a1 = jr && i0 + 1 >= rg.i0 && i0 + 1 < rg.i1;
    const bool jo = j >= oy && j < oy + tl.wy;                // inside the output tile
    const bool o0 = a0 && jo && i0 >= ox && i0 < ox + tl.wx;
    const bool o1 = a1 && jo && i0 + 1 >= ox && i0 + 1 < ox + tl.wx;
    // offset of (i0, j) inside a plane; only dereferenced where a0 / a1 hold, i.e. inside the array
    const long long rowoff = static_cast<long long>(kOff) + i0 + d.pitch * static_cast<long long>(max(0, min(j, d.ny + 1)));
    const int rm = max(row - 1, 0), rp = min(row + 1, FY - 1);
    const T* __restrict__ pF = FIRST ? nullptr : io.Yin + rowoff;
    const T* __restrict__ pZ = FIRST ? nullptr : io.Zin + rowoff;
    const double* __restrict__ pB = io.B + rowoff;

    const V zero = cheb_make2<T>(T(0), T(0));
    V w[NLEV][3];        // level l: rotating window of three planes
    V bq[NLEV + 1];      // B at planes s, s-1, .., s-NLEV (cast to T)
    V zq = zero;         // Zin at plane s-1 (not FIRST)
#pragma unroll
    for (int l = 0; l < NLEV; l++) { w[l][0] = zero; w[l][1] = zero; w[l][2] = zero; }
#pragma unroll
    for (int m = 0; m <= NLEV; m++) bq[m] = zero;

    // predicated loads: no branch on the thread's own activity; the plane test is uniform over the CTA
    auto load_T = [&](const T* p, int k) -> V {
        V v = zero;
        if (k >= rg.k0 && k < rg.k1) {
            const T* q = p + k * d.plane;
            if (a0) v.x = __ldg(q);
            if (a1) v.y = __ldg(q + 1);
        }
        return v;
    };
    auto load_B = [&](int k) -> V {
        V v = zero;
        if (k >= rg.k0 && k < rg.k1) {
            const double* q = pB + k * d.plane;
            if (a0) v.x = fm.cast_in(__ldg(q));
            if (a1) v.y = fm.cast_in(__ldg(q + 1));
        }
        return v;
    };

    const int s0 = ka - NLEV, s1 = kb - 1 + NLEV;   // march steps: level 0 touches planes s0 .. s1, level NLEV planes ka .. kb-1
    // two planes of prefetch for everything that is read from HBM
    V pf_f[2], pf_b[2] = {zero, zero}, pf_z[2] = {zero, zero};
    pf_f[0] = FIRST ? load_B(s0) : load_T(pF, s0);
    pf_f[1] = FIRST ? load_B(s0 + 1) : load_T(pF, s0 + 1);
    if (!FIRST) {
        pf_b[0] = load_B(s0); pf_b[1] = load_B(s0 + 1);
        pf_z[0] = load_T(pZ, s0); pf_z[1] = load_T(pZ, s0 + 1);
    }

    // one march step with window rotation R (compile time) and, for boundary CTAs, the x / y Neumann remaps
    auto step = [&](auto Rc, auto REMAPc, int s, int xl0, int xh0, int xl1, int xh1, int yl, int yh) {
        constexpr int R = decltype(Rc)::value, RM = (R + 2) % 3, RO = (R + 1) % 3;   // new, mid, old slots
        constexpr bool REMAP = decltype(REMAPc)::value != 0;
        const int cur = s & 1, prv = cur ^ 1;
        // ---- level 0: the plane that arrives from HBM
        const V f0v = pf_f[0];
        const V zprev = zq;                  // Zin at plane s-1
        pf_f[0] = pf_f[1];
        pf_f[1] = FIRST ? load_B(s + 2) : load_T(pF, s + 2);
#pragma unroll
        for (int m = NLEV; m > 0; m--) bq[m] = bq[m - 1];
        if (FIRST) {
            bq[0] = f0v;
        } else {
            bq[0] = pf_b[0];
            zq = pf_z[0];
            pf_b[0] = pf_b[1]; pf_z[0] = pf_z[1];
            pf_b[1] = load_B(s + 2);
            pf_z[1] = load_T(pZ, s + 2);
        }
        w[0][R] = f0v;
        *reinterpret_cast<V*>(&plane[0][cur][row][2 * lane]) = f0v;

        // ---- levels 1 .. NLEV, each one plane behind the previous
#pragma unroll
        for (int l = 1; l <= NLEV; l++) {
            const int k = s - l;
            const V c = w[l - 1][RM], zmv = w[l - 1][RO], zpv = w[l - 1][R];
            const V ymv = *reinterpret_cast<const V*>(&plane[l - 1][prv][rm][2 * lane]);
            const V ypv = *reinterpret_cast<const V*>(&plane[l - 1][prv][rp][2 * lane]);
            const T xl = __shfl_up_sync(kFullMask, c.y, 1);     // lane 0 / 31: a halo column, any value will do
            const T xr = __shfl_down_sync(kFullMask, c.x, 1);
            T xm0 = xl, xp0 = c.y, xm1 = c.x, xp1 = xr;
            T ym0 = ymv.x, yp0 = ypv.x, ym1 = ymv.y, yp1 = ypv.y;
            T zm0 = zmv.x, zp0 = zpv.x, zm1 = zmv.y, zp1 = zpv.y;
            if (REMAP) {
                // ghost value = opposite neighbour (orderNeumanBcs = 2) or the cell itself (= 1)
                if (xl0) xm0 = (xl0 == 2) ? xp0 : c.x;
                if (xh0) xp0 = (xh0 == 2) ? xm0 : c.x;
                if (xl1) xm1 = (xl1 == 2) ? xp1 : c.y;
                if (xh1) xp1 = (xh1 == 2) ? xm1 : c.y;
                if (yl) { ym0 = (yl == 2) ? yp0 : c.x; ym1 = (yl == 2) ? yp1 : c.y; }
                if (yh) { yp0 = (yh == 2) ? ym0 : c.x; yp1 = (yh == 2) ? ym1 : c.y; }
            }
            V z = zero;
            if (!(FIRST && l == 1)) {
                if (l == 1) z = zprev;                                                                       // y_{c0-1} from HBM
                else if (FIRST && l == 2) z = cheb_make2<T>(fm.y0(w[0][RO].x), fm.y0(w[0][RO].y));           // y_0 = B / theta
                else z = w[l >= 2 ? l - 2 : 0][RO];
            }
            auto eval = [&](T a_zm0, T a_zp0, T a_zm1, T a_zp1) -> V {
                V r;
                if (FIRST && l == 1) {
                    r.x = fm.first(c.x, xm0, xp0, ym0, yp0, a_zm0, a_zp0);
                    r.y = fm.first(c.y, xm1, xp1, ym1, yp1, a_zm1, a_zp1);
                } else {
                    r.x = fm.step(l, c.x, xm0, xp0, ym0, yp0, a_zm0, a_zp0, bq[l].x, z.x);
                    r.y = fm.step(l, c.y, xm1, xp1, ym1, yp1, a_zm1, a_zp1, bq[l].y, z.y);
                }
                return r;
            };
            V v;
            const int zl = (k == 1) ? tl.nm[4] : 0, zh = (k == d.nz) ? tl.nm[5] : 0;   // uniform over the CTA, non-zero on two planes only
            if (zl | zh) {
                // rare path, evaluated on its own so that the common path carries no merges
                if (zl) { zm0 = (zl == 2) ? zp0 : c.x; zm1 = (zl == 2) ? zp1 : c.y; }
                if (zh) { zp0 = (zh == 2) ? zm0 : c.x; zp1 = (zh == 2) ? zm1 : c.y; }
                v = eval(zm0, zp0, zm1, zp1);
            } else {
                v = eval(zmv.x, zpv.x, zmv.y, zpv.y);
            }
            const bool kin = k >= rg.k0 && k < rg.k1;
            v.x = (kin && a0) ? v.x : T(0);
            v.y = (kin && a1) ? v.y : T(0);
            if (l < NLEV) {
                w[l < NLEV ? l : 0][R] = v;
                *reinterpret_cast<V*>(&plane[l < NLEV ? l : 0][cur][row][2 * lane]) = v;
            } else if (k >= ka && k < kb && (o0 || o1)) {
                const long long idx = rowoff + k * d.plane;
                if (LAST) {
                    st2(io.X + idx, make_double2(fm.out_x(v.x), fm.out_x(v.y)), o0, o1);
                } else {
                    // y_{c0+NLEV} and y_{c0+NLEV-1} at plane k
                    const V zo = (FIRST && NLEV == 1) ? cheb_make2<T>(fm.y0(w[0][RM].x), fm.y0(w[0][RM].y)) : w[NLEV - 1][RM];
                    if (o0) { io.Yout[idx] = v.x; io.Zout[idx] = zo.x; }
                    if (o1) { io.Yout[idx + 1] = v.y; io.Zout[idx + 1] = zo.y; }
                }
            }
        }
        __syncthreads();
    };

    // does this CTA's footprint touch a Neumann face in x or y?  (uniform)
    const int fx0 = ox - tl.le, fx1 = fx0 + 63, fy0 = oy - NLEV, fy1 = fy0 + FY - 1;
    const bool remap = (tl.nm[0] && fx0 <= 1 && 1 <= fx1) || (tl.nm[1] && fx0 <= d.nx && d.nx <= fx1) ||
                       (tl.nm[2] && fy0 <= 1 && 1 <= fy1) || (tl.nm[3] && fy0 <= d.ny && d.ny <= fy1);
    if (remap) {
        const int xl0 = (i0 == 1) ? tl.nm[0] : 0, xh0 = (i0 == d.nx) ? tl.nm[1] : 0;
        const int xl1 = (i0 + 1 == 1) ? tl.nm[0] : 0, xh1 = (i0 + 1 == d.nx) ? tl.nm[1] : 0;
        const int yl = (j == 1) ? tl.nm[2] : 0, yh = (j == d.ny) ? tl.nm[3] : 0;
        for (int s = s0; s <= s1; s += 3) {
            step(ChebInt<0>{}, ChebInt<1>{}, s, xl0, xh0, xl1, xh1, yl, yh);
            step(ChebInt<1>{}, ChebInt<1>{}, s + 1, xl0, xh0, xl1, xh1, yl, yh);
            step(ChebInt<2>{}, ChebInt<1>{}, s + 2, xl0, xh0, xl1, xh1, yl, yh);
        }
    } else {
        for (int s = s0; s <= s1; s += 3) {
            step(ChebInt<0>{}, ChebInt<0>{}, s, 0, 0, 0, 0, 0, 0);
            step(ChebInt<1>{}, ChebInt<0>{}, s + 1, 0, 0, 0, 0, 0, 0);
            step(ChebInt<2>{}, ChebInt<0>{}, s + 2, 0, 0, 0, 0, 0, 0);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host interface (defined in cheb_blocked.cu, its own translation unit: 4 forms x 4 depths x first/last instantiations)
// ------------------------------------------------------------------------------------------------
enum ChebFormKind : int { CHEB_FORM_CPU_FAST = 0, CHEB_FORM_CPU_PARITY = 1, CHEB_FORM_ALPAKA_F32_FAST = 2, CHEB_FORM_ALPAKA_F32_PARITY = 3 };
constexpr int kChebFY = 16;   // footprint rows = warps per CTA

struct ChebPassDesc {
    int nlev;          // sweeps advanced by this pass (1 .. kChebMaxLev)
    int first, last;   // first pass of a call (level 0 is B) / last pass (writes X = -y)
    int form;          // ChebFormKind
    // CPU form
    Coef cf;
    double theta, c1, two_sigma, two_over_delta;
    double rho[kChebMaxLev + 1], rho_old[kChebMaxLev + 1];
    // alpaka form (fp32): coefficients of kernelsAlpakaChebyshev.hpp:148-151,240-246 evaluated on the host in float
    float a_theta, a_fc0[kChebMaxLev + 1], a_f0[kChebMaxLev + 1], a_f1[kChebMaxLev + 1], a_f2[kChebMaxLev + 1], a_fB[kChebMaxLev + 1],
        a_fZ[kChebMaxLev + 1];
    float a_fc0_first, a_f0_first, a_f1_first, a_f2_first;
    // arrays (iterates are double or float depending on the form)
    const double* B;
    const void* Yin;
    const void* Zin;
    void* Yout;
    void* Zout;
    double* X;
};

// enqueue one pass; returns the number of CTAs launched
unsigned int cheb_blocked_launch(cudaStream_t stream, const Dims& d, const Box& box, const ChebTile& tl, const ChebPassDesc& p, const Ctl* ctl);
// output tile of a pass with `nlev` levels
ChebTile cheb_make_tile(int nlev, int zchunk, const int nm[6]);

}  // namespace pps

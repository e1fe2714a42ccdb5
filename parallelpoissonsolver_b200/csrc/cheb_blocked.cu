// cheb_blocked.cu -- instantiations and launcher of the temporally blocked Chebyshev kernel (cheb_blocked.cuh)
#include "cheb_blocked.cuh"

#include <mutex>
#include <set>

namespace pps {

ChebTile cheb_make_tile(int nlev, int zchunk, const int nm[6]) {
    ChebTile t;
    t.le = nlev + (nlev & 1);
    t.wx = 64 - 2 * t.le;
    t.wy = kChebFY - 2 * nlev;
    t.zchunk = zchunk;
    for (int f = 0; f < 6; f++) t.nm[f] = nm[f];
    return t;
}

template <bool PARITY>
static ChebFormCpu<PARITY> make_cpu_form(const ChebPassDesc& p) {
    ChebFormCpu<PARITY> f;
    f.cf = p.cf;
    f.theta = p.theta; f.inv_theta = 1.0 / p.theta; f.c1 = p.c1; f.two_sigma = p.two_sigma; f.two_over_delta = p.two_over_delta;
    for (int l = 0; l <= kChebMaxLev; l++) { f.rho[l] = p.rho[l]; f.rho_old[l] = p.rho_old[l]; }
    return f;
}
template <bool PARITY>
static ChebFormAlpaka<float, PARITY> make_alpaka_form(const ChebPassDesc& p) {
    ChebFormAlpaka<float, PARITY> f;
    f.theta = p.a_theta;
    for (int l = 0; l <= kChebMaxLev; l++) {
        f.fc0[l] = p.a_fc0[l]; f.f0[l] = p.a_f0[l]; f.f1[l] = p.a_f1[l]; f.f2[l] = p.a_f2[l]; f.fB[l] = p.a_fB[l]; f.fZ[l] = p.a_fZ[l];
    }
    f.fc0_first = p.a_fc0_first; f.f0_first = p.a_f0_first; f.f1_first = p.a_f1_first; f.f2_first = p.a_f2_first;
    return f;
}

template <int NLEV, bool FIRST, bool LAST, class Form>
static unsigned int launch_inst(cudaStream_t stream, const Dims& d, const Box& box, const ChebTile& tl, const Form& fm, const ChebPassDesc& p,
                                const Ctl* ctl) {
    using T = typename Form::T;
    auto kern = cheb_blocked_kernel<NLEV, kChebFY, FIRST, LAST, Form>;
    constexpr int smem = NLEV * 2 * kChebFY * 64 * static_cast<int>(sizeof(T));
    if (smem > 48 * 1024) {
        // per-device opt-in above 48 KB; cheap enough to repeat (the attribute is sticky per function and device)
        static std::mutex mu;
        static std::set<std::pair<const void*, int>> done;
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        if (done.insert({reinterpret_cast<const void*>(kern), dev}).second)
            PPS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    ChebIO<T> io{p.B, static_cast<const T*>(p.Yin), static_cast<const T*>(p.Zin), static_cast<T*>(p.Yout), static_cast<T*>(p.Zout), p.X};
    const int nx = d.nx, ny = d.ny;
    dim3 grid((nx + tl.wx - 1) / tl.wx, (ny + tl.wy - 1) / tl.wy, (std::max(1, box.k1 - box.k0) + tl.zchunk - 1) / tl.zchunk);
    kern<<<grid, dim3(32, kChebFY, 1), smem, stream>>>(d, box, tl, fm, io, ctl);
    return grid.x * grid.y * grid.z;
}

template <int NLEV, class Form>
static unsigned int launch_flags(cudaStream_t stream, const Dims& d, const Box& box, const ChebTile& tl, const Form& fm, const ChebPassDesc& p,
                                 const Ctl* ctl) {
    if (p.first && p.last) return launch_inst<NLEV, true, true>(stream, d, box, tl, fm, p, ctl);
    if (p.first) return launch_inst<NLEV, true, false>(stream, d, box, tl, fm, p, ctl);
    if (p.last) return launch_inst<NLEV, false, true>(stream, d, box, tl, fm, p, ctl);
    return launch_inst<NLEV, false, false>(stream, d, box, tl, fm, p, ctl);
}

template <class Form>
static unsigned int launch_depth(cudaStream_t stream, const Dims& d, const Box& box, const ChebTile& tl, const Form& fm, const ChebPassDesc& p,
                                 const Ctl* ctl) {
    switch (p.nlev) {
        case 1: return launch_flags<1>(stream, d, box, tl, fm, p, ctl);
        case 2: return launch_flags<2>(stream, d, box, tl, fm, p, ctl);
        case 3: return launch_flags<3>(stream, d, box, tl, fm, p, ctl);
        case 4: return launch_flags<4>(stream, d, box, tl, fm, p, ctl);
        default: throw std::runtime_error("cheb_blocked_launch: 1 .. 4 sweeps per pass");
    }
}

unsigned int cheb_blocked_launch(cudaStream_t stream, const Dims& d, const Box& box, const ChebTile& tl, const ChebPassDesc& p, const Ctl* ctl) {
    switch (p.form) {
        case CHEB_FORM_CPU_FAST: return launch_depth(stream, d, box, tl, make_cpu_form<false>(p), p, ctl);
        case CHEB_FORM_CPU_PARITY: return launch_depth(stream, d, box, tl, make_cpu_form<true>(p), p, ctl);
        case CHEB_FORM_ALPAKA_F32_FAST: return launch_depth(stream, d, box, tl, make_alpaka_form<false>(p), p, ctl);
        case CHEB_FORM_ALPAKA_F32_PARITY: return launch_depth(stream, d, box, tl, make_alpaka_form<true>(p), p, ctl);
        default: throw std::runtime_error("cheb_blocked_launch: unknown arithmetic form");
    }
}

}  // namespace pps

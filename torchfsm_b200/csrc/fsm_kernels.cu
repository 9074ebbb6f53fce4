// Kernel instantiations for ONE line length: compile with -DFSM_N=<N> (8..1024).
#include "fsm_launch.h"
#include <cerrno>
#include <mutex>
#include <vector>

#ifndef FSM_N
#error "compile with -DFSM_N=<line length>"
#endif

namespace fsm {

// FFT decomposition per line length: N = R0*R1*R2, EPT elements per thread, TL = N/EPT threads per line.
template <int N> struct CfgFor;
template <> struct CfgFor<8> { using type = FftCfg<8, 4, 4, 2>; };
template <> struct CfgFor<16> { using type = FftCfg<16, 4, 4, 4>; };
template <> struct CfgFor<32> { using type = FftCfg<32, 8, 8, 4>; };
template <> struct CfgFor<64> { using type = FftCfg<64, 8, 8, 8>; };
template <> struct CfgFor<128> { using type = FftCfg<128, 16, 16, 8>; };
template <> struct CfgFor<256> { using type = FftCfg<256, 16, 16, 16>; };
#ifndef FSM_C512_R1
#define FSM_C512_R1 8    // measured on C5: (16,8,4) 38.1 ms/step, (16,16,2) 39.0
#define FSM_C512_R2 4
#endif
template <> struct CfgFor<512> { using type = FftCfg<512, 16, 16, FSM_C512_R1, FSM_C512_R2>; };
#ifndef FSM_C1024_R0   // tuning hooks: -DFSM_C1024_R0=.. -DFSM_C1024_R1=.. -DFSM_C1024_R2=.. (EPT stays 16)
#define FSM_C1024_R0 16   // measured on C3: (16,8,8) 2.10 ms/step, (16,16,4) 2.17, (8,16,8) 2.19, (8,8,16) 2.23, (4,16,16) 2.28
#define FSM_C1024_R1 8
#define FSM_C1024_R2 8
#endif
template <> struct CfgFor<1024> { using type = FftCfg<1024, 16, FSM_C1024_R0, FSM_C1024_R1, FSM_C1024_R2>; };

// Pass-specific decompositions. The 3-D last-axis pass carries three accumulators of u.grad(u) plus the
// working line, and the 3-channel FX pass three spectra (NS pressure projection couples the channels):
// with 8 elements per thread they stay in registers (measured: C4 59 -> 34 ms/step, no spills).
template <int N, int NDIM> struct CfgPhys { using type = typename CfgFor<N>::type; };
template <> struct CfgPhys<128, 3> { using type = FftCfg<128, 8, 8, 8, 2>; };
template <> struct CfgPhys<256, 3> { using type = FftCfg<256, 8, 8, 8, 4>; };
template <> struct CfgPhys<512, 3> { using type = FftCfg<512, 8, 8, 8, 8>; };
template <int N, int C> struct CfgFx { using type = typename CfgFor<N>::type; };
template <> struct CfgFx<128, 3> { using type = FftCfg<128, 8, 8, 8, 2>; };
template <> struct CfgFx<256, 3> { using type = FftCfg<256, 8, 8, 8, 4>; };
template <> struct CfgFx<512, 3> { using type = FftCfg<512, 8, 8, 8, 8>; };

#ifndef FSM_EMU
template <class K>
static inline int set_smem(K kern, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return -EIO;
    }
    return 0;
}
static inline int check_launch() { return cudaGetLastError() == cudaSuccess ? 0 : -EIO; }
// CTAs of this kernel resident on the whole GPU at once = the L2 prefetch distance (Geom::pf_wave)
template <class K>
static inline int resident_ctas(K kern, int block, size_t smem) {
    int per_sm = 0, dev = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem) != cudaSuccess) return 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return per_sm * sms;
}
#else
template <class K> static inline int resident_ctas(K, int, size_t) { return 0; }
template <class K> static inline int set_smem(K, size_t) { return 0; }
static inline int check_launch() { return 0; }
#endif

template <typename T, int N, int PROG>
static int launch_ix_p(const IxArgs<T>& a, cudaStream_t s) {
    using Cfg = typename CfgFor<N>::type;
    auto kern = k_pass_ix<T, Cfg, PROG>;
    const size_t smem = Smem<Cfg, T>::bytes(2 * kKL);
    if (int e = set_smem(kern, smem)) return e;
    dim3 grid((a.n_t + kKL - 1) / kKL, a.n_outer, a.nbc), block(kKL * Cfg::TL);
    static const int wave = resident_ctas(kern, kKL * Cfg::TL, smem);
    Geom<T> g = a.g;
    g.pf_wave = wave;
    FSM_LAUNCH(kern, grid, block, smem, s, g, a.state, a.w1, a.state_bstride, a.w1_fstride, kKL, a.in_t_stride,
               a.in_o_stride, a.out_o_stride, a.out_e_stride, a.n_t, a.eb, a.pe);
    return check_launch();
}
#ifndef FSM_IXZ
#define FSM_IXZ 1   // Z-line programs on one GPU run the phase-overlapping inverse-x kernel (k_pass_ixz)
#endif
template <typename T, int N, int PROG>
static int launch_ixz_p(const IxArgs<T>& a, cudaStream_t s) {
    using Cfg = typename CfgFor<N>::type;
    using Z = PhysZ<Cfg>;
    auto kern = k_pass_ixz<T, Cfg, PROG>;
    const size_t smem = Z::smem_bytes(sizeof(cplx<T>));
    if (int e = set_smem(kern, smem)) return e;
    dim3 grid((a.n_t + Z::KS - 1) / Z::KS, 1, a.nbc), block(Z::NT);
    FSM_LAUNCH(kern, grid, block, smem, s, a.g, a.state, a.w1, a.state_bstride, a.w1_fstride, a.in_t_stride, a.out_e_stride,
               a.n_t);
    return check_launch();
}
#ifndef FSM_IXP
// Persistent inverse-x kernel with cp.async staging of the next tile's state lines (k_pass_ixp). Measured on C3
// (profiles/r2_kernel_variants.md): 0.700 ms/step against 0.685 for the one-tile-per-CTA kernel, 0.363 against
// 0.319 at 512 points: hiding the first-load latency buys nothing, the pass is not bound by it. Kept off.
#define FSM_IXP 0
#endif
#ifndef FSM_EMU
static inline int sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return sms;
}
#else
static inline int sm_count() { return 3; }   // the emulator runs a few "SMs" so that the persistent loops iterate
#endif
template <typename T, int N, int PROG>
static int launch_ixp_p(const IxArgs<T>& a, cudaStream_t s) {
    using Cfg = typename CfgFor<N>::type;
    if constexpr (Cfg::NST < 2) {
        return -ENOSYS;
    } else {
        auto kern = k_pass_ixp<T, Cfg, PROG>;
        const size_t smem = Smem<Cfg, T>::bytes(2 * kKL) + sizeof(cplx<T>) * (size_t)kKL * Cfg::N;
        if (int e = set_smem(kern, smem)) return e;
        const int tiles = (a.n_t + kKL - 1) / kKL;
        const int items = tiles * a.nbc;
        static const int per_sm = [&] {
#ifndef FSM_EMU
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kKL * Cfg::TL, smem) != cudaSuccess) n = 1;
            return n < 1 ? 1 : n;
#else
            return 1;
#endif
        }();
        int ctas = sm_count() * per_sm;
        if (ctas < 1) ctas = 1;
        if (ctas > items) ctas = items;
        FSM_LAUNCH(kern, dim3(ctas), dim3(kKL * Cfg::TL), smem, s, a.g, a.state, a.w1, a.state_bstride, a.w1_fstride,
                   a.in_t_stride, a.out_e_stride, a.n_t, tiles, items);
        return check_launch();
    }
}
template <typename T, int N>
static int launch_ix(int prog, const IxArgs<T>& a, cudaStream_t s) {
    if (a.n_outer == 1 && a.eb.shift >= 30 && a.g.ndim == 2 && (prog == PROG_NS2D || prog == PROG_KS2D)) {
        using Cfg = typename CfgFor<N>::type;
        if (FSM_IXZ && Cfg::TL <= 16) {      // measured: the two-CTA form wins up to 256 points, loses from 512 on
            if (prog == PROG_NS2D) return launch_ixz_p<T, N, PROG_NS2D>(a, s);
            return launch_ixz_p<T, N, PROG_KS2D>(a, s);
        }
        if constexpr (FSM_IXP != 0) {
            if (Cfg::TL >= 32 && sizeof(T) == 4) {
                if (prog == PROG_NS2D) return launch_ixp_p<T, N, PROG_NS2D>(a, s);
                return launch_ixp_p<T, N, PROG_KS2D>(a, s);
            }
        }
    }
    switch (prog) {
        case PROG_NS2D: return launch_ix_p<T, N, PROG_NS2D>(a, s);
        case PROG_KS2D: return launch_ix_p<T, N, PROG_KS2D>(a, s);
        case PROG_C2R: return launch_ix_p<T, N, PROG_C2R>(a, s);
        case PROG_CONV: case PROG_KS: case PROG_NS3D: return launch_ix_p<T, N, PROG_CONV>(a, s);
        default: return -ENOSYS;
    }
}

template <typename T, int N, int DIR, int IMODE, int EMODE>
static int launch_mid_m(const MidArgs<T>& a, cudaStream_t s) {
    using Cfg = typename CfgFor<N>::type;
    auto kern = k_pass_mid<T, Cfg, DIR, IMODE, EMODE>;
    const size_t smem = Smem<Cfg, T>::bytes(2 * kKL);
    if (int e = set_smem(kern, smem)) return e;
    dim3 grid((a.n_t + kKL - 1) / kKL, a.n_outer, a.nb), block(kKL * Cfg::TL);
    FSM_LAUNCH(kern, grid, block, smem, s, a.g, a.in, a.out, a.in_fstride, a.out_fstride, a.nfi, a.spec, kKL,
               a.in_t_stride, a.in_o_stride, a.out_o_stride, a.out_e_stride, a.n_t, a.ib, a.eb, a.pe);
    return check_launch();
}
// inverse side (dir > 0): contiguous input (one GPU) or the cyclic rank blocks of a receive buffer, plain store;
// forward side: contiguous input, plain store / cyclic rank blocks of a send buffer / the peers' receive buffers
template <typename T, int N>
static int launch_mid(int dir, const MidArgs<T>& a, cudaStream_t s) {
    const bool in_blocked = a.ib.shift < 30, out_blocked = a.eb.shift < 30;
    if (dir > 0) {
        if (out_blocked || a.pe.n > 0) return -EINVAL;
        if (in_blocked && a.ib.cyc <= 0) return -EINVAL;
        return in_blocked ? launch_mid_m<T, N, +1, 1, 0>(a, s) : launch_mid_m<T, N, +1, 0, 0>(a, s);
    }
    if (in_blocked) return -EINVAL;
    if (!out_blocked) return launch_mid_m<T, N, -1, 0, 0>(a, s);
    if (a.eb.cyc <= 0) return -EINVAL;
    return a.pe.n > 0 ? launch_mid_m<T, N, -1, 0, 2>(a, s) : launch_mid_m<T, N, -1, 0, 1>(a, s);
}

#ifndef FSM_PHYS3D_NL512
// thread-lines per CTA of the 3-D convection last-axis pass at 512 points. Round 1 measured 4 (two 256-thread CTAs per
// SM, 32-byte store segments, 80 B of spills) 9.68 ms/step against 9.75 with 8: no gain. With the third component parked
// in shared memory and the conflict-free last-stage layout (round 2) 4 lines win: 9.68 -> 9.39 ms per C5 step
// (profiles/r2_kernel_variants.md).
#define FSM_PHYS3D_NL512 4
#endif
#ifndef FSM_PHYS3D_NL256
#define FSM_PHYS3D_NL256 8
#endif
template <typename T, int N, int PROG, int NDIM>
static int launch_phys_p(const PhysArgs<T>& a, cudaStream_t s) {
    using Cfg = typename CfgPhys<N, NDIM>::type;
    using PT = PhysTraits<PROG, NDIM>;
    constexpr int NFW = (PT::NOUT * PT::RPT + 1) / 2;
    constexpr int NLP = (PROG == PROG_CONV && NDIM == 3 && N == 512 && sizeof(T) == 4) ? FSM_PHYS3D_NL512
                        : ((PROG == PROG_CONV && NDIM == 3 && N == 256 && sizeof(T) == 4) ? FSM_PHYS3D_NL256 : kKL);
    auto kern = k_pass_phys<T, Cfg, PROG, NDIM, NLP>;
    const int K = NLP * PT::RPT;
    const size_t smem = Smem<Cfg, T>::bytes(NLP * (1 + (NFW > 0 ? NFW : 0)));
    if (int e = set_smem(kern, smem)) return e;
    dim3 grid((a.n_t + K - 1) / K, a.n_outer, a.nb), block(NLP * Cfg::TL);
    static const int wave = resident_ctas(kern, NLP * Cfg::TL, smem);
    Geom<T> g = a.g;
    g.pf_wave = wave;
    FSM_LAUNCH(kern, grid, block, smem, s, g, a.win, a.wout, a.phys_in, a.phys_out, a.win_fstride, a.wout_fstride, K,
               a.in_t_stride, a.in_o_stride, a.out_o_stride, a.out_e_stride, a.n_t, 1);
    return check_launch();
}
#ifndef FSM_PHYSZ
#define FSM_PHYSZ 1   // Z-line programs run the phase-overlapping last-axis kernel (k_pass_physz)
#endif
template <typename T, int N, int PROG>
static int launch_physz_p(const PhysArgs<T>& a, cudaStream_t s) {
    using Cfg = typename CfgFor<N>::type;
    using Z = PhysZ<Cfg>;
    auto kern = k_pass_physz<T, Cfg, PROG>;
    const size_t smem = Z::smem_bytes(sizeof(cplx<T>));
    if (int e = set_smem(kern, smem)) return e;
    dim3 grid((a.n_t + Z::K - 1) / Z::K, 1, a.nb), block(Z::NT);
    FSM_LAUNCH(kern, grid, block, smem, s, a.g, a.win, a.wout, a.win_fstride, a.wout_fstride, a.in_t_stride, a.out_e_stride,
               a.n_t);
    return check_launch();
}
template <typename T, int N>
static int launch_phys(int prog, int ndim, const PhysArgs<T>& a, cudaStream_t s) {
    if (FSM_PHYSZ && ndim == 2 && a.n_outer == 1) {
        if (prog == PROG_NS2D) return launch_physz_p<T, N, PROG_NS2D>(a, s);
        if (prog == PROG_KS2D) return launch_physz_p<T, N, PROG_KS2D>(a, s);
    }
    if (prog == PROG_C2R) return launch_phys_p<T, N, PROG_C2R, 2>(a, s);
    if (prog == PROG_R2C) return launch_phys_p<T, N, PROG_R2C, 2>(a, s);
    if (ndim == 2) {
        if (prog == PROG_NS2D) return launch_phys_p<T, N, PROG_NS2D, 2>(a, s);
        if (prog == PROG_KS) return launch_phys_p<T, N, PROG_KS, 2>(a, s);
        if (prog == PROG_KS2D) return launch_phys_p<T, N, PROG_KS2D, 2>(a, s);
        if (prog == PROG_CONV) return launch_phys_p<T, N, PROG_CONV, 2>(a, s);
    } else if (ndim == 3) {
        if (prog == PROG_KS) return launch_phys_p<T, N, PROG_KS, 3>(a, s);
        if constexpr (N <= 512) {
            if (prog == PROG_CONV || prog == PROG_NS3D) return launch_phys_p<T, N, PROG_CONV, 3>(a, s);
        }
    }
    return -ENOSYS;
}

template <typename T, int N, int C>
static int launch_fx_c(const FxArgs<T>& a, cudaStream_t s) {
    using Cfg = typename CfgFx<N, C>::type;
    auto kern = k_pass_fx<T, Cfg, C>;
    constexpr int KF = kFxLines<Cfg>;
    const size_t smem = Smem<Cfg, T>::bytes(KF);
    if (int e = set_smem(kern, smem)) return e;
    dim3 grid((a.nlines + KF - 1) / KF, 1, a.nb), block(KF * Cfg::TL);
    static const int wave = resident_ctas(kern, KF * Cfg::TL, smem);
    Geom<T> g = a.g;
    g.pf_wave = wave;
    FSM_LAUNCH(kern, grid, block, smem, s, g, a.win, a.win_fstride, a.cb, a.ep, a.nlines, a.b0, a.ib,
               a.line_stride ? a.line_stride : (long)Cfg::N);
    return check_launch();
}
template <typename T, int N>
static int launch_fx(int C, const FxArgs<T>& a, cudaStream_t s) {
    if (C == 1) return launch_fx_c<T, N, 1>(a, s);
    if constexpr (CfgFx<N, 2>::type::EPT * 2 <= 32) {
        if (C == 2) return launch_fx_c<T, N, 2>(a, s);
    }
    if constexpr (CfgFx<N, 3>::type::EPT * 3 <= 48) {
        if (C == 3) return launch_fx_c<T, N, 3>(a, s);
    }
    return -ENOSYS;
}

template <typename T, int N>
static int launch_step1d(const Step1dArgs<T>& a, cudaStream_t s) {
    using Cfg = typename CfgFor<N>::type;
    auto kern = k_step1d<T, Cfg>;
    const size_t smem = Smem<Cfg, T>::bytes(1) + sizeof(cplx<T>) * (size_t)FSM_1D_ARRAYS * (N / 2 + 2);   // + resident arrays
    if (int e = set_smem(kern, smem)) return e;
    FSM_LAUNCH(kern, dim3(a.nb), dim3(Cfg::TL), smem, s, a.g, a.sl, a.ep, a.n_steps);
    return check_launch();
}
template <typename T, int N>
static int launch_line1d(int mode, const void* in, void* out, long nfields, cudaStream_t s) {
    using Cfg = typename CfgFor<N>::type;
    const size_t smem = Smem<Cfg, T>::bytes(1);
    if (mode == MODE1D_R2C) {
        auto kern = k_line1d<T, Cfg, MODE1D_R2C>;
        FSM_LAUNCH(kern, dim3((unsigned)nfields), dim3(Cfg::TL), smem, s, in, out);
    } else {
        auto kern = k_line1d<T, Cfg, MODE1D_C2R>;
        FSM_LAUNCH(kern, dim3((unsigned)nfields), dim3(Cfg::TL), smem, s, in, out);
    }
    return check_launch();
}

// ---- static twiddle tables -------------------------------------------------------------------
#ifndef FSM_EMU
template <typename T, class Cfg>
static int fill_table() {
    if constexpr (Cfg::TW_TOTAL > 0) {
        std::vector<cplx<T>> host(Cfg::TW_TOTAL);
        fill_twiddles<Cfg, T>(host.data());
        if (cudaMemcpyToSymbol(g_twiddles<T, Cfg>, host.data(), sizeof(cplx<T>) * Cfg::TW_TOTAL) != cudaSuccess) return -EIO;
    }
    return 0;
}
template <typename T, int N>
static int prepare_tables() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -EIO;
    std::lock_guard<std::mutex> lk(mu);
    if (done[dev]) return 0;
    if (int e = fill_table<T, typename CfgFor<N>::type>()) return e;
    if constexpr (!std::is_same<typename CfgPhys<N, 3>::type, typename CfgFor<N>::type>::value) {
        if (int e = fill_table<T, typename CfgPhys<N, 3>::type>()) return e;
    }
    if constexpr (!std::is_same<typename CfgFx<N, 3>::type, typename CfgFor<N>::type>::value &&
                  !std::is_same<typename CfgFx<N, 3>::type, typename CfgPhys<N, 3>::type>::value) {
        if (int e = fill_table<T, typename CfgFx<N, 3>::type>()) return e;
    }
    done[dev] = true;
    return 0;
}
#else
template <typename T, int N> static int prepare_tables() { return 0; }
#endif

#define FSM_CAT2(a, b) a##b
#define FSM_CAT(a, b) FSM_CAT2(a, b)
const LaunchTable<float>* FSM_CAT(table_f32_, FSM_N)() {
    static const LaunchTable<float> t = {FSM_N, launch_ix<float, FSM_N>, launch_mid<float, FSM_N>,
                                         launch_phys<float, FSM_N>, launch_fx<float, FSM_N>,
                                         launch_step1d<float, FSM_N>, launch_line1d<float, FSM_N>,
                                         prepare_tables<float, FSM_N>};
    return &t;
}
const LaunchTable<double>* FSM_CAT(table_f64_, FSM_N)() {
    static const LaunchTable<double> t = {FSM_N, launch_ix<double, FSM_N>, launch_mid<double, FSM_N>,
                                          launch_phys<double, FSM_N>, launch_fx<double, FSM_N>,
                                          launch_step1d<double, FSM_N>, launch_line1d<double, FSM_N>,
                                          prepare_tables<double, FSM_N>};
    return &t;
}

}  // namespace fsm

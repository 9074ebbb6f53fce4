// Line-pass kernels of the fused pseudo-spectral step (2-D and 3-D grids).
//
// Layouts (complex unless noted; B = batch, F = field index, C = channels):
//   physical   (B, C, n0, n1[, n2])   real, last axis contiguous ("rows")
//   spectral   "rot-half": 2-D [ky<nh][kx<n0];  3-D [ky<n1][kz<nh][kx<n0]   (nh = nlast/2+1)
//   W1  (after inverse-x)   2-D [x][ky<ph];      3-D [kz][x][ky]
//   W3  (after inverse-y)   3-D [x][y][kz<ph]
//   W2a (after forward-last) 2-D [ky][x];        3-D [kz][x][y]
//   W2b (after forward-y)   3-D [ky][kz][x]
// Every pass reads contiguous lines, transforms them in registers/shared memory and writes
// either contiguous lines or "rotated" (line index becomes the fastest output index), so no
// pass ever makes strided global accesses.
//
// One nonlinear evaluation = IX -> [MID inverse] -> PHYS -> [MID forward] -> FX(+combine).
#pragma once
#include "fsm_fft.cuh"

namespace fsm {

#ifndef FSM_KL
#define FSM_KL 8
#endif
// ask ptxas for at least 512 resident threads per SM (register cap 128)
#ifndef FSM_TARGET_THREADS
#define FSM_TARGET_THREADS 512
#endif
#define FSM_MINB(nt) (((nt) >= FSM_TARGET_THREADS) ? 1 : (FSM_TARGET_THREADS / (nt)))
constexpr int kKL = FSM_KL;  // thread-lines per CTA in every line pass
#ifndef FSM_KL_FX
#define FSM_KL_FX 4   // FX has no rotated store: fewer lines per CTA -> more CTAs per SM overlap their phases
#endif
#ifndef FSM_FX_MINB
#define FSM_FX_MINB 3 // resident FX CTAs per SM the register allocation aims for (single-channel lines)
#endif
// measured: the smaller FX tile helps when a line already spans >= 32 threads (N >= 512)
template <class Cfg> constexpr int kFxLines = (Cfg::TL >= 32 && FSM_KL_FX < kKL) ? FSM_KL_FX : kKL;

enum Prog : int {
    PROG_NONE = 0,
    PROG_CONV = 1,     // u.grad(u), C == ndim            (reference: generic/_convection.py:43-46)
    PROG_KS = 2,       // 1/2 |grad phi|^2, C == 1         (dedicated/_ks_convection.py:33-38)
    PROG_NS2D = 3,     // vorticity convection, 2-D, C == 1 (dedicated/_navier_stokes.py:41-46)
    PROG_NS3D = 4,     // conv + pressure projection, 3-D   (dedicated/_navier_stokes.py:241-254)
    PROG_C2R = 5,      // plain inverse transform to physical space (mesh.py:487-491, .real)
    PROG_R2C = 6,      // plain forward transform of a real field   (mesh.py:481-485)
    PROG_KS2D = 7,     // 2-D KS with paired Z-lines (kernel-side variant of PROG_KS)
};

template <typename T>
struct Geom {
    int ndim;
    int n[3];      // physical sizes (x, y, z); unused entries = 1
    int nh;        // nlast/2 + 1
    int ph;        // padded pitch of the half axis in W1 (2-D) / W3 (3-D)
    int kmax[3];   // dealiasing: keep |m_i| <= kmax[i]
    long nmodes;   // complex modes per field in the rot-half layout
    T inv_ntot;    // 1 / (n0*n1*n2)
    // slab decomposition (one GPU: ky0 = 0, kys = 1, gap = 0): ky ownership is CYCLIC, local line t holds the global
    // ky = ky0 + kys * t (ky0 = rank, kys = ranks), so every rank owns the same share of the dealiased band. On the
    // inverse side only the kept local lines are processed and shipped: compact index t' -> local line
    // t' < gap_at ? t' : t' + gap (the dropped middle of the spectrum is a contiguous run of local lines).
    int ky0, kys, gap_at, gap;
    int pf;        // L2 prefetch switches (PF_* bits)
    int pf_wave;   // CTAs resident on the whole GPU for this launch = prefetch distance in CTAs
    const T* dk[3];     // 2*pi*f_i(m), Nyquist entry zeroed (Hermitian projection of i*k)
    const T* dkraw[3];  // 2*pi*f_i(m) as the reference computes it (Nyquist kept, negative)
};

// Rank-blocked index (slab decomposition): index i lives in block i >> shift (block stride `stride`
// elements) at position i & mask inside it. One GPU: shift = 30, i.e. a single block.
struct Blk {
    int shift;      // rank block:     i >> shift
    long stride;
    int shift2;     // sub-slab block: (i & mask) >> shift2   (30 = none)
    long stride2;
    // cyclic ownership (the ky axis of a slab-decomposed grid): rank = i & (2^cyc - 1), local index t = i >> cyc,
    // stored at the compact position t < gap_at ? t : t - gap. cyc = 0: blocked ownership as described above.
    int cyc, gap_at, gap;
};
// Direct exchange (slab decomposition): instead of a local send buffer the pass stores straight into the
// receive buffers of the destination ranks (peer memory over NVLink; plain addresses in the emulator).
// Element i of the exchanged axis belongs to rank i >> Blk::shift; that rank's buffer holds one block per
// source rank, ours at self_off.
#define FSM_MAX_PEERS 8
struct Peers {
    void* base[FSM_MAX_PEERS];
    int n;            // 0: no direct exchange
    long self_off;    // elements: (this rank) * Blk::stride
};
__device__ __forceinline__ long blk_off(int i, const Blk& b, long elem_stride) {
    if (b.cyc > 0) {
        const int t = i >> b.cyc;
        return (long)(i & ((1 << b.cyc) - 1)) * b.stride + (long)(t < b.gap_at ? t : t - b.gap) * elem_stride;
    }
    const int r = i & ((1 << b.shift) - 1);
    const int lo = (b.shift2 < b.shift) ? b.shift2 : b.shift;
    return (long)(i >> b.shift) * b.stride + (long)(r >> b.shift2) * b.stride2 + (long)(r & ((1 << lo) - 1)) * elem_stride;
}

// ------------------------------------------------------------------------------------------
// L2 prefetch. Every pass is latency bound on its first loads (16 warps per SM leave nothing to switch
// to), so a CTA asks the L2 for data ahead of time with cp.async.bulk.prefetch.L2: no registers, no
// shared memory, nothing to wait on. Two uses: (1) its OWN later operands (FX combine rows) while its
// transform runs; (2) the input lines of the CTA that will occupy this slot one wave later
// (linear block id + pf_wave), so that CTA's first loads hit the L2 instead of HBM.
// ------------------------------------------------------------------------------------------
enum : int { PF_IX = 1, PF_PHYS = 2, PF_FX_OPS = 4, PF_FX_WIN = 8, PF_MID = 16, PF_TW_EARLY = 32, PF_PHYS_SELF = 64 };

__device__ __forceinline__ void l2_prefetch(const void* p, long bytes) {
#ifndef FSM_EMU
    const unsigned long long a = reinterpret_cast<unsigned long long>(p);
    const unsigned long long lo = a & ~15ull, hi = (a + (unsigned long long)bytes + 15ull) & ~15ull;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((unsigned)(hi - lo)) : "memory");
#else
    (void)p; (void)bytes;
#endif
}
// the kept part of one spectral line of n points: modes [0, kmax] and [n - kmax, n)
template <typename T>
__device__ __forceinline__ void l2_prefetch_line(const cplx<T>* line, int n, int kmax) {
    if (2 * kmax + 1 >= n) {
        l2_prefetch(line, (long)n * sizeof(cplx<T>));
    } else {
        l2_prefetch(line, (long)(kmax + 1) * sizeof(cplx<T>));
        l2_prefetch(line + (n - kmax), (long)kmax * sizeof(cplx<T>));
    }
}
// block coordinates of the CTA `wave` positions further in launch order (x fastest); false past the end
__device__ __forceinline__ bool next_wave_block(int wave, int& x, int& y, int& z) {
    const unsigned gx = gridDim.x, gy = gridDim.y, gz = gridDim.z;   // launch grids stay far below 2^31 CTAs
    const unsigned lin = blockIdx.x + gx * (blockIdx.y + gy * blockIdx.z) + (unsigned)wave;
    if (wave <= 0 || lin >= gx * gy * gz) return false;
    const unsigned r = lin / gx;
    x = (int)(lin - r * gx);
    z = (int)(r / gy);
    y = (int)(r - (unsigned)z * gy);
    return true;
}

template <typename T>
__device__ __forceinline__ cplx<T>* peer_ptr(const Peers& pe, const Blk& b, int i, long off0, long elem_stride) {
    if (b.cyc > 0) {
        const int t = i >> b.cyc;
        return static_cast<cplx<T>*>(pe.base[i & ((1 << b.cyc) - 1)]) + pe.self_off + off0 +
               (long)(t < b.gap_at ? t : t - b.gap) * elem_stride;
    }
    const int r = i & ((1 << b.shift) - 1);
    return static_cast<cplx<T>*>(pe.base[i >> b.shift]) + pe.self_off + off0 + (long)r * elem_stride;
}

template <int N>
__device__ __forceinline__ int signed_mode(int p) { return (p <= N / 2) ? p : p - N; }
__device__ __forceinline__ int signed_mode_rt(int p, int n) { return (p <= n / 2) ? p : p - n; }
__device__ __forceinline__ int iabs(int a) { return a < 0 ? -a : a; }

// sin/cos of pi*x (emulator build of the stage twiddles; the CUDA build reads a static table)
__device__ __forceinline__ void fsm_sincospi(float x, float* s, float* c) {
#ifdef FSM_EMU
    *s = (float)std::sin(3.14159265358979323846 * (double)x);
    *c = (float)std::cos(3.14159265358979323846 * (double)x);
#else
    sincospif(x, s, c);
#endif
}
__device__ __forceinline__ void fsm_sincospi(double x, double* s, double* c) {
#ifdef FSM_EMU
    *s = std::sin(3.14159265358979323846 * x);
    *c = std::cos(3.14159265358979323846 * x);
#else
    sincospi(x, s, c);
#endif
}

// -1/x: fp32 uses the hardware reciprocal (<= 1 ulp), fp64 the exact division
__device__ __forceinline__ float neg_recip(float x) {
#ifdef FSM_EMU
    return -1.0f / x;
#else
    return -__frcp_rn(x);
#endif
}
__device__ __forceinline__ double neg_recip(double x) { return -1.0 / x; }

#ifndef FSM_TW_CPASYNC
#define FSM_TW_CPASYNC 1
#endif
#ifndef FSM_EMU
// Stage twiddles of one FFT configuration: static device memory, filled once per device by the launch layer
// (LaunchTable::prepare, called from fsm_plan_create); every CTA copies them into shared memory.
template <typename T, class Cfg>
__device__ cplx<T> g_twiddles[Cfg::TW_TOTAL + 2];
#endif

// twiddles_begin starts the copy into shared memory (cp.async: no registers, nothing waited for) so that the
// caller can put its own first global loads in flight; twiddles_ready waits for the copy and publishes it to
// the CTA. The first use of a twiddle is after stage 0 of the first transform.
template <class Cfg, typename T, bool ASYNC = true>
__device__ __forceinline__ void twiddles_begin(cplx<T>* tw) {
#ifndef FSM_EMU
    if constexpr (Cfg::TW_TOTAL > 0) {
#pragma unroll 1
        for (int i = threadIdx.x; i < Cfg::TW_TOTAL; i += blockDim.x) {
            if constexpr (ASYNC && FSM_TW_CPASYNC) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(tw + i);
                asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(&g_twiddles<T, Cfg>[i]), "n"(sizeof(cplx<T>)) : "memory");
            } else {
                tw[i] = g_twiddles<T, Cfg>[i];
            }
        }
    }
#else
    if constexpr (Cfg::R1 > 1) {
        constexpr int Ns = Cfg::R0;
        for (int i = threadIdx.x; i < Cfg::TW1; i += blockDim.x) {
            const int t = i / Ns + 1, j = i % Ns;
            T s, c;
            fsm_sincospi(T(-2) * T(j * t) / T(Ns * Cfg::R1), &s, &c);
            tw[i] = mk<T>(c, s);
        }
    }
    if constexpr (Cfg::R2 > 1) {
        constexpr int Ns = Cfg::R0 * Cfg::R1;
        for (int i = threadIdx.x; i < Cfg::TW2; i += blockDim.x) {
            const int t = i / Ns + 1, j = i % Ns;
            T s, c;
            fsm_sincospi(T(-2) * T(j * t) / T(Ns * Cfg::R2), &s, &c);
            tw[Cfg::TW1 + i] = mk<T>(c, s);
        }
    }
#endif
}
__device__ __forceinline__ void twiddles_ready() {
#ifndef FSM_EMU
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
    __syncthreads();
}
template <class Cfg, typename T>
__device__ __forceinline__ void make_twiddles(cplx<T>* tw) {
    twiddles_begin<Cfg, T>(tw);
    twiddles_ready();
}

// Shared-memory carve-up used by every pass: [twiddles][NBUF line buffers]
template <class Cfg, typename T>
struct Smem {
    static constexpr int TWPAD = (Cfg::TW_TOTAL + 1) & ~1;
    static size_t bytes(int nbuf) { return sizeof(cplx<T>) * (size_t)(TWPAD + nbuf * Cfg::LINE_PITCH); }
};

// Rotated ("transposed") store of K line buffers that hold natural-order results:
//   dst[e * e_stride + t]  for e < n_e, t < K (t = line slot, fastest)
template <class Cfg, typename T, int SIGN = 1>
__device__ __forceinline__ void rotated_store(const cplx<T>* bufs, int k_lo, int k_hi, cplx<T>* dst, long e_stride) {
    constexpr int K = kKL, NT = kKL * Cfg::TL, N = Cfg::N;
    static_assert((N * K) % NT == 0, "store loop must divide evenly");
    const int t = threadIdx.x % K;
    const int e0 = threadIdx.x / K;
    const cplx<T>* src = bufs + t * Cfg::LINE_PITCH + e0;
    cplx<T>* d = dst + (long)e0 * e_stride + SIGN * t;
    if (t >= k_lo && t < k_hi) {
        FSM_UNROLL
        for (int j = 0; j < (N * K) / NT; ++j) d[(long)j * (NT / K) * e_stride] = src[j * (NT / K)];
    }
}

// Rotated last stage: once line_fft_head has run on every line of the CTA (and a block barrier has
// passed), thread (t = tid % K, widx = tid / K) finishes the work items w = widx + q*TL of line t and
// hands each output X_t[e] to emit(e, value). Consecutive lanes hold consecutive lines, so a store to
// dst[e * stride + t] is coalesced and the natural-order staging round trip through shared memory
// disappears.
template <class Cfg, int DIR, typename T, class F>
__device__ __forceinline__ void rotated_last_stage(const cplx<T>* bufs, const cplx<T>* tw, F&& emit) {
    constexpr int RL = LastStage<Cfg>::RL, NS = LastStage<Cfg>::NS;
    const int t = threadIdx.x % kKL, widx = threadIdx.x / kKL;
    const cplx<T>* buf = bufs + t * Cfg::LINE_PITCH;
    static_for<0, Cfg::EPT / RL>([&](auto qc) FSM_INLINE_LAMBDA {
        constexpr int q = decltype(qc)::value;
        const int w = widx + q * Cfg::TL;
        cplx<T> a[RL];
        fft_last_item<Cfg, DIR, T, q * Cfg::TL>(buf, tw, widx, a);
        static_for<0, RL>([&](auto tc) FSM_INLINE_LAMBDA {
            constexpr int tp = decltype(tc)::value;
            emit(w + tp * NS, a[tp]);
        });
    });
}

// ------------------------------------------------------------------------------------------
// Pass IX: inverse transform along x of symbol-multiplied spectra (dealias mask, 1/N scale and
// the x-derivative / stream-function symbols are applied while loading).
//   grid = (tiles of K lines over the line-minor index, outer index, batch*channels)
// Line-minor index t: ky (2-D and 3-D). Outer index: kz (3-D). Output rotated: [.. x][ky].
// Derived fields per channel (NF):
//   PROG_CONV/NS3D/KS: {u_hat, i kx u_hat}                       (2 fields)
//   PROG_NS2D:         {i ky psi, -i kx psi, i kx w, i ky w}     (4 fields), psi = -w / lap
//   PROG_C2R:          {u_hat} without mask                       (1 field)
// ------------------------------------------------------------------------------------------
template <int PROG>
struct IxFields {
    static constexpr int NF = (PROG == PROG_NS2D) ? 4 : ((PROG == PROG_C2R) ? 1 : 2);
    static constexpr int NPAIR = (PROG == PROG_NS2D) ? 2 : 1;   // Z-line programs (NS2D, KS2D)
};

template <typename T, class Cfg, int PROG>
__global__ void __launch_bounds__(kKL * Cfg::TL, FSM_MINB(kKL * Cfg::TL)) k_pass_ix(Geom<T> g, const cplx<T>* __restrict__ state, cplx<T>* __restrict__ w1,
                                                  long state_bstride /*per (b,c)*/, long w1_fstride, int K,
                                                  long in_t_stride, long in_o_stride, long out_o_stride,
                                                  long out_e_stride, int n_t, Blk eb, Peers pe) {
    constexpr int N = Cfg::N, EPT = Cfg::EPT, TL = Cfg::TL;
    constexpr int NF = IxFields<PROG>::NF;
    FSM_DYN_SMEM(smem_raw);
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* bufs = tw + Smem<Cfg, T>::TWPAD;
    const int lt = threadIdx.x / TL, tau = threadIdx.x % TL;
    if ((g.pf & PF_IX) && tau == 0) {
        int x2, o2, z2;
        if (next_wave_block(g.pf_wave, x2, o2, z2)) {
            const int t2 = x2 * K + lt;
            bool want = t2 < n_t;
            if (want && PROG != PROG_C2R) {
                want = iabs(signed_mode_rt(g.ky0 + g.kys * (t2 < g.gap_at ? t2 : t2 + g.gap), g.n[1])) <= g.kmax[1];
                if (g.ndim == 3) want = want && (o2 <= g.kmax[2]);
            }
            if (want)
                l2_prefetch_line<T>(state + (long)z2 * state_bstride + (long)(t2 < g.gap_at ? t2 : t2 + g.gap) * in_t_stride + (long)o2 * in_o_stride, N,
                                    PROG == PROG_C2R ? N : g.kmax[0]);
        }
    }
    twiddles_begin<Cfg, T>(tw);
    if (g.pf & PF_TW_EARLY) twiddles_ready();

    const int t0 = blockIdx.x * K;
    const int o = blockIdx.y;
    const long bc = blockIdx.z;
    const int t = t0 + lt;  // ky index of this thread-line
    const int k_valid = (K < n_t - t0) ? K : (n_t - t0);
    // dealias test on the line coordinates
    const int tloc = (t < n_t) ? (t < g.gap_at ? t : t + g.gap) : 0;   // local line behind the compact index t
    const int tglob = g.ky0 + g.kys * tloc;                             // its global ky
    const int my = signed_mode_rt(tglob, g.n[1]);
    bool line_kept = (t < n_t);
    if (PROG != PROG_C2R) {
        line_kept = line_kept && (iabs(my) <= g.kmax[1]);
        if (g.ndim == 3) line_kept = line_kept && (o <= g.kmax[2]);
    }
    const T dky = (t < n_t) ? g.dk[1][tglob] : T(0);
    const T dkyraw = (t < n_t) ? g.dkraw[1][tglob] : T(0);

    cplx<T> u[EPT];
    const cplx<T>* src = state + bc * state_bstride + (long)tloc * in_t_stride + (long)o * in_o_stride + tau;
    FSM_PIN(src);
    LineSync<TL> sync{1 + lt};
    cplx<T>* mybuf = bufs + lt * Cfg::LINE_PITCH;

    if constexpr (PROG == PROG_NS2D || PROG == PROG_KS2D) {
        // "Z-lines": the two real fields of a pair ride in ONE complex line Z = f_a + i f_b, so the last-axis
        // pass needs no pairing work. Pairs: Z1 = u_x + i d_x w, Z2 = u_y + i d_y w with psi = -w/lap,
        // u_x = d_y psi, u_y = -d_x psi (_navier_stokes.py:41-45). With the four plain x-transforms
        //   A = T[kx w], B = T[i ky ninv w], C = T[w], D = T[i kx ninv w]      (ninv = -1/lap, T = IFFT_x)
        // the stored +ky line and its mirror -ky (w_hat(kx,-ky) = conj(w_hat(-kx,ky))) are
        //   Z1+ = B - A, Z1- = conj(A + B), Z2+ = -ky C - D, Z2- = conj(-ky C + D),
        // formed while the data crosses shared memory for the rotated store.
        // W1 layout: [pair][x][ky' < n1], the -ky line stored at ky' = n1 - ky.
        // KS2D (1/2 |grad phi|^2): one pair Z = phi_x + i phi_y from A = T[kx phi], C = T[phi]:
        //   Z+ = i A - ky C,  Z- = conj(i A + ky C).
        const int n1 = g.n[1];
        constexpr int NPAIR = IxFields<PROG>::NPAIR;
        // the x wavenumbers of this thread's elements are loaded once, together with the line itself
        T dkxr[EPT];
        const T* dkx_t = g.dkraw[0] + tau;
        FSM_PIN(dkx_t);
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) {
            const int p = tau + m * TL;
            const bool kept = line_kept && (iabs(signed_mode<N>(p)) <= g.kmax[0]);
            u[m] = kept ? src[m * TL] : mk<T>(T(0), T(0));
            dkxr[m] = dkx_t[m * TL];
        }
        twiddles_ready();   // the line is in flight; now wait for the twiddle copy
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) u[m] = cscale(u[m], g.inv_ntot);
        static_for<0, 2 * NPAIR>([&](auto fc) FSM_INLINE_LAMBDA {
            constexpr int f = decltype(fc)::value;
            cplx<T> v[EPT];
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) {
                const int p = tau + m * TL;
                // first-derivative symbol: zero on the Nyquist index (Hermitian projection), else the raw value
                const T dkx_h = (2 * p == N) ? T(0) : dkxr[m];
                if constexpr (PROG == PROG_KS2D) {
                    if constexpr (f == 0) v[m] = cscale(u[m], dkx_h);
                    else v[m] = u[m];
                } else if constexpr (f == 2) {
                    v[m] = u[m];
                } else {
                    const T dkx = dkx_h;
                    if constexpr (f == 0) {
                        v[m] = cscale(u[m], dkx);
                    } else {
                        const T dkxraw = dkxr[m];
                        // lap = (i dkxraw)^2 + (i dkyraw)^2 (mesh.py:406-426); psi = -w * where(lap==0, 1, 1/lap)
                        const T lap = -(dkxraw * dkxraw) - (dkyraw * dkyraw);
                        const T ninv = (lap == T(0)) ? T(-1) : neg_recip(lap);
                        v[m] = cmul_i(u[m], (f == 1 ? dky : dkx) * ninv);
                    }
                }
            }
            cplx<T>* pbufs = bufs + (f & 1) * kKL * Cfg::LINE_PITCH;
            line_fft_head<Cfg, +1, T>(v, pbufs + lt * Cfg::LINE_PITCH, tw, tau, sync);
            if constexpr (f & 1) {
                __syncthreads();
                // last stage of both members of the pair in the rotated distribution, combine, store
                constexpr int pair = f >> 1;
                constexpr int RL = LastStage<Cfg>::RL, NS = LastStage<Cfg>::NS;
                const int ts = threadIdx.x % kKL, widx = threadIdx.x / kKL;
                const int tg = t0 + ts;
                const bool valid = ts < k_valid;
                const bool selfc = (tg == 0) || (2 * tg == n1);
                const T dky_s = valid ? g.dk[1][tg] : T(0);
                const cplx<T>* s0 = bufs + ts * Cfg::LINE_PITCH;
                const cplx<T>* s1 = s0 + kKL * Cfg::LINE_PITCH;
                cplx<T>* dp = w1 + (bc * NPAIR + pair) * w1_fstride + tg;
                cplx<T>* dm = w1 + (bc * NPAIR + pair) * w1_fstride + (n1 - tg);
                FSM_PIN(dp);
                FSM_PIN(dm);
                static_for<0, EPT / RL>([&](auto qc) FSM_INLINE_LAMBDA {
                    constexpr int q = decltype(qc)::value;
                    const int w = widx + q * TL;
                    cplx<T> a[RL], b[RL];
                    fft_last_item<Cfg, +1, T, q * TL>(s0, tw, widx, a);
                    fft_last_item<Cfg, +1, T, q * TL>(s1, tw, widx, b);
                    if (valid) {
                        static_for<0, RL>([&](auto tc) FSM_INLINE_LAMBDA {
                            constexpr int tp = decltype(tc)::value;
                            const int off = (w + tp * NS) * (int)out_e_stride;   // inside one field: fits 32 bits
                            cplx<T> zp, zm;
                            if constexpr (PROG == PROG_KS2D) {      // a = A = T[kx phi], b = C = T[phi]
                                zp = selfc ? mk<T>(-a[tp].y, -dky_s * b[tp].y)
                                           : mk<T>(-a[tp].y - dky_s * b[tp].x, a[tp].x - dky_s * b[tp].y);
                                zm = mk<T>(dky_s * b[tp].x - a[tp].y, -(a[tp].x + dky_s * b[tp].y));
                            } else if constexpr (pair == 0) {
                                zp = selfc ? mk<T>(b[tp].x, -a[tp].y) : b[tp] - a[tp];
                                zm = mk<T>(a[tp].x + b[tp].x, -(a[tp].y + b[tp].y));
                            } else {
                                zp = selfc ? mk<T>(-b[tp].x, -dky_s * a[tp].y)
                                           : mk<T>(-dky_s * a[tp].x - b[tp].x, -dky_s * a[tp].y - b[tp].y);
                                zm = mk<T>(dky_s * a[tp].x - b[tp].x, b[tp].y - dky_s * a[tp].y);
                            }
                            dp[off] = zp;
                            if (!selfc) dm[off] = zm;
                        });
                    }
                });
                if constexpr (f == 1 && NPAIR == 2) __syncthreads();
            }
        });
        return;
    }

    FSM_UNROLL
    for (int m = 0; m < EPT; ++m) {
        const int p = tau + m * TL;
        bool kept = line_kept;
        if (PROG != PROG_C2R) kept = kept && (iabs(signed_mode<N>(p)) <= g.kmax[0]);
        u[m] = kept ? src[m * TL] : mk<T>(T(0), T(0));
    }
    twiddles_ready();
    FSM_UNROLL
    for (int m = 0; m < EPT; ++m) u[m] = cscale(u[m], g.inv_ntot);

    static_for<0, NF>([&](auto fc) FSM_INLINE_LAMBDA {
        constexpr int f = decltype(fc)::value;
        cplx<T> v[EPT];
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) {
            const int p = tau + m * TL;
            if constexpr (f == 0) v[m] = u[m];
            else v[m] = cmul_i(u[m], g.dk[0][p]);                               // d_x
        }
        cplx<T>* pbufs = bufs + (f & 1) * kKL * Cfg::LINE_PITCH;
        line_fft_head<Cfg, +1, T>(v, pbufs + lt * Cfg::LINE_PITCH, tw, tau, sync);
        __syncthreads();
        cplx<T>* dst = w1 + (bc * NF + f) * w1_fstride + (long)o * out_o_stride + t0 + threadIdx.x % kKL;
        const bool valid = (int)(threadIdx.x % kKL) < k_valid;
        if (eb.shift >= 30) {   // one GPU: plain strided store, 32-bit index inside the field
            const int es = (int)out_e_stride;
            rotated_last_stage<Cfg, +1, T>(pbufs, tw, [&](int e, cplx<T> val) FSM_INLINE_LAMBDA {
                if (valid) dst[e * es] = val;
            });
        } else if (pe.n > 0) {  // direct exchange: the x-slab owner's receive buffer
            const long off0 = dst - w1;
            rotated_last_stage<Cfg, +1, T>(pbufs, tw, [&](int e, cplx<T> val) FSM_INLINE_LAMBDA {
                if (valid) *peer_ptr<T>(pe, eb, e, off0, out_e_stride) = val;
            });
        } else {
            rotated_last_stage<Cfg, +1, T>(pbufs, tw, [&](int e, cplx<T> val) FSM_INLINE_LAMBDA {
                if (valid) dst[blk_off(e, eb, out_e_stride)] = val;
            });
        }
    });
}

// ------------------------------------------------------------------------------------------
// Pass IX-P: k_pass_ix's Z-line branch (NS2D, KS2D) as a PERSISTENT kernel for long lines. One 512-thread CTA per
// SM (register file) loops over (ky tile, sample) work items; the state lines of item i+1 are copied into a
// staging area of shared memory with cp.async (thread-private: every thread copies exactly the elements it will
// read back, so no barrier is involved) while the four transforms of item i run, and the stores of item i drain
// while item i+1 computes. ncu on the one-tile-per-CTA kernel: 18 % of the warp-stall samples sat on the first
// use of the freshly loaded line, 2 % on the store drain at exit, plus launch gaps.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void async_copy8(void* smem_dst, const void* gsrc) {
#ifndef FSM_EMU
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gsrc) : "memory");
#else
    *static_cast<unsigned long long*>(smem_dst) = *static_cast<const unsigned long long*>(gsrc);
#endif
}
__device__ __forceinline__ void async_copy16(void* smem_dst, const void* gsrc) {
#ifndef FSM_EMU
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gsrc) : "memory");
#else
    static_cast<unsigned long long*>(smem_dst)[0] = static_cast<const unsigned long long*>(gsrc)[0];
    static_cast<unsigned long long*>(smem_dst)[1] = static_cast<const unsigned long long*>(gsrc)[1];
#endif
}
__device__ __forceinline__ void async_copy_wait() {
#ifndef FSM_EMU
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

template <typename T, class Cfg, int PROG>
__global__ void __launch_bounds__(kKL * Cfg::TL, FSM_MINB(kKL * Cfg::TL))
k_pass_ixp(Geom<T> g, const cplx<T>* __restrict__ state, cplx<T>* __restrict__ w1, long state_bstride, long w1_fstride,
           long in_t_stride, long out_e_stride, int n_t, int tiles_per_sample, int n_items) {
    static_assert(PROG == PROG_NS2D || PROG == PROG_KS2D, "Z-line programs only");
    constexpr int N = Cfg::N, EPT = Cfg::EPT, TL = Cfg::TL;
    constexpr int NPAIR = IxFields<PROG>::NPAIR;
    FSM_DYN_SMEM(smem_raw);
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* bufs = tw + Smem<Cfg, T>::TWPAD;
    cplx<T>* stage = bufs + 2 * kKL * Cfg::LINE_PITCH;          // [kKL][N] staging of the next item's state lines
    const int lt = threadIdx.x / TL, tau = threadIdx.x % TL;
    const int n1 = g.n[1];
    twiddles_begin<Cfg, T>(tw);
    LineSync<TL> sync{1 + lt};
    cplx<T>* mystage = stage + lt * N + tau;
    // element-wise constants of this thread: x wavenumbers and the dealiasing predicate along x
    T dkxr[EPT];
    unsigned keepx = 0;
    {
        const T* dkx_t = g.dkraw[0] + tau;
        FSM_PIN(dkx_t);
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) {
            dkxr[m] = dkx_t[m * TL];
            if (iabs(signed_mode<N>(tau + m * TL)) <= g.kmax[0]) keepx |= 1u << m;
        }
    }
    auto line_of = [&](int item, int& t, long& bc) FSM_INLINE_LAMBDA {
        bc = item / tiles_per_sample;
        t = (item - (int)bc * tiles_per_sample) * kKL + lt;
    };
    auto line_mask = [&](int t) FSM_INLINE_LAMBDA -> unsigned {
        const bool kept = (t < n_t) && (iabs(signed_mode_rt(t, n1)) <= g.kmax[1]);
        return kept ? keepx : 0u;
    };
    auto prefetch_item = [&](int item) FSM_INLINE_LAMBDA {
        int t; long bc;
        line_of(item, t, bc);
        const unsigned km = line_mask(t);
        const cplx<T>* src = state + bc * state_bstride + (long)(km ? t : 0) * in_t_stride + tau;
        FSM_PIN(src);
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) {
            if ((km >> m) & 1u) {
                if constexpr (sizeof(cplx<T>) == 8) async_copy8(mystage + m * TL, src + m * TL);
                else async_copy16(mystage + m * TL, src + m * TL);
            }
        }
    };
    int item = blockIdx.x;
    if (item < n_items) prefetch_item(item);
    twiddles_ready();
    for (; item < n_items; item += gridDim.x) {
        int t; long bc;
        line_of(item, t, bc);
        const int t0 = t - lt;
        const int k_valid = (kKL < n_t - t0) ? kKL : (n_t - t0);
        const unsigned km = line_mask(t);
        const T dky = km ? g.dk[1][t] : T(0);
        const T dkyraw = km ? g.dkraw[1][t] : T(0);
        cplx<T> u[EPT];
        async_copy_wait();
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) u[m] = ((km >> m) & 1u) ? cscale(mystage[m * TL], g.inv_ntot) : mk<T>(T(0), T(0));
        if (item + (int)gridDim.x < n_items) prefetch_item(item + gridDim.x);   // own elements only: no barrier needed
        static_for<0, 2 * NPAIR>([&](auto fc) FSM_INLINE_LAMBDA {
            constexpr int f = decltype(fc)::value;
            cplx<T> v[EPT];
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) {
                const int p = tau + m * TL;
                const T dkx_h = (2 * p == N) ? T(0) : dkxr[m];
                if constexpr (PROG == PROG_KS2D) {
                    if constexpr (f == 0) v[m] = cscale(u[m], dkx_h);
                    else v[m] = u[m];
                } else if constexpr (f == 2) {
                    v[m] = u[m];
                } else if constexpr (f == 0) {
                    v[m] = cscale(u[m], dkx_h);
                } else {
                    const T lap = -(dkxr[m] * dkxr[m]) - (dkyraw * dkyraw);
                    const T ninv = (lap == T(0)) ? T(-1) : neg_recip(lap);
                    v[m] = cmul_i(u[m], (f == 1 ? dky : dkx_h) * ninv);
                }
            }
            cplx<T>* pbufs = bufs + (f & 1) * kKL * Cfg::LINE_PITCH;
            line_fft_head<Cfg, +1, T>(v, pbufs + lt * Cfg::LINE_PITCH, tw, tau, sync);
            if constexpr (f & 1) {
                __syncthreads();
                constexpr int pair = f >> 1;
                constexpr int RL = LastStage<Cfg>::RL, NS = LastStage<Cfg>::NS;
                const int ts = threadIdx.x % kKL, widx = threadIdx.x / kKL;
                const int tg = t0 + ts;
                const bool valid = ts < k_valid;
                const bool selfc = (tg == 0) || (2 * tg == n1);
                const T dky_s = valid ? g.dk[1][tg] : T(0);
                const cplx<T>* s0 = bufs + ts * Cfg::LINE_PITCH;
                const cplx<T>* s1 = s0 + kKL * Cfg::LINE_PITCH;
                cplx<T>* dp = w1 + (bc * NPAIR + pair) * w1_fstride + tg;
                cplx<T>* dm = w1 + (bc * NPAIR + pair) * w1_fstride + (n1 - tg);
                FSM_PIN(dp);
                FSM_PIN(dm);
                static_for<0, EPT / RL>([&](auto qc) FSM_INLINE_LAMBDA {
                    constexpr int q = decltype(qc)::value;
                    const int w = widx + q * TL;
                    cplx<T> a[RL], b[RL];
                    fft_last_item<Cfg, +1, T, q * TL>(s0, tw, widx, a);
                    fft_last_item<Cfg, +1, T, q * TL>(s1, tw, widx, b);
                    if (valid) {
                        static_for<0, RL>([&](auto tc) FSM_INLINE_LAMBDA {
                            constexpr int tp = decltype(tc)::value;
                            const int off = (w + tp * NS) * (int)out_e_stride;
                            cplx<T> zp, zm;
                            if constexpr (PROG == PROG_KS2D) {
                                zp = selfc ? mk<T>(-a[tp].y, -dky_s * b[tp].y)
                                           : mk<T>(-a[tp].y - dky_s * b[tp].x, a[tp].x - dky_s * b[tp].y);
                                zm = mk<T>(dky_s * b[tp].x - a[tp].y, -(a[tp].x + dky_s * b[tp].y));
                            } else if constexpr (pair == 0) {
                                zp = selfc ? mk<T>(b[tp].x, -a[tp].y) : b[tp] - a[tp];
                                zm = mk<T>(a[tp].x + b[tp].x, -(a[tp].y + b[tp].y));
                            } else {
                                zp = selfc ? mk<T>(-b[tp].x, -dky_s * a[tp].y)
                                           : mk<T>(-dky_s * a[tp].x - b[tp].x, -dky_s * a[tp].y - b[tp].y);
                                zm = mk<T>(dky_s * a[tp].x - b[tp].x, b[tp].y - dky_s * a[tp].y);
                            }
                            dp[off] = zp;
                            if (!selfc) dm[off] = zm;
                        });
                    }
                });
                __syncthreads();   // the next heads (this item's second pair or the next item) overwrite the buffers
            }
        });
    }
}

// Tile geometry shared by the two Z-line kernels (k_pass_ixz below, k_pass_physz further down)
template <class Cfg>
struct PhysZ {
    static constexpr int NLZ = (Cfg::TL <= 16) ? 8 : 4;     // thread-lines per CTA
    static constexpr int KS = 2 * NLZ;                      // head buffers = rotated lanes
    static constexpr int NT = NLZ * Cfg::TL;
    static constexpr int K = 4 * NLZ;                       // rows per CTA
    static constexpr bool PARK = (Cfg::EPT >= 16);          // product of the even row waits in shared memory
    static constexpr int PARK_PITCH = Cfg::N / 2;           // complex slots per thread-line
    static size_t smem_bytes(size_t elem) {
        return elem * (size_t)(((Cfg::TW_TOTAL + 1) & ~1) + KS * Cfg::LINE_PITCH + (PARK ? NLZ * PARK_PITCH : 0));
    }
};

// ------------------------------------------------------------------------------------------
// Pass IX-Z: inverse-x pass of the Z-line programs (NS2D, KS2D) on one GPU, organised like PHYS-Z for phase
// overlap: NLZ thread-lines per CTA, each taking TWO ky lines (256 threads and 70 KB of shared memory at 1024
// points -> two independent CTAs per SM). Every stored line is ONE plain transform of (complex symbol) x state:
//   NS2D  Z1+ = T[(-kx + i ky ninv) w]        Z1- = conj T[( kx + i ky ninv) w]      (psi = ninv w, ninv = -1/lap)
//         Z2+ = T[(-ky - i kx ninv) w]        Z2- = conj T[( ky - i kx ninv) w]
//   KS2D  Z+  = T[(-ky + i kx) phi]           Z-  = conj T[( ky + i kx) phi]
// (the linear combinations of k_pass_ix's four transforms A..D moved in front of the transform), so a field needs
// one set of head buffers only and the last stage stores straight from registers in the rotated distribution.
// The state line is re-read per field (L1/L2 hits after the first touch) instead of being held in registers.
// Lines that are their own mirror (ky = 0, ky = n1/2) keep the Hermitian projection k_pass_ix applies: the
// component a real field cannot have is dropped.
// ------------------------------------------------------------------------------------------
#ifndef FSM_IXZ_PREFETCH
#define FSM_IXZ_PREFETCH 1
#endif
template <typename T, class Cfg, int PROG>
__global__ void __launch_bounds__(PhysZ<Cfg>::NT, FSM_MINB(PhysZ<Cfg>::NT))
k_pass_ixz(Geom<T> g, const cplx<T>* __restrict__ state, cplx<T>* __restrict__ w1, long state_bstride, long w1_fstride,
           long in_t_stride, long out_e_stride, int n_t) {
    static_assert(PROG == PROG_NS2D || PROG == PROG_KS2D, "Z-line programs only");
    constexpr int N = Cfg::N, EPT = Cfg::EPT, TL = Cfg::TL;
    using Z = PhysZ<Cfg>;
    constexpr int KS = Z::KS;                    // ky lines per CTA = head buffers = rotated lanes
    constexpr int NPAIR = IxFields<PROG>::NPAIR;
    constexpr int RL = LastStage<Cfg>::RL, NS = LastStage<Cfg>::NS;
    FSM_DYN_SMEM(smem_raw);
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* hb = tw + Smem<Cfg, T>::TWPAD;
    const int lt = threadIdx.x / TL, tau = threadIdx.x % TL;
    twiddles_begin<Cfg, T>(tw);
    const int t0 = blockIdx.x * KS;
    const long bc = blockIdx.z;
    const int n1 = g.n[1];
    LineSync<TL> sync{1 + lt};
    // per element of this thread: x wavenumber (raw, and zero on the Nyquist index) and the dealiasing predicate
    T dkxr[EPT];
    unsigned keepmask = 0;
    {
        const T* dkx_t = g.dkraw[0] + tau;
        FSM_PIN(dkx_t);
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) {
            const int p = tau + m * TL;
            dkxr[m] = dkx_t[m * TL];
            if (iabs(signed_mode<N>(p)) <= g.kmax[0]) keepmask |= 1u << m;
        }
    }
    // rotated distribution of the last stage / store
    const int s = threadIdx.x % KS, widx = threadIdx.x / KS;     // widx < TL / 2
    const int tg = t0 + s;
    const bool valid = tg < n_t;
    const bool selfc = (tg == 0) || (2 * tg == n1);
    const cplx<T>* sbuf = hb + s * Cfg::LINE_PITCH;
    const int es = (int)out_e_stride;
    bool tw_pending = true;

    // per line of this thread-line: y wavenumbers, source pointer, keep mask (loaded once, not per field)
    T dky2[2], dkyraw2[2];
    const cplx<T>* src2[2];
    unsigned km2[2];
    FSM_UNROLL
    for (int r = 0; r < 2; ++r) {
        const int t = t0 + 2 * lt + r;
        const bool line_kept = (t < n_t) && (iabs(signed_mode_rt(t, n1)) <= g.kmax[1]);
        dky2[r] = line_kept ? g.dk[1][t] * g.inv_ntot : T(0);
        dkyraw2[r] = line_kept ? g.dkraw[1][t] : T(0);
        src2[r] = state + bc * state_bstride + (long)(line_kept ? t : 0) * in_t_stride + tau;
        FSM_PIN(src2[r]);
        km2[r] = line_kept ? keepmask : 0u;
    }
    auto load_raw = [&](int r, cplx<T>* raw) FSM_INLINE_LAMBDA {
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) raw[m] = ((km2[r] >> m) & 1u) ? src2[r][m * TL] : mk<T>(T(0), T(0));
    };
    cplx<T> raw[EPT];
    load_raw(0, raw);
    static_for<0, 2 * NPAIR>([&](auto fc) FSM_INLINE_LAMBDA {
        constexpr int f = decltype(fc)::value;
        constexpr int pair = f >> 1;
        constexpr bool minus = (f & 1) != 0;
        static_for<0, 2>([&](auto rc) FSM_INLINE_LAMBDA {
            constexpr int r = decltype(rc)::value;
            const T dky = dky2[r], dkyraw = dkyraw2[r];     // dky carries the 1/N scale
            cplx<T> v[EPT];
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) {
                const int p = tau + m * TL;
                const T dkx_h = ((2 * p == N) ? T(0) : dkxr[m]) * g.inv_ntot;
                T sr, si;
                if constexpr (PROG == PROG_KS2D) {
                    sr = minus ? dky : -dky;
                    si = dkx_h;
                } else {
                    const T lap = -(dkxr[m] * dkxr[m]) - (dkyraw * dkyraw);
                    const T ninv = (lap == T(0)) ? T(-1) : neg_recip(lap);
                    if constexpr (pair == 0) {
                        sr = minus ? dkx_h : -dkx_h;
                        si = dky * ninv;
                    } else {
                        sr = minus ? dky : -dky;
                        si = -(dkx_h * ninv);
                    }
                }
                v[m] = cmul(raw[m], mk<T>(sr, si));
            }
            if (tw_pending) { twiddles_ready(); tw_pending = false; }
            line_fft_head<Cfg, +1, T>(v, hb + (2 * lt + r) * Cfg::LINE_PITCH, tw, tau, sync);
            // next item's state line: (f, 1) is needed at once; (f + 1, 0) is requested before the block barrier and the
            // last stage of this field, which hide its latency (holding it across a head would spill)
            if constexpr (r == 0) load_raw(1, raw);
            else if constexpr (FSM_IXZ_PREFETCH && f + 1 < 2 * NPAIR) load_raw(0, raw);
        });
        __syncthreads();
        {
            cplx<T>* dst = w1 + (bc * NPAIR + pair) * w1_fstride + (minus ? (n1 - tg) : tg);
            FSM_PIN(dst);
            const bool store = valid && !(minus && selfc);
            static_for<0, 2 * Cfg::EPT / RL>([&](auto qc) FSM_INLINE_LAMBDA {
                constexpr int q = decltype(qc)::value;
                cplx<T> a[RL];
                fft_last_item<Cfg, +1, T, q * (TL / 2)>(sbuf, tw, widx, a);
                const int w = widx + q * (TL / 2);
                if (store) {
                    static_for<0, RL>([&](auto tc) FSM_INLINE_LAMBDA {
                        constexpr int tp = decltype(tc)::value;
                        cplx<T> val = a[tp];
                        if constexpr (minus) {
                            val.y = -val.y;
                        } else {
                            // own mirror: NS2D pair 0 and KS2D keep (u . ) as real/imag parts of real fields
                            if (selfc) {
                                if constexpr (PROG == PROG_NS2D && pair == 0) val.x = T(0);
                                else val.y = T(0);
                            }
                        }
                        dst[(w + tp * NS) * es] = val;
                    });
                }
            });
        }
        if constexpr (f + 1 < 2 * NPAIR) {
            __syncthreads();
            if constexpr (!FSM_IXZ_PREFETCH) load_raw(0, raw);
        }
    });
}

// ------------------------------------------------------------------------------------------
// Pass MID (3-D only): C2C transform along y on NFI input fields, producing NFO output fields;
// output field j = transform(in[src[j]] * (deriv[j] ? i*dk_y : 1)). Rotated output.
//   inverse: W1 [kz][x][ky] -> W3 [x][y][kz];   forward: W2a [kz][x][y] -> W2b [ky][kz][x]
// ------------------------------------------------------------------------------------------
struct MidSpec {
    int nfo;
    int src[12];
    int deriv[12];
};

// IMODE: 0 = input lines are contiguous (one GPU, forward side of a slab), 1 = the ky axis of the input is spread over
// the rank blocks of a receive buffer with cyclic ownership (inverse side of a slab-decomposed grid).
// EMODE: 0 = plain rotated store, 1 = the ky axis of the output goes to the rank blocks of a send buffer (cyclic
// ownership), 2 = the same straight into the peers' receive buffers. One instantiation per combination keeps the
// address arithmetic of the other paths (and their registers) out of each kernel.
template <typename T, class Cfg, int DIR, int IMODE, int EMODE>
__global__ void __launch_bounds__(kKL * Cfg::TL, FSM_MINB(kKL * Cfg::TL)) k_pass_mid(Geom<T> g, const cplx<T>* __restrict__ in, cplx<T>* __restrict__ out,
                                                   long in_fstride, long out_fstride, int nfi, MidSpec spec, int K,
                                                   long in_t_stride, long in_o_stride, long out_o_stride,
                                                   long out_e_stride, int n_t, Blk ib, Blk eb, Peers pe) {
    constexpr int N = Cfg::N, EPT = Cfg::EPT, TL = Cfg::TL;
    FSM_DYN_SMEM(smem_raw);
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* bufs = tw + Smem<Cfg, T>::TWPAD;
    twiddles_begin<Cfg, T, false>(tw);   // measured (C4/C5): the y pass is 7 % faster with the plain copy than with cp.async
    if (g.pf & PF_TW_EARLY) twiddles_ready();
    const int lt = threadIdx.x / TL, tau = threadIdx.x % TL;
    const int t0 = blockIdx.x * K;
    const int o = blockIdx.y;
    const long b = blockIdx.z;
    const int t = t0 + lt;
    const int k_valid = (K < n_t - t0) ? K : (n_t - t0);
    const bool line_ok = t < n_t;
    LineSync<TL> sync{1 + lt};
    cplx<T>* mybuf = bufs + lt * Cfg::LINE_PITCH;
    // the line of output field j+1 is loaded (registers) before field j is transformed and stored, so its
    // latency hides behind the butterflies, the block barrier and the rotated stores of field j
    unsigned keepmask = 0;   // bit m: element p = tau + m*TL of an input line is read (dealiasing, valid line)
    FSM_UNROLL
    for (int m = 0; m < EPT; ++m) {
        const int p = tau + m * TL;
        if (line_ok && (DIR < 0 || iabs(signed_mode<N>(p)) <= g.kmax[1])) keepmask |= 1u << m;
    }
    // cyclic input (IMODE 1): ky = tau + m*TL lives in rank block ky & (P-1) at slot ky >> log2(P), compacted by the
    // dropped run of slots; TL is a multiple of P, so the block is fixed per thread and the slot advances by TL/P
    const int icm = (IMODE == 1) ? ((1 << ib.cyc) - 1) : 0;
    const long irank_off = (IMODE == 1) ? (long)(tau & icm) * ib.stride : 0;
    const int islot0 = (IMODE == 1) ? (tau >> ib.cyc) : 0, islot_step = (IMODE == 1) ? (TL >> ib.cyc) : 0;
    auto load_line = [&](int j, cplx<T>* raw) FSM_INLINE_LAMBDA {
        const cplx<T>* src = in + (b * nfi + spec.src[j]) * in_fstride + (long)(line_ok ? t : 0) * in_t_stride + (long)o * in_o_stride;
        if constexpr (IMODE == 0) {
            src += tau;
            FSM_PIN(src);
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) raw[m] = ((keepmask >> m) & 1u) ? src[m * TL] : mk<T>(T(0), T(0));
        } else {
            if constexpr (TL >= FSM_MAX_PEERS) {
                src += irank_off;
                FSM_PIN(src);
                FSM_UNROLL
                for (int m = 0; m < EPT; ++m) {
                    const int slot = islot0 + m * islot_step;
                    raw[m] = ((keepmask >> m) & 1u) ? src[slot - (slot >= ib.gap_at ? ib.gap : 0)] : mk<T>(T(0), T(0));
                }
            } else {   // lines shorter than the rank count times the thread stride: plain per-element form
                FSM_UNROLL
                for (int m = 0; m < EPT; ++m)
                    raw[m] = ((keepmask >> m) & 1u) ? src[blk_off(tau + m * TL, ib, 1)] : mk<T>(T(0), T(0));
            }
        }
    };
    cplx<T> raw[EPT];
    load_line(0, raw);
    twiddles_ready();
    const int ecm = (EMODE != 0) ? ((1 << eb.cyc) - 1) : 0;
    const int es = (int)out_e_stride;     // one field of one rank block: fits 32 bits
    for (int j = 0; j < spec.nfo; ++j) {
        cplx<T> v[EPT];
        const bool deriv = spec.deriv[j] != 0;
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) v[m] = deriv ? cmul_i(raw[m], g.dk[1][tau + m * TL]) : raw[m];
        if (j + 1 < spec.nfo) load_line(j + 1, raw);
        cplx<T>* pbufs = bufs + (j & 1) * kKL * Cfg::LINE_PITCH;
        line_fft_head<Cfg, DIR, T>(v, pbufs + lt * Cfg::LINE_PITCH, tw, tau, sync);
        __syncthreads();
        cplx<T>* dst = out + (b * spec.nfo + j) * out_fstride + (long)o * out_o_stride + t0 + threadIdx.x % kKL;
        const bool valid = (int)(threadIdx.x % kKL) < k_valid;
        if constexpr (EMODE == 0) {
            FSM_PIN(dst);
            rotated_last_stage<Cfg, DIR, T>(pbufs, tw, [&](int e, cplx<T> val) FSM_INLINE_LAMBDA {
                if (valid) dst[e * es] = val;
            });
        } else if constexpr (EMODE == 2) {  // direct exchange: the ky-slab owner's receive buffer
            const long off0 = dst - out;
            rotated_last_stage<Cfg, DIR, T>(pbufs, tw, [&](int e, cplx<T> val) FSM_INLINE_LAMBDA {
                if (valid) *peer_ptr<T>(pe, eb, e, off0, out_e_stride) = val;
            });
        } else {                            // send buffer: rank block e & (P-1), slot e >> log2(P)
            rotated_last_stage<Cfg, DIR, T>(pbufs, tw, [&](int e, cplx<T> val) FSM_INLINE_LAMBDA {
                if (valid) dst[(long)(e & ecm) * eb.stride + (e >> eb.cyc) * es] = val;
            });
        }
    }
}

// ------------------------------------------------------------------------------------------
// Pass PHYS: last-axis pass. Two real fields ride in one complex transform:
//   inverse: Z = A_hat + i B_hat (mirrored/conjugated for the negative half) -> (a, b) = (Re, Im)
//   forward: z = s0 + i s1 -> Z; S0(k) = (Z(k)+conj Z(N-k))/2, S1(k) = (Z(k)-conj Z(N-k))/(2i)
// The nonlinear product is formed in registers between the two.
//   rows: 2-D x;  3-D (x, y) with y the line-minor index. Output rotated.
// ------------------------------------------------------------------------------------------
#ifndef FSM_PHYS_PARK
#define FSM_PHYS_PARK 1
#ifndef FSM_PHYS_PARK3
#define FSM_PHYS_PARK3 1
#endif
#endif
template <int PROG, int NDIM>
struct PhysTraits;
// NFI = input fields per sample, NOUT = output fields per sample, RPT = rows per thread-line
template <> struct PhysTraits<PROG_NS2D, 2> { static constexpr int NFI = 2, NOUT = 1, RPT = 2; };
template <> struct PhysTraits<PROG_KS, 2> { static constexpr int NFI = 2, NOUT = 1, RPT = 2; };
template <> struct PhysTraits<PROG_KS2D, 2> { static constexpr int NFI = 1, NOUT = 1, RPT = 2; };
template <> struct PhysTraits<PROG_KS, 3> { static constexpr int NFI = 3, NOUT = 1, RPT = 2; };
template <> struct PhysTraits<PROG_CONV, 2> { static constexpr int NFI = 4, NOUT = 2, RPT = 1; };
template <> struct PhysTraits<PROG_CONV, 3> { static constexpr int NFI = 9, NOUT = 3, RPT = 2; };
template <int NDIM> struct PhysTraits<PROG_C2R, NDIM> { static constexpr int NFI = 1, NOUT = 0, RPT = 2; };
template <int NDIM> struct PhysTraits<PROG_R2C, NDIM> { static constexpr int NFI = 0, NOUT = 1, RPT = 2; };

// Build the paired line Z(p) = A(p) + i B(p), p = tau + m*TL, from two half-lines (k <= kmax kept; the
// imaginary parts of the DC / Nyquist entries are dropped as a C2R transform does). The first half of a
// thread's elements is read directly (p < N/2), the second half mirrored and conjugated (k = N - p), so the
// index and conjugation selects are resolved at compile time; only tau == 0 touches p = 0 and p = N/2.
template <typename T, class Cfg>
__device__ __forceinline__ void pair_load_raw(cplx<T>* A, cplx<T>* B, const cplx<T>* a, const cplx<T>* b, int tau, int kmax,
                                              bool row_ok) {
    constexpr int N = Cfg::N, EPT = Cfg::EPT, TL = Cfg::TL;
    // element m reads k = p (first half) or k = N - p (mirrored half): two base pointers per field, constant offsets
    const bool ha = row_ok && a != nullptr, hb = row_ok && b != nullptr;
    const cplx<T>* alo = (a ? a : b) + tau;
    const cplx<T>* ahi = (a ? a : b) + (N - tau);
    const cplx<T>* blo = (b ? b : a) + tau;
    const cplx<T>* bhi = (b ? b : a) + (N - tau);
    FSM_PIN(alo); FSM_PIN(ahi); FSM_PIN(blo); FSM_PIN(bhi);   // (never dereferenced for a row beyond the end)
    static_for<0, EPT>([&](auto mc) FSM_INLINE_LAMBDA {
        constexpr int m = decltype(mc)::value;
        constexpr bool mirrored = (2 * m * TL >= N);
        const int p = tau + m * TL;
        const int k = mirrored ? N - p : p;
        const bool kept = k <= kmax;
        if constexpr (mirrored) {
            A[m] = (kept && ha) ? ahi[-m * TL] : mk<T>(T(0), T(0));
            B[m] = (kept && hb) ? bhi[-m * TL] : mk<T>(T(0), T(0));
        } else {
            A[m] = (kept && ha) ? alo[m * TL] : mk<T>(T(0), T(0));
            B[m] = (kept && hb) ? blo[m * TL] : mk<T>(T(0), T(0));
        }
    });
}
template <typename T, class Cfg>
__device__ __forceinline__ void pair_combine(cplx<T>* v, const cplx<T>* Ar, const cplx<T>* Br, bool da, bool db, const T* dk,
                                             int tau, int kmax) {
    constexpr int N = Cfg::N, EPT = Cfg::EPT, TL = Cfg::TL;
    static_for<0, EPT>([&](auto mc) FSM_INLINE_LAMBDA {
        constexpr int m = decltype(mc)::value;
        constexpr bool mirrored = (2 * m * TL >= N);          // p >= N/2 for every tau (p == N/2 only if tau == 0)
        const int p = tau + m * TL;
        const int k = mirrored ? N - p : p;
        cplx<T> A = Ar[m], B = Br[m];
        if (da || db) {
            const T d = (k <= kmax) ? dk[k] : T(0);
            if (da) A = cmul_i(A, d);
            if (db) B = cmul_i(B, d);
        }
        if constexpr (m == 0 || 2 * m * TL == N) {
            if (tau == 0) { A.y = T(0); B.y = T(0); }           // k == 0 or k == N/2
        }
        if constexpr (mirrored) { A.y = -A.y; B.y = -B.y; }
        v[m] = mk<T>(A.x - B.y, A.y + B.x);
    });
}
template <typename T, class Cfg>
__device__ __forceinline__ void pair_fill(cplx<T>* v, const cplx<T>* a, const cplx<T>* b, bool da, bool db, const T* dk,
                                          int tau, int kmax, bool row_ok) {
    cplx<T> A[Cfg::EPT], B[Cfg::EPT];
    pair_load_raw<T, Cfg>(A, B, a, b, tau, kmax, row_ok);
    pair_combine<T, Cfg>(v, A, B, da, db, dk, tau, kmax);
}

// NLP = thread-lines per CTA (kKL by default; a tuning parameter, see FSM_PHYS3D_NL512 in fsm_kernels.cu).
template <typename T, class Cfg, int PROG, int NDIM, int NLP = kKL>
__global__ void __launch_bounds__(NLP * Cfg::TL, FSM_MINB(NLP * Cfg::TL)) k_pass_phys(Geom<T> g, const cplx<T>* __restrict__ win, cplx<T>* __restrict__ wout,
                                                    const T* __restrict__ phys_in, T* __restrict__ phys_out,
                                                    long win_fstride, long wout_fstride, int K /*rows per CTA*/,
                                                    long in_t_stride, long in_o_stride, long out_o_stride,
                                                    long out_e_stride, int n_t, int nsamp_fields /*C2R/R2C: fields per z*/) {
    constexpr int N = Cfg::N, EPT = Cfg::EPT, TL = Cfg::TL;
    using PT = PhysTraits<PROG, NDIM>;
    constexpr int RPT = PT::RPT, NFI = PT::NFI, NOUT = PT::NOUT;
    constexpr int NFW = (NOUT * RPT + 1) / 2;  // forward transforms (= staging lines) per thread-line
    FSM_DYN_SMEM(smem_raw);
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* bufs = tw + Smem<Cfg, T>::TWPAD;
    constexpr int NL = NLP;  // thread-lines per CTA (K == NLP * RPT rows)
    const int lt = threadIdx.x / TL, tau = threadIdx.x % TL;
    const int kmaxl = (PROG == PROG_C2R) ? N / 2 : g.kmax[NDIM - 1];
    if constexpr (PROG == PROG_NS2D || PROG == PROG_KS2D) {
        if ((g.pf & PF_PHYS_SELF) && tau == 0) {
            // this thread-line's later rows: asked for now, read after the first transform(s)
            const cplx<T>* wb1 = win + (long)blockIdx.z * NFI * win_fstride + (long)blockIdx.y * in_o_stride;
            FSM_UNROLL
            for (int r = 0; r < RPT; ++r) {
                const int row1 = blockIdx.x * K + lt * RPT + r;
                if (row1 < n_t) {
                    FSM_UNROLL
                    for (int f = 0; f < NFI; ++f)
                        if (r + f > 0) l2_prefetch_line<T>(wb1 + f * win_fstride + (long)row1 * in_t_stride, N, kmaxl);
                }
            }
        }
        if ((g.pf & PF_PHYS) && tau == 0) {
            int x2, o2, z2;
            if (next_wave_block(g.pf_wave, x2, o2, z2)) {
                const cplx<T>* wb2 = win + (long)z2 * NFI * win_fstride + (long)o2 * in_o_stride;
                FSM_UNROLL
                for (int r = 0; r < RPT; ++r) {
                    const int row2 = x2 * K + lt * RPT + r;
                    if (row2 < n_t) {
                        FSM_UNROLL
                        for (int f = 0; f < NFI; ++f) l2_prefetch_line<T>(wb2 + f * win_fstride + (long)row2 * in_t_stride, N, kmaxl);
                    }
                }
            }
        }
    }
    twiddles_begin<Cfg, T>(tw);
    if (g.pf & PF_TW_EARLY) twiddles_ready();
    bool tw_pending = true;   // resolved at compile time: the code below is straight-line
    auto tw_ready_once = [&]() FSM_INLINE_LAMBDA {
        if (tw_pending) { twiddles_ready(); tw_pending = false; }
    };
    const int t0 = blockIdx.x * K;
    const int o = blockIdx.y;
    const long b = blockIdx.z;
    const T* dkl = g.dk[NDIM - 1];
    LineSync<TL> sync{1 + lt};
    // smem: NL fft buffers, then NL*NFW staging lines
    cplx<T>* mybuf = bufs + lt * Cfg::LINE_PITCH;
    cplx<T>* stage = bufs + (NL + lt * (NFW > 0 ? NFW : 1)) * Cfg::LINE_PITCH;
    const cplx<T>* wb = win + b * NFI * win_fstride + (long)o * in_o_stride;

    unsigned keepmask = 0;   // bit m: element p = tau + m*TL of a full complex row survives the dealiasing
    FSM_UNROLL
    for (int m = 0; m < EPT; ++m) {
        const int p = tau + m * TL;
        if (p <= kmaxl || p >= N - kmaxl) keepmask |= 1u << m;
    }
    T keep2[EPT];   // CONV3D: third component of row 0 waits for row 1
    cplx<T> rawA[(PROG == PROG_CONV && NDIM == 3) ? EPT : 1], rawB[(PROG == PROG_CONV && NDIM == 3) ? EPT : 1];
    (void)rawA; (void)rawB;
    T acc[(NOUT > 0 ? NOUT : 1)][EPT];
    (void)keep2;
    static_for<0, RPT>([&](auto rc) FSM_INLINE_LAMBDA {
        constexpr int r = decltype(rc)::value;
        const int row = t0 + lt * RPT + r;
        const bool row_ok = row < n_t;
        const long roff = (long)row * in_t_stride;
        cplx<T> v[EPT];
        auto inverse_pair = [&](int fa, int fb, bool da, bool db) FSM_INLINE_LAMBDA {
            const cplx<T>* pa = (fa >= 0) ? wb + fa * win_fstride + roff : nullptr;
            const cplx<T>* pb = (fb >= 0) ? wb + fb * win_fstride + roff : nullptr;
            pair_fill<T, Cfg>(v, pa, pb, da, db, dkl, tau, kmaxl, row_ok);
            tw_ready_once();
            sync();
            line_fft<Cfg, +1, T>(v, mybuf, tw, tau, sync);
        };
        if constexpr (PROG == PROG_NS2D || PROG == PROG_KS2D) {
            // Z-lines written by IX: NS2D field 0 = u_x + i d_x w, field 1 = u_y + i d_y w; KS2D field 0 =
            // phi_x + i phi_y (full complex rows)
            auto inverse_z = [&](int f) FSM_INLINE_LAMBDA {
                const cplx<T>* z = wb + f * win_fstride + (row_ok ? roff : 0) + tau;
                FSM_PIN(z);
                const unsigned km = row_ok ? keepmask : 0u;
                FSM_UNROLL
                for (int m = 0; m < EPT; ++m) v[m] = ((km >> m) & 1u) ? z[m * TL] : mk<T>(T(0), T(0));
                tw_ready_once();
                sync();
                line_fft<Cfg, +1, T>(v, mybuf, tw, tau, sync);
            };
            inverse_z(0);
            if constexpr (PROG == PROG_KS2D) {
                FSM_UNROLL
                for (int m = 0; m < EPT; ++m) acc[0][m] = T(0.5) * (v[m].x * v[m].x + v[m].y * v[m].y);
            } else {
                FSM_UNROLL
                for (int m = 0; m < EPT; ++m) acc[0][m] = v[m].x * v[m].y;
                inverse_z(1);
                FSM_UNROLL
                for (int m = 0; m < EPT; ++m) acc[0][m] += v[m].x * v[m].y;
            }
        } else if constexpr (PROG == PROG_KS && NDIM == 2) {
            // fields: 0 phi (x-transformed), 1 d_x phi  ->  1/2 (phi_x^2 + phi_y^2)
            inverse_pair(1, 0, false, true);
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) acc[0][m] = T(0.5) * (v[m].x * v[m].x + v[m].y * v[m].y);
        } else if constexpr (PROG == PROG_KS && NDIM == 3) {
            // fields: 0 phi, 1 d_x phi, 2 d_y phi
            inverse_pair(1, 2, false, false);
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) acc[0][m] = v[m].x * v[m].x + v[m].y * v[m].y;
            inverse_pair(0, -1, true, false);
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) acc[0][m] = T(0.5) * (acc[0][m] + v[m].x * v[m].x);
        } else if constexpr (PROG == PROG_CONV && NDIM == 2) {
            // fields per channel c: 2c = u_c, 2c+1 = d_x u_c ; d_y applied here
            T u0[EPT], u1[EPT];
            inverse_pair(0, 2, false, false);
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) { u0[m] = v[m].x; u1[m] = v[m].y; }
            inverse_pair(1, 3, false, false);
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) { acc[0][m] = u0[m] * v[m].x; acc[1][m] = u0[m] * v[m].y; }
            inverse_pair(0, 2, true, true);
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) { acc[0][m] += u1[m] * v[m].x; acc[1][m] += u1[m] * v[m].y; }
        } else if constexpr (PROG == PROG_CONV && NDIM == 3) {
            // fields: c = u_c (0..2), 3+c = d_x u_c, 6+c = d_y u_c ; d_z applied here. Six paired transforms per
            // row: for direction j  (d_j u_2, u_j)  then  (d_j u_0, d_j u_1). The half-lines of transform i+1
            // are loaded (registers) before transform i runs, so the load latency hides behind the butterflies.
            constexpr int FA[6] = {5, 3, 8, 6, 2, 0}, FB[6] = {0, 4, 1, 7, 2, 1};
            constexpr bool DA[6] = {false, false, false, false, true, true}, DB[6] = {false, false, false, false, false, true};
            T uj[EPT];
            if constexpr (r == 0) pair_load_raw<T, Cfg>(rawA, rawB, wb + FA[0] * win_fstride + roff, wb + FB[0] * win_fstride + roff,
                                                        tau, kmaxl, row_ok);
            tw_ready_once();
            static_for<0, 6>([&](auto ic) FSM_INLINE_LAMBDA {
                constexpr int i = decltype(ic)::value;
                pair_combine<T, Cfg>(v, rawA, rawB, DA[i], DB[i], dkl, tau, kmaxl);
                if constexpr (i + 1 < 6) {
                    pair_load_raw<T, Cfg>(rawA, rawB, wb + FA[i + 1] * win_fstride + roff, wb + FB[i + 1] * win_fstride + roff,
                                          tau, kmaxl, row_ok);
                } else if constexpr (r + 1 < RPT) {
                    const int row2 = row + 1;
                    const long roff2 = (long)row2 * in_t_stride;
                    pair_load_raw<T, Cfg>(rawA, rawB, wb + FA[0] * win_fstride + roff2, wb + FB[0] * win_fstride + roff2, tau,
                                          kmaxl, row2 < n_t);
                }
                sync();
                line_fft<Cfg, +1, T>(v, mybuf, tw, tau, sync);
                constexpr int j = i / 2;
                if constexpr ((i & 1) == 0) {
                    FSM_UNROLL
                    for (int m = 0; m < EPT; ++m) {
                        uj[m] = v[m].y;
                        acc[2][m] = (j == 0) ? uj[m] * v[m].x : acc[2][m] + uj[m] * v[m].x;
                    }
                } else {
                    FSM_UNROLL
                    for (int m = 0; m < EPT; ++m) {
                        acc[0][m] = (j == 0) ? uj[m] * v[m].x : acc[0][m] + uj[m] * v[m].x;
                        acc[1][m] = (j == 0) ? uj[m] * v[m].y : acc[1][m] + uj[m] * v[m].y;
                    }
                }
            });
        } else if constexpr (PROG == PROG_C2R) {
            // one field, rows r (real part) and r+1 (imag part) of the same field share a transform
        } else if constexpr (PROG == PROG_R2C) {
            const T* prow = phys_in + (b * (long)n_t * gridDim.y + ((long)o * n_t + row)) * N;
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) acc[0][m] = row_ok ? prow[tau + m * TL] : T(0);
            tw_ready_once();
        }

        // ---- forward side: pack real results pairwise into complex lines and transform
        if constexpr (NOUT == 1) {
            // the product of row 0 waits for row 1: with 16 elements per thread it is parked in this thread-line's
            // (still unused) staging line instead of 16 registers, which the two inverse transforms of row 1 need
            constexpr bool kPark = (EPT >= 16) && FSM_PHYS_PARK;
            cplx<T>* park = stage;   // two values per 2*sizeof(T) slot
            if constexpr (r == 0) {
                FSM_UNROLL
                for (int m = 0; m < EPT; m += 2) {
                    if constexpr (kPark) park[tau + (m / 2) * TL] = mk<T>(acc[0][m], acc[0][m + 1]);
                    else { keep2[m] = acc[0][m]; keep2[m + 1] = acc[0][m + 1]; }
                }
            } else {
                FSM_UNROLL
                for (int m = 0; m < EPT; m += 2) {
                    if constexpr (kPark) {
                        const cplx<T> pk2 = park[tau + (m / 2) * TL];
                        v[m] = mk<T>(pk2.x, acc[0][m]);
                        v[m + 1] = mk<T>(pk2.y, acc[0][m + 1]);
                    } else {
                        v[m] = mk<T>(keep2[m], acc[0][m]);
                        v[m + 1] = mk<T>(keep2[m + 1], acc[0][m + 1]);
                    }
                }
                sync();
                line_fft<Cfg, -1, T>(v, mybuf, tw, tau, sync);
                FSM_UNROLL
                for (int m = 0; m < EPT; ++m) stage[tau + m * TL] = v[m];
            }
        } else if constexpr (NOUT == 2) {
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) v[m] = mk<T>(acc[0][m], acc[1][m]);
            sync();
            line_fft<Cfg, -1, T>(v, mybuf, tw, tau, sync);
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) stage[tau + m * TL] = v[m];
        } else if constexpr (NOUT == 3) {
            // staging line r*? : line 0 = (c0,c1) of row 0, line 1 = (c2 row0, c2 row1), line 2 = (c0,c1) of row 1
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) v[m] = mk<T>(acc[0][m], acc[1][m]);
            sync();
            line_fft<Cfg, -1, T>(v, mybuf, tw, tau, sync);
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) stage[(r == 0 ? 0 : 2) * Cfg::LINE_PITCH + tau + m * TL] = v[m];
            // the third component of row 0 waits for row 1: parked in staging line 1 (written only after row 1) as real
            // values instead of EPT registers held across the six inverse transforms of row 1 (the kernel sits at the
            // 128-register cap: cuobjdump showed 72 B of spills without this)
            // (measured: C5, 512-point lines, last-axis pass 9.98 -> 9.82 ms per step; at 256 points the registers are
            // there and parking costs 5 % of the pass, so shorter lines keep the registers)
            constexpr bool kPark3 = FSM_PHYS_PARK3 && (N >= 512);
            T* park3 = reinterpret_cast<T*>(stage + 1 * Cfg::LINE_PITCH);
            if constexpr (r == 0) {
                FSM_UNROLL
                for (int m = 0; m < EPT; ++m) {
                    if constexpr (kPark3) park3[tau + m * TL] = acc[2][m];
                    else keep2[m] = acc[2][m];
                }
            } else {
                FSM_UNROLL
                for (int m = 0; m < EPT; ++m) v[m] = mk<T>(kPark3 ? park3[tau + m * TL] : keep2[m], acc[2][m]);
                sync();
                line_fft<Cfg, -1, T>(v, mybuf, tw, tau, sync);
                FSM_UNROLL
                for (int m = 0; m < EPT; ++m) stage[1 * Cfg::LINE_PITCH + tau + m * TL] = v[m];
            }
        }
    });

    if constexpr (PROG == PROG_C2R) {
        // rows (lt*2, lt*2+1) of field f = blockIdx.z % nsamp_fields handled as one complex inverse transform
        const int row = t0 + lt * 2;
        const bool ok0 = row < n_t, ok1 = row + 1 < n_t;
        const cplx<T>* p0 = win + b * win_fstride + (long)o * in_o_stride + (long)row * in_t_stride;
        cplx<T> v[EPT];
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) {
            const int p = tau + m * TL;
            const int k = (p <= N / 2) ? p : N - p;
            cplx<T> A = ok0 ? p0[k] : mk<T>(T(0), T(0));
            cplx<T> B = ok1 ? p0[in_t_stride + k] : mk<T>(T(0), T(0));
            if (k == 0 || k == N / 2) { A.y = T(0); B.y = T(0); }
            if (p > N / 2) { A.y = -A.y; B.y = -B.y; }
            v[m] = mk<T>(A.x - B.y, A.y + B.x);
        }
        tw_ready_once();
        line_fft<Cfg, +1, T>(v, mybuf, tw, tau, sync);
        T* prow = phys_out + (b * (long)n_t * gridDim.y + ((long)o * n_t + row)) * N;
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) {
            if (ok0) prow[tau + m * TL] = v[m].x;
            if (ok1) prow[N + tau + m * TL] = v[m].y;
        }
        (void)nsamp_fields;
        return;
    }

    if constexpr (NOUT > 0) {
        __syncthreads();
        // ---- split the packed spectra and store rotated: wout[f][k * out_e_stride + row]
        // Thread (l = tid % kKL, k0 = tid / kKL) handles line slot l and the modes k0 + it*TL < N/2 (compile-time
        // trip count, constant strides); the kKL Nyquist modes k = N/2 are a tail for the first kKL threads.
        cplx<T>* ob = wout + b * NOUT * wout_fstride + (long)o * out_o_stride + t0;
        FSM_PIN(ob);
        const cplx<T>* st0 = bufs + NL * Cfg::LINE_PITCH;
        constexpr int NLc = NLP;
        const int es = (int)out_e_stride;   // offsets inside one field fit 32 bits
        auto emit = [&](int l, int k, int kn) FSM_INLINE_LAMBDA {
            const cplx<T>* sl = st0 + l * NFW * Cfg::LINE_PITCH;
            auto split = [&](const cplx<T>* line, cplx<T>& s0, cplx<T>& s1) FSM_INLINE_LAMBDA {
                const cplx<T> zk = line[k], zn = line[kn];
                s0 = mk<T>(T(0.5) * (zk.x + zn.x), T(0.5) * (zk.y - zn.y));
                s1 = mk<T>(T(0.5) * (zk.y + zn.y), T(0.5) * (zn.x - zk.x));
            };
            cplx<T> s0, s1;
            cplx<T>* dst = ob + k * es;
            if constexpr (NOUT == 1) {
                split(sl, s0, s1);
                const int row = l * 2;
                if (t0 + row < n_t) dst[row] = s0;
                if (t0 + row + 1 < n_t) dst[row + 1] = s1;
            } else if constexpr (NOUT == 2) {
                split(sl, s0, s1);
                if (t0 + l < n_t) {
                    dst[l] = s0;
                    dst[wout_fstride + l] = s1;
                }
            } else {
                const int row = l * 2;
                const bool ok0 = t0 + row < n_t, ok1 = t0 + row + 1 < n_t;
                split(sl, s0, s1);
                if (ok0) { dst[row] = s0; dst[wout_fstride + row] = s1; }
                split(sl + 2 * Cfg::LINE_PITCH, s0, s1);
                if (ok1) { dst[row + 1] = s0; dst[wout_fstride + row + 1] = s1; }
                split(sl + Cfg::LINE_PITCH, s0, s1);
                if (ok0) dst[2 * wout_fstride + row] = s0;
                if (ok1) dst[2 * wout_fstride + row + 1] = s1;
            }
        };
        {
            const int l = threadIdx.x % NLc, k0 = threadIdx.x / NLc;   // k0 < TL
            static_for<0, EPT / 2>([&](auto itc) FSM_INLINE_LAMBDA {
                constexpr int it = decltype(itc)::value;
                const int k = k0 + it * TL;
                if constexpr (it == 0) emit(l, k, (k == 0) ? 0 : N - k);
                else emit(l, k, N - k);
            });
            if (threadIdx.x < NLc) emit(threadIdx.x, N / 2, N / 2);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Pass PHYS-Z: last-axis pass of the Z-line programs (NS2D, KS2D), organised for phase overlap and for the
// shared-memory pipe (ncu, round 1: the generic kernel above ran "DRAM phase + shared-memory phase" back to
// back in the ONE 512-thread CTA an SM could hold -- 152 us + 225 us of the 388 us it took on C3).
//   * NLZ thread-lines per CTA, each taking FOUR rows -> at 1024 points 256 threads and 86 KB of shared memory
//     per CTA: two independent CTAs per SM whose load / transform / store phases interleave;
//   * rows (4 lt + 2h, 4 lt + 2h + 1) ride as one packed forward transform whose head lands in head buffer
//     2 lt + h; the LAST stage of all 2 NLZ heads runs in the rotated distribution (lanes across buffers), each
//     thread taking the work items j and NS - j so that Z(k) and Z(N - k) meet in registers: the split
//     S0 = (Z(k) + conj Z(N-k))/2, S1 = (Z(k) - conj Z(N-k))/2i needs no staging line, and the two rows of a
//     buffer leave as ONE 16-byte store, 128-byte segments per k across the lanes.
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void store_pair(cplx<T>* dst, cplx<T> a, cplx<T> b) {
    dst[0] = a;
    dst[1] = b;
}
#if defined(__CUDACC__) && !defined(FSM_EMU)
__device__ __forceinline__ void store_pair(cplx<float>* dst, cplx<float> a, cplx<float> b) {
    *reinterpret_cast<float4*>(dst) = make_float4(a.x, a.y, b.x, b.y);   // dst is 16-byte aligned (even row)
}
#endif

template <typename T, class Cfg, int PROG>
__global__ void __launch_bounds__(PhysZ<Cfg>::NT, FSM_MINB(PhysZ<Cfg>::NT))
k_pass_physz(Geom<T> g, const cplx<T>* __restrict__ win, cplx<T>* __restrict__ wout, long win_fstride, long wout_fstride,
             long in_t_stride, long out_e_stride, int n_t) {
    static_assert(PROG == PROG_NS2D || PROG == PROG_KS2D, "Z-line programs only");
    constexpr int N = Cfg::N, EPT = Cfg::EPT, TL = Cfg::TL;
    using Z = PhysZ<Cfg>;
    constexpr int NLZ = Z::NLZ, KS = Z::KS, K = Z::K;
    constexpr int NFI = (PROG == PROG_NS2D) ? 2 : 1;
    constexpr int RL = LastStage<Cfg>::RL, NS = LastStage<Cfg>::NS;
    static_assert(Cfg::NST >= 2 && TL % 2 == 0 && (NS / 2) % (TL / 2) == 0, "rotated paired last stage");
    FSM_DYN_SMEM(smem_raw);
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* hb = tw + Smem<Cfg, T>::TWPAD;
    const int lt = threadIdx.x / TL, tau = threadIdx.x % TL;
    twiddles_begin<Cfg, T>(tw);
    const int t0 = blockIdx.x * K;
    const long b = blockIdx.z;
    const int kmaxl = g.kmax[1];
    LineSync<TL> sync{1 + lt};
    cplx<T>* bufB = hb + (2 * lt) * Cfg::LINE_PITCH;        // head of rows 0,1 of this thread-line
    cplx<T>* bufA = bufB + Cfg::LINE_PITCH;                 // exchange buffer of the inverse transforms, head of rows 2,3
    cplx<T>* park = hb + KS * Cfg::LINE_PITCH + lt * Z::PARK_PITCH;
    const cplx<T>* wb = win + b * NFI * win_fstride;

    unsigned keepmask = 0;   // bit m: element p = tau + m*TL of a full complex row survives the dealiasing
    FSM_UNROLL
    for (int m = 0; m < EPT; ++m) {
        const int p = tau + m * TL;
        if (p <= kmaxl || p >= N - kmaxl) keepmask |= 1u << m;
    }
    bool tw_pending = true;
    T keep2[Z::PARK ? 1 : EPT];
    (void)keep2;
    static_for<0, 4>([&](auto rc) FSM_INLINE_LAMBDA {
        constexpr int r = decltype(rc)::value;
        const int row = t0 + 4 * lt + r;
        const bool row_ok = row < n_t;
        cplx<T> v[EPT];
        T acc[EPT];
        auto inverse_z = [&](int f) FSM_INLINE_LAMBDA {
            const cplx<T>* z = wb + f * win_fstride + (row_ok ? (long)row * in_t_stride : 0) + tau;
            FSM_PIN(z);
            const unsigned km = row_ok ? keepmask : 0u;
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) v[m] = ((km >> m) & 1u) ? z[m * TL] : mk<T>(T(0), T(0));
            if (tw_pending) { twiddles_ready(); tw_pending = false; }
            sync();
            line_fft<Cfg, +1, T>(v, bufA, tw, tau, sync);
        };
        inverse_z(0);
        if constexpr (PROG == PROG_KS2D) {
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) acc[m] = T(0.5) * (v[m].x * v[m].x + v[m].y * v[m].y);
        } else {
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) acc[m] = v[m].x * v[m].y;
            inverse_z(1);
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) acc[m] += v[m].x * v[m].y;
        }
        if constexpr ((r & 1) == 0) {
            FSM_UNROLL
            for (int m = 0; m < EPT; m += 2) {
                if constexpr (Z::PARK) park[tau + (m / 2) * TL] = mk<T>(acc[m], acc[m + 1]);
                else { keep2[m] = acc[m]; keep2[m + 1] = acc[m + 1]; }
            }
        } else {
            FSM_UNROLL
            for (int m = 0; m < EPT; m += 2) {
                if constexpr (Z::PARK) {
                    const cplx<T> pk2 = park[tau + (m / 2) * TL];
                    v[m] = mk<T>(pk2.x, acc[m]);
                    v[m + 1] = mk<T>(pk2.y, acc[m + 1]);
                } else {
                    v[m] = mk<T>(keep2[m], acc[m]);
                    v[m + 1] = mk<T>(keep2[m + 1], acc[m + 1]);
                }
            }
            sync();   // every thread of the line is done reading bufA (last stage of the inverse transform)
            line_fft_head<Cfg, -1, T>(v, (r == 1) ? bufB : bufA, tw, tau, sync);
        }
    });
    __syncthreads();

    // ---- rotated, paired last stage of the KS head buffers: split and store wout[k * out_e_stride + row]
    const int s = threadIdx.x % KS, widx = threadIdx.x / KS;     // widx < TL / 2
    const cplx<T>* buf = hb + s * Cfg::LINE_PITCH;
    const bool ok0 = t0 + 2 * s < n_t, ok1 = t0 + 2 * s + 1 < n_t;
    cplx<T>* ob = wout + b * wout_fstride + t0 + 2 * s;
    FSM_PIN(ob);
    const int es = (int)out_e_stride;   // offsets inside one field fit 32 bits
    auto emit = [&](int k, cplx<T> zk, cplx<T> zn) FSM_INLINE_LAMBDA {
        const cplx<T> s0 = mk<T>(T(0.5) * (zk.x + zn.x), T(0.5) * (zk.y - zn.y));
        const cplx<T> s1 = mk<T>(T(0.5) * (zk.y + zn.y), T(0.5) * (zn.x - zk.x));
        cplx<T>* dst = ob + k * es;
        if (ok1) store_pair(dst, s0, s1);
        else if (ok0) dst[0] = s0;
    };
    static_for<0, EPT / RL>([&](auto qc) FSM_INLINE_LAMBDA {
        constexpr int q = decltype(qc)::value;
        const int u = widx + q * (TL / 2);                       // pair unit: work items u and NS - u
        const int w1 = u, w2 = (u == 0) ? NS / 2 : NS - u;
        cplx<T> a[RL], c[RL];
        fft_last_item<Cfg, -1, T, 0>(buf, tw, w1, a);            // a[t] = X[w1 + t*NS]
        fft_last_item<Cfg, -1, T, 0>(buf, tw, w2, c);            // c[t] = X[w2 + t*NS]
        if (q == 0 && u == 0) {
            // item 0 pairs with itself: k = t*NS <-> N - k = (RL - t)*NS; item NS/2 likewise: t <-> RL-1-t
            static_for<0, RL / 2 + 1>([&](auto tc) FSM_INLINE_LAMBDA {
                constexpr int t = decltype(tc)::value;
                emit(t * NS, a[t], a[(RL - t) % RL]);
            });
            static_for<0, RL / 2>([&](auto tc) FSM_INLINE_LAMBDA {
                constexpr int t = decltype(tc)::value;
                emit(NS / 2 + t * NS, c[t], c[RL - 1 - t]);
            });
        } else {
            static_for<0, RL / 2>([&](auto tc) FSM_INLINE_LAMBDA {
                constexpr int t = decltype(tc)::value;
                emit(w1 + t * NS, a[t], c[RL - 1 - t]);
                emit(w2 + t * NS, c[t], a[RL - 1 - t]);
            });
        }
    });
}

// ------------------------------------------------------------------------------------------
// Pass FX: forward transform along x of the C nonlinear components, spectral epilogue
// (coefficient, NS pressure projection, constant source, KS zero-mode capture) and the
// data-driven integrator combine. Contiguous in, contiguous out (rot-half layout).
// ------------------------------------------------------------------------------------------
#define FSM_MAX_IN 3
#define FSM_MAX_OUT 3
#define FSM_MAX_TAB 4
template <typename T>
struct Combine {
    int n_in, n_out, n_tab;
    int use_fresh;                 // 0: no nonlinear term (pure linear step)
    const cplx<T>* in[FSM_MAX_IN];
    cplx<T>* out[FSM_MAX_OUT];
    const T* tab[FSM_MAX_TAB];     // real tables [tab_channels][nmodes]
    long tab_cstride;              // 0 if one table serves every channel
    long tab_bstride;              // 0 if one table serves every sample (else per-sample tables: batched coefficients,
                                   // operator/_base.py:339-357 with tensor-valued coefficients)
    // row[r] = sum_m (ca + cb * tab[ct] + cb2 * tab[ct2])[r][m] * X_m,  X_0 = fresh N, X_{1+i} = in[i].
    // Rows 0..n_out-1 are stored to out[r]; row FSM_MAX_OUT (if has_next) is the next stage state, handed
    // in registers to the fused inverse transforms and never stored.
    int has_next;
    int kind;                                   // index of the matching CombineShape (compile-time structure), -1 = generic
    int tab_cplx;                               // tables hold cplx<T> entries (1-D grids, complex linear symbol)
    int any_ct2;                                // some coefficient uses a second table
    int ct[FSM_MAX_OUT + 1][FSM_MAX_IN + 1];    // table index, -1 = scalar only, -2 = term absent
    int ct2[FSM_MAX_OUT + 1][FSM_MAX_IN + 1];   // optional second table, -1 = none
    T ca[FSM_MAX_OUT + 1][FSM_MAX_IN + 1];
    T cb[FSM_MAX_OUT + 1][FSM_MAX_IN + 1];
    T cb2[FSM_MAX_OUT + 1][FSM_MAX_IN + 1];
};

template <typename T>
struct FxEpilogue {
    T nl_coef;                     // scalar coefficient of the convective term
    const T* nl_coef_b;            // optional per-sample coefficient [B] (tensor-valued coefficient on the nonlinear term,
                                   // operator/_base.py:375-403: `result += coef * fun(...)`); replaces nl_coef
    const cplx<T>* source;         // optional constant source spectrum [C][nmodes] (coef folded in)
    T* dc_out;                     // KS: per-sample zero-mode of the nonlinear term (captured, then zeroed)
    int project;                   // NS pressure projection (2-D and 3-D velocity form)
    const cplx<T>* force;          // optional constant term added BEFORE the projection [C][nmodes]: -coef * f_hat of
                                   // NSPressureConvection(external_force) (_navier_stokes.py:237-254)
    const cplx<T>* force_dyn;      // optional per-evaluation force spectrum f_hat(u) [B][C][nmodes] (a force operator that
                                   // depends on the state, evaluated by the caller on the stage input): -coef * f before the
                                   // projection and +coef * f after it (the reference adds the force twice, :241-254)
};

// ---- compile-time combine structures ---------------------------------------------------------------------
// The stage programs of the ETD integrators (fsm_plan.cu: build_stages) come in a handful of shapes. Which
// operand and which table slot feeds which output row is then known at compile time and the data-driven
// selects of the generic path fold away; the numeric coefficients (ca, cb) stay run-time data. The host
// (match_combine_shape) picks the shape; anything else (RK4, rhs, linear-only, fused) runs the generic path.
// ct(r, m): table slot of the coefficient of operand m (0 = fresh nonlinear term, 1.. = in[m-1]) in output
// row r; -1 = scalar only; -2 = term absent. Slots are numbered in order of appearance (make_combine).
constexpr int kCombineShapes = 7;
template <int KIND>
struct CShape {   // generic: structure read from the Combine at run time
    static constexpr int n_in = FSM_MAX_IN, n_out = FSM_MAX_OUT, n_tab = FSM_MAX_TAB;
    static FSM_HD constexpr int ct(int, int) { return -2; }
};
#define FSM_CSHAPE(K, NIN, NOUT, NTAB, ...)                              \
    template <>                                                          \
    struct CShape<K> {                                                   \
        static constexpr int n_in = NIN, n_out = NOUT, n_tab = NTAB;     \
        static FSM_HD constexpr int ct(int r, int m) {                   \
            constexpr int t[FSM_MAX_OUT][FSM_MAX_IN + 1] = __VA_ARGS__;  \
            return t[r][m];                                              \
        }                                                                \
    }
// u' = c1 N + E u                                               (ETDRK1 / SETDRK1)
FSM_CSHAPE(0, 1, 1, 2, {{0, 1, -2, -2}, {-2, -2, -2, -2}, {-2, -2, -2, -2}});
// a = c1 N + E u ; d = (c1 - c2) N + E u                        (ETDRK2 / SETDRK2 stage 1; stage 2 is shape 6, u' = c2 N + d)
FSM_CSHAPE(1, 1, 2, 3, {{0, 1, -2, -2}, {2, 1, -2, -2}, {-2, -2, -2, -2}});
// u' = c2 N + a - c2 N0                                         (ETDRK2 / SETDRK2 stage 2 when no c1 - c2 table is supplied)
FSM_CSHAPE(2, 2, 1, 1, {{0, -1, 0, -2}, {-2, -2, -2, -2}, {-2, -2, -2, -2}});
// a = c1 N + E2 u ; N0 = N ; sum = c4 N + E u                   (SETDRK3 / SETDRK4 stage 1)
FSM_CSHAPE(3, 1, 3, 4, {{0, 1, -2, -2}, {-1, -2, -2, -2}, {2, 3, -2, -2}});
// b = c2 N + E2 u ; sum += 2 c5 N                               (SETDRK4 stage 2)
FSM_CSHAPE(4, 2, 2, 3, {{0, 1, -2, -2}, {2, -2, -1, -2}, {-2, -2, -2, -2}});
// c = 2 c3 N + E2 a - c3 N0 ; sum += 2 c5 N                     (SETDRK4 stage 3, SETDRK3 stage 2)
FSM_CSHAPE(5, 3, 2, 3, {{0, 1, 0, -2}, {2, -2, -2, -1}, {-2, -2, -2, -2}});
// u' = c6 N + sum                                               (last stage of SETDRK3 / SETDRK4)
FSM_CSHAPE(6, 1, 1, 1, {{0, -1, -2, -2}, {-2, -2, -2, -2}, {-2, -2, -2, -2}});
#undef FSM_CSHAPE

// host and device: does this Combine have the structure of shape KIND?
template <int KIND, typename T>
FSM_HD inline bool combine_matches(const Combine<T>& cb) {
    using S = CShape<KIND>;
    if (cb.n_in != S::n_in || cb.n_out != S::n_out || cb.n_tab != S::n_tab || cb.has_next || cb.any_ct2 || !cb.use_fresh || cb.tab_cplx)
        return false;
    for (int r = 0; r < FSM_MAX_OUT; ++r)
        for (int m = 0; m < FSM_MAX_IN + 1; ++m)
            if (cb.ct[r][m] != S::ct(r, m)) return false;
    return cb.ct[FSM_MAX_OUT][0] == -2 || !cb.has_next;
}
template <typename T>
inline int match_combine_shape(const Combine<T>& cb) {
    if (combine_matches<0>(cb)) return 0;
    if (combine_matches<1>(cb)) return 1;
    if (combine_matches<2>(cb)) return 2;
    if (combine_matches<3>(cb)) return 3;
    if (combine_matches<4>(cb)) return 4;
    if (combine_matches<5>(cb)) return 5;
    if (combine_matches<6>(cb)) return 6;
    return -1;
}
// run f(integral_constant<int, KIND>) for the run-time kind (uniform across the grid)
template <class F>
__device__ __forceinline__ void with_combine_kind(int kind, F&& f) {
    switch (kind) {
        case 0: f(std::integral_constant<int, 0>{}); break;
        case 1: f(std::integral_constant<int, 1>{}); break;
        case 2: f(std::integral_constant<int, 2>{}); break;
        case 3: f(std::integral_constant<int, 3>{}); break;
        case 4: f(std::integral_constant<int, 4>{}); break;
        case 5: f(std::integral_constant<int, 5>{}); break;
        case 6: f(std::integral_constant<int, 6>{}); break;
        default: f(std::integral_constant<int, -1>{}); break;
    }
}

// Operands of NB modes (mode0 + j*mstride) of one combine: every global load is issued up front so the
// memory system sees them all in flight.
template <typename T, int NB>
struct CombineOperands {
    cplx<T> X[FSM_MAX_IN][NB];
    T tv[FSM_MAX_TAB][NB];
};
template <typename T, int NB, int KIND = -1>
__device__ __forceinline__ void combine_load(const Combine<T>& cb, long bc_off, long tab_off, long mode0, long mstride,
                                             CombineOperands<T, NB>& op) {
    using S = CShape<KIND>;
    FSM_UNROLL
    for (int i = 0; i < FSM_MAX_IN; ++i)
        if ((KIND >= 0) ? (i < S::n_in) : (i < cb.n_in)) {
            const cplx<T>* src = cb.in[i] + bc_off + mode0;
            FSM_UNROLL
            for (int j = 0; j < NB; ++j) op.X[i][j] = src[j * mstride];
        }
    FSM_UNROLL
    for (int q = 0; q < FSM_MAX_TAB; ++q)
        if ((KIND >= 0) ? (q < S::n_tab) : (q < cb.n_tab)) {
            const T* src = cb.tab[q] + tab_off + mode0;
            FSM_UNROLL
            for (int j = 0; j < NB; ++j) op.tv[q][j] = src[j * mstride];
        }
}
template <typename T, int NB, int KIND = -1>
__device__ __forceinline__ void combine_apply(const Combine<T>& cb, const cplx<T>* fresh, const CombineOperands<T, NB>& op,
                                              long bc_off, long mode0, long mstride, cplx<T>* next = nullptr) {
    using S = CShape<KIND>;
    FSM_UNROLL
    for (int r = 0; r < FSM_MAX_OUT + 1; ++r) {
        const bool is_next = (r == FSM_MAX_OUT);
        bool row_on;
        if constexpr (KIND >= 0) row_on = !is_next && r < S::n_out;
        else row_on = is_next ? (cb.has_next && next != nullptr) : (r < cb.n_out);
        if (row_on) {
            cplx<T> s[NB];
            FSM_UNROLL
            for (int j = 0; j < NB; ++j) s[j] = mk<T>(T(0), T(0));
            FSM_UNROLL
            for (int m = 0; m < FSM_MAX_IN + 1; ++m) {
                const int ti = (KIND >= 0) ? S::ct(r < FSM_MAX_OUT ? r : 0, m) : cb.ct[r][m];
                if (ti != -2) {
                    const int ti2 = (KIND >= 0) ? -1 : cb.ct2[r][m];
                    FSM_UNROLL
                    for (int j = 0; j < NB; ++j) {
                        T coef = cb.ca[r][m];
                        FSM_UNROLL
                        for (int q = 0; q < FSM_MAX_TAB; ++q)
                            if (ti == q) coef = fsm_fma(cb.cb[r][m], op.tv[q][j], coef);
                        if (KIND < 0 && cb.any_ct2) {
                            FSM_UNROLL
                            for (int q = 0; q < FSM_MAX_TAB; ++q)
                                if (ti2 == q) coef = fsm_fma(cb.cb2[r][m], op.tv[q][j], coef);
                        }
                        const cplx<T> x = (m == 0) ? fresh[j] : op.X[(m == 0) ? 0 : m - 1][j];
                        s[j] = cfma_s(coef, x, s[j]);
                    }
                }
            }
            if (is_next) {
                FSM_UNROLL
                for (int j = 0; j < NB; ++j) next[j] = s[j];
            } else {
                cplx<T>* dst = cb.out[r] + bc_off + mode0;
                FSM_UNROLL
                for (int j = 0; j < NB; ++j) dst[j * mstride] = s[j];
            }
        }
    }
}
template <typename T, int NB, int KIND = -1>
__device__ __forceinline__ void combine_block(const Combine<T>& cb, const cplx<T>* fresh, long bc_off, long tab_off,
                                              long mode0, long mstride, cplx<T>* next = nullptr, bool store = true) {
    (void)store;
    CombineOperands<T, NB> op;
    combine_load<T, NB, KIND>(cb, bc_off, tab_off, mode0, mstride, op);
    combine_apply<T, NB, KIND>(cb, fresh, op, bc_off, mode0, mstride, next);
}

// One mode with COMPLEX coefficient tables (odd-order linear terms make exp(L dt) complex; 1-D kernels and the
// point-wise linear step only — the fused 2-D/3-D epilogue keeps real tables).
template <typename T>
__device__ __forceinline__ void combine_mode_cplx(const Combine<T>& cb, cplx<T> fresh, long bc_off, long tab_off, long mode) {
    cplx<T> X[FSM_MAX_IN + 1];
    cplx<T> tv[FSM_MAX_TAB];
    X[0] = fresh;
    FSM_UNROLL
    for (int i = 0; i < FSM_MAX_IN; ++i) X[i + 1] = (i < cb.n_in) ? cb.in[i][bc_off + mode] : mk<T>(T(0), T(0));
    FSM_UNROLL
    for (int q = 0; q < FSM_MAX_TAB; ++q)
        tv[q] = (q < cb.n_tab) ? reinterpret_cast<const cplx<T>*>(cb.tab[q])[tab_off + mode] : mk<T>(T(0), T(0));
    FSM_UNROLL
    for (int r = 0; r < FSM_MAX_OUT; ++r) {
        if (r < cb.n_out) {
            cplx<T> s = mk<T>(T(0), T(0));
            FSM_UNROLL
            for (int m = 0; m < FSM_MAX_IN + 1; ++m) {
                const int ti = cb.ct[r][m], ti2 = cb.ct2[r][m];
                if (ti != -2) {
                    cplx<T> coef = mk<T>(cb.ca[r][m], T(0));
                    FSM_UNROLL
                    for (int q = 0; q < FSM_MAX_TAB; ++q) {
                        if (ti == q) coef = coef + cscale(tv[q], cb.cb[r][m]);
                        if (ti2 == q) coef = coef + cscale(tv[q], cb.cb2[r][m]);
                    }
                    s = s + cmul(coef, X[m]);
                }
            }
            cb.out[r][bc_off + mode] = s;
        }
    }
}

template <typename T>
__device__ __forceinline__ void combine_mode(const Combine<T>& cb, cplx<T> fresh, long bc_off, long tab_off, long mode) {
    if (cb.tab_cplx) combine_mode_cplx<T>(cb, fresh, bc_off, tab_off, mode);
    else combine_block<T, 1>(cb, &fresh, bc_off, tab_off, mode, 0);
}

template <typename T, class Cfg, int C>
__global__ void __launch_bounds__(kFxLines<Cfg> * Cfg::TL, (C == 1 && kFxLines<Cfg> < kKL) ? FSM_FX_MINB : FSM_MINB(kFxLines<Cfg> * Cfg::TL))
k_pass_fx(Geom<T> g, const cplx<T>* __restrict__ win, long win_fstride,
                                                  Combine<T> cb, FxEpilogue<T> ep, int nlines, int b0, Blk ib, long line_stride) {
    constexpr int N = Cfg::N, EPT = Cfg::EPT, TL = Cfg::TL;
    FSM_DYN_SMEM(smem_raw);
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* bufs = tw + Smem<Cfg, T>::TWPAD;
    const int K = blockDim.x / TL;
    const int lt = threadIdx.x / TL, tau = threadIdx.x % TL;
    const int line = blockIdx.x * K + lt;
    const long bl = blockIdx.z;          // sample index inside the chunk (W buffers)
    const long b = b0 + bl;              // global sample index (state arrays)
    if ((g.pf & (PF_FX_OPS | PF_FX_WIN)) && tau == 0) {
        if (C == 1 && (g.pf & PF_FX_OPS) && line < nlines) {   // measured: hurts the 3-channel lines of C4/C5
            // this line's combine operands: requested now, used after the transform
            FSM_UNROLL
            for (int c = 0; c < C; ++c) {
                FSM_UNROLL
                for (int i = 0; i < FSM_MAX_IN; ++i)
                    if (i < cb.n_in) l2_prefetch(cb.in[i] + (b * C + c) * g.nmodes + (long)line * N, (long)N * sizeof(cplx<T>));
                FSM_UNROLL
                for (int q = 0; q < FSM_MAX_TAB; ++q)
                    if (q < cb.n_tab) l2_prefetch(cb.tab[q] + b * cb.tab_bstride + c * cb.tab_cstride + (long)line * N, (long)N * sizeof(T));
            }
        }
        int x2, y2, z2;
        if ((g.pf & PF_FX_WIN) && ib.shift >= 30 && next_wave_block(g.pf_wave, x2, y2, z2)) {
            const int line2 = x2 * K + lt;
            if (line2 < nlines) {
                FSM_UNROLL
                for (int c = 0; c < C; ++c)
                    l2_prefetch(win + ((long)z2 * C + c) * win_fstride + (long)line2 * line_stride, (long)N * sizeof(cplx<T>));
            }
        }
    }
    twiddles_begin<Cfg, T>(tw);
    if (g.pf & PF_TW_EARLY) twiddles_ready();
    LineSync<TL> sync{1 + lt};
    cplx<T>* mybuf = bufs + lt * Cfg::LINE_PITCH;
    constexpr int NB = (EPT >= 4) ? 4 : EPT;
    const long line_mode0 = (long)line * N;
    // single-channel lines: the operands of the first combine block are requested before the transform runs and
    // every later block is requested while the previous one is being combined (software pipeline)
    // (measured: at N = 1024 the extra operand registers spill and cost more than the hidden latency, so the
    // pipeline is used for lines up to 512 points only)
    constexpr bool kPipe = (C == 1) && (N <= 512);
    CombineOperands<T, NB> opq[2];   // dead (optimised away) when !kPipe
    if constexpr (kPipe) {
        if (line < nlines) combine_load<T, NB>(cb, b * g.nmodes, b * cb.tab_bstride, line_mode0 + tau, TL, opq[0]);
    }
    cplx<T> nhat[C][EPT];
    auto load_channel = [&](int c, cplx<T>* dst) FSM_INLINE_LAMBDA {
        const cplx<T>* src = win + (bl * C + c) * win_fstride + (long)line * line_stride;
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) dst[m] = src[(ib.shift >= 30) ? (long)(tau + m * TL) : blk_off(tau + m * TL, ib, 1)];
    };
    if (line < nlines) load_channel(0, nhat[0]);
    twiddles_ready();      // first channel in flight while the twiddle copy lands
    if (line >= nlines) return;
    static_for<0, C>([&](auto cc) FSM_INLINE_LAMBDA {
        constexpr int c = decltype(cc)::value;
        if constexpr (c > 0) load_channel(c, nhat[c]);
        if (c > 0) sync();
        line_fft<Cfg, -1, T>(nhat[c], mybuf, tw, tau, sync);
    });
    // line coordinates
    int ky, kz = 0;
    if (g.ndim == 3) { ky = g.ky0 + g.kys * (line / g.nh); kz = line % g.nh; } else { ky = line; }
    // the epilogue is instantiated once per compile-time combine shape (and once generic); cb.kind is uniform
    with_combine_kind(cb.kind, [&](auto kindc) FSM_INLINE_LAMBDA {
    constexpr int KIND = decltype(kindc)::value;
    static_for<0, EPT / NB>([&](auto mbc) FSM_INLINE_LAMBDA {
        constexpr int mb = decltype(mbc)::value * NB;
        constexpr int cur = decltype(mbc)::value & 1;
        if constexpr (kPipe && mb + NB < EPT)
            combine_load<T, NB, KIND>(cb, b * g.nmodes, b * cb.tab_bstride, line_mode0 + tau + (mb + NB) * TL, TL, opq[cur ^ 1]);
        cplx<T> f[C][NB];
        const T nlc = ep.nl_coef_b ? ep.nl_coef_b[b] : ep.nl_coef;
        FSM_UNROLL
        for (int j = 0; j < NB; ++j) {
            const int p = tau + (mb + j) * TL;
            FSM_UNROLL
            for (int c = 0; c < C; ++c) f[c][j] = cscale(nhat[c][mb + j], nlc);
            if constexpr (C == 3 || C == 2) {
                if (ep.project) {
                    // result_i = (ik_i) lap^-1 sum_j (ik_j) c_j - c_i   (_navier_stokes.py:249-254), with the
                    // Hermitian projection of the composite symbol: a cross term is dropped when exactly one
                    // of its two axes sits on its Nyquist index. c = coef * (conv - force).
                    if (ep.force) {
                        FSM_UNROLL
                        for (int c = 0; c < C; ++c) f[c][j] = f[c][j] + ep.force[(long)c * g.nmodes + line_mode0 + p];
                    }
                    if (ep.force_dyn) {
                        FSM_UNROLL
                        for (int c = 0; c < C; ++c)
                            f[c][j] = f[c][j] - cscale(ep.force_dyn[(b * C + c) * g.nmodes + line_mode0 + p], ep.nl_coef);
                    }
                    const int kk[3] = {p, ky, kz};
                    T dd[C];
                    bool qq[C];
                    T k2 = T(0);
                    FSM_UNROLL
                    for (int i = 0; i < C; ++i) {
                        dd[i] = g.dkraw[i][kk[i]];
                        qq[i] = (kk[i] == g.n[i] / 2);
                        k2 = fsm_fma(dd[i], dd[i], k2);
                    }
                    const T ik2 = (k2 == T(0)) ? T(0) : T(1) / k2;
                    cplx<T> r[C];
                    FSM_UNROLL
                    for (int i = 0; i < C; ++i) {
                        cplx<T> s = mk<T>(T(0), T(0));
                        FSM_UNROLL
                        for (int jj = 0; jj < C; ++jj) {
                            const T w = (i == jj || qq[i] == qq[jj]) ? dd[i] * dd[jj] * ik2 : T(0);
                            s.x = fsm_fma(w, f[jj][j].x, s.x);
                            s.y = fsm_fma(w, f[jj][j].y, s.y);
                        }
                        r[i] = s - f[i][j];
                    }
                    FSM_UNROLL
                    for (int i = 0; i < C; ++i) f[i][j] = r[i];
                }
            }
        }
        if (ep.source) {
            FSM_UNROLL
            for (int c = 0; c < C; ++c) {
                FSM_UNROLL
                for (int j = 0; j < NB; ++j)
                    f[c][j] = f[c][j] + ep.source[(long)c * g.nmodes + line_mode0 + tau + (mb + j) * TL];
            }
        }
        if constexpr (C == 3 || C == 2) {
            if (ep.project && ep.force_dyn) {
                FSM_UNROLL
                for (int c = 0; c < C; ++c) {
                    FSM_UNROLL
                    for (int j = 0; j < NB; ++j)
                        f[c][j] = f[c][j] + cscale(ep.force_dyn[(b * C + c) * g.nmodes + line_mode0 + tau + (mb + j) * TL], ep.nl_coef);
                }
            }
        }
        if (ep.dc_out && line == 0 && g.ky0 == 0 && tau == 0 && mb == 0) {
            ep.dc_out[b] = f[0][0].x;
            f[0][0] = mk<T>(T(0), T(0));
        }
        if constexpr (kPipe) {
            combine_apply<T, NB, KIND>(cb, f[0], opq[cur], b * g.nmodes, line_mode0 + tau + mb * TL, TL);
        } else {
            FSM_UNROLL
            for (int c = 0; c < C; ++c)
                combine_block<T, NB, KIND>(cb, f[c], (b * C + c) * g.nmodes, b * cb.tab_bstride + c * cb.tab_cstride, line_mode0 + tau + mb * TL, TL);
        }
    });
    });
}

// ------------------------------------------------------------------------------------------
// 1-D grids: the whole time loop of one sample runs inside ONE kernel launch (one CTA per sample,
// state and stage arrays stay in L1/L2; the path is latency-bound).
//   per stage: Z = u_hat + i (i k u_hat) (dealiased, mirrored) -> inverse FFT -> (u, u_x) -> u u_x
//              -> forward FFT -> N_hat(k <= N/2) -> combine
// Also provides the plain 1-D transforms (MODE_R2C / MODE_C2R).
// ------------------------------------------------------------------------------------------
#define FSM_MAX_STAGES 4
template <typename T>
struct StageList {
    int n_stages;
    const cplx<T>* input[FSM_MAX_STAGES];
    Combine<T> cb[FSM_MAX_STAGES];
};

// The state and stage arrays of the sample (at most FSM_1D_ARRAYS distinct ones, N/2+1 modes each) are copied into shared
// memory once, every stage of every step works on those copies (the Combine descriptors are re-pointed at them), and
// all of them are written back at the end: between two stages nothing but the coefficient tables is read from global
// memory. (First version: every stage re-read its input and operands from L2 and wrote its outputs there, ~1 us per
// dependent round trip; C1 ran at 44 us per step.)
#define FSM_1D_ARRAYS 6
template <typename T, class Cfg>
__global__ void __launch_bounds__(Cfg::TL) k_step1d(Geom<T> g, StageList<T> sl, FxEpilogue<T> ep, int n_steps) {
    constexpr int N = Cfg::N, EPT = Cfg::EPT, TL = Cfg::TL, NH = N / 2 + 1, APITCH = NH + 1;
    FSM_DYN_SMEM(smem_raw);
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* buf = tw + Smem<Cfg, T>::TWPAD;
    cplx<T>* arrs = buf + Cfg::LINE_PITCH;
    make_twiddles<Cfg, T>(tw);
    const int tau = threadIdx.x;
    const long b = blockIdx.x;
    const long boff = b * g.nmodes;
    const int kmax = g.kmax[0];
    const T* dk = g.dk[0];
    LineSync<TL> sync{1};
    // distinct global arrays touched by the stage list (every thread builds the same list)
    const cplx<T>* base[FSM_1D_ARRAYS];
    int nbase = 0;
    auto slot_of = [&](const cplx<T>* p) FSM_INLINE_LAMBDA {
        for (int i = 0; i < nbase; ++i)
            if (base[i] == p) return i;
        if (nbase < FSM_1D_ARRAYS) base[nbase] = p;
        return nbase++;
    };
    for (int si = 0; si < sl.n_stages; ++si) {
        slot_of(sl.input[si]);
        for (int i = 0; i < sl.cb[si].n_in; ++i) slot_of(sl.cb[si].in[i]);
        for (int r = 0; r < sl.cb[si].n_out; ++r) slot_of(sl.cb[si].out[r]);
    }
    const bool resident = nbase <= FSM_1D_ARRAYS;     // always true for the stage programs of fsm_plan.cu
    if (resident) {
        for (int a = 0; a < nbase; ++a)
            for (int k = tau; k < NH; k += TL) arrs[a * APITCH + k] = base[a][boff + k];
    }
    __syncthreads();
    for (int step = 0; step < n_steps; ++step) {
        for (int si = 0; si < sl.n_stages; ++si) {
            Combine<T> cb = sl.cb[si];
            const cplx<T>* in = sl.input[si] + boff;
            if (resident) {        // re-point the descriptor: element [boff + k] of an array is arrs[slot * APITCH + k]
                in = arrs + slot_of(sl.input[si]) * APITCH;
                for (int i = 0; i < cb.n_in; ++i) cb.in[i] = arrs + slot_of(cb.in[i]) * APITCH - boff;
                for (int r = 0; r < cb.n_out; ++r) cb.out[r] = arrs + slot_of(sl.cb[si].out[r]) * APITCH - boff;
            }
            cplx<T> v[EPT];
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) {
                const int p = tau + m * TL;
                const int k = (p <= N / 2) ? p : N - p;
                cplx<T> A = mk<T>(T(0), T(0));
                if (k <= kmax) A = cscale(in[k], g.inv_ntot);
                cplx<T> Bq = cmul_i(A, dk[k]);
                if (k == 0 || k == N / 2) { A.y = T(0); Bq.y = T(0); }
                if (p > N / 2) { A.y = -A.y; Bq.y = -Bq.y; }
                v[m] = mk<T>(A.x - Bq.y, A.y + Bq.x);
            }
            line_fft<Cfg, +1, T>(v, buf, tw, tau, sync);
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) v[m] = mk<T>(v[m].x * v[m].y, T(0));
            sync();
            line_fft<Cfg, -1, T>(v, buf, tw, tau, sync);
            FSM_UNROLL
            for (int m = 0; m < EPT; ++m) {
                const int p = tau + m * TL;
                if (p <= N / 2) {
                    cplx<T> f = cscale(v[m], ep.nl_coef_b ? ep.nl_coef_b[b] : ep.nl_coef);
                    if (ep.source) f = f + ep.source[p];
                    combine_mode<T>(cb, f, boff, b * cb.tab_bstride, p);
                }
            }
            __syncthreads();
        }
    }
    if (resident) {
        for (int a = 0; a < nbase; ++a) {
            cplx<T>* dst = const_cast<cplx<T>*>(base[a]);
            for (int k = tau; k < NH; k += TL) dst[boff + k] = arrs[a * APITCH + k];
        }
    }
}

enum { MODE1D_R2C = 0, MODE1D_C2R = 1 };
template <typename T, class Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::TL) k_line1d(const void* __restrict__ in_v, void* __restrict__ out_v) {
    constexpr int N = Cfg::N, EPT = Cfg::EPT, TL = Cfg::TL, NH = N / 2 + 1;
    FSM_DYN_SMEM(smem_raw);
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* buf = tw + Smem<Cfg, T>::TWPAD;
    make_twiddles<Cfg, T>(tw);
    const int tau = threadIdx.x;
    const long f = blockIdx.x;
    LineSync<TL> sync{1};
    cplx<T> v[EPT];
    if constexpr (MODE == MODE1D_R2C) {
        const T* in = static_cast<const T*>(in_v) + f * N;
        cplx<T>* out = static_cast<cplx<T>*>(out_v) + f * NH;
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) v[m] = mk<T>(in[tau + m * TL], T(0));
        line_fft<Cfg, -1, T>(v, buf, tw, tau, sync);
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) {
            const int p = tau + m * TL;
            if (p <= N / 2) out[p] = v[m];
        }
    } else {
        const cplx<T>* in = static_cast<const cplx<T>*>(in_v) + f * NH;
        T* out = static_cast<T*>(out_v) + f * N;
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) {
            const int p = tau + m * TL;
            const int k = (p <= N / 2) ? p : N - p;
            cplx<T> A = cscale(in[k], T(1) / T(N));
            if (k == 0 || k == N / 2) A.y = T(0);
            if (p > N / 2) A.y = -A.y;
            v[m] = A;
        }
        line_fft<Cfg, +1, T>(v, buf, tw, tau, sync);
        FSM_UNROLL
        for (int m = 0; m < EPT; ++m) out[tau + m * TL] = v[m].x;
    }
}

// Combine without a nonlinear term (ETDRK0) or point-wise fix-ups: one thread per mode.
// The fresh term X_0 is `fresh` ([B][C][nmodes], a nonlinear term evaluated outside the fused passes:
// fsm_stage_combine) plus `source` ([C][nmodes], constant spectrum), either may be null.
template <typename T>
__global__ void k_combine_only(Combine<T> cb, long nmodes, int C, long total /*B*C*nmodes*/,
                               const cplx<T>* __restrict__ fresh, const cplx<T>* __restrict__ source) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long mode = i % nmodes;
    const long bc = i / nmodes;
    const int c = (int)(bc % C);
    cplx<T> f = mk<T>(T(0), T(0));
    if (fresh) f = fresh[i];
    if (source) f = f + source[(long)c * nmodes + mode];
    combine_mode<T>(cb, f, bc * nmodes, (bc / C) * cb.tab_bstride + c * cb.tab_cstride, mode);
}

// KS: N_hat_b(0) = coef * (S_b - mean_b S_b). The FX pass stored coef*S_b in dc[b] and combined a zero.
// This kernel adds the missing contribution (coefficient of the fresh term at mode 0) to every output.
template <typename T>
__global__ void k_ks_dc_fix(Combine<T> cb, const T* dc, int B, long nmodes, T* log_slot) {
    FSM_DYN_SMEM(smem_raw);
    T* s_mean = reinterpret_cast<T*>(smem_raw);
    if (threadIdx.x == 0) {
        T s = T(0);
        for (int b = 0; b < B; ++b) s += dc[b];
        s_mean[0] = s / T(B);
        if (log_slot) *log_slot = s;   // local sum of this evaluation (multi-rank zero-mode correction)
    }
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const T delta = dc[b] - s_mean[0];
        for (int r = 0; r < cb.n_out; ++r) {
            const int ti = cb.ct[r][0];
            if (ti == -2) continue;
            T coef = cb.ca[r][0];
            if (ti >= 0) coef += cb.cb[r][0] * cb.tab[ti][(long)b * cb.tab_bstride];
            cplx<T>* o = cb.out[r] + (long)b * nmodes;  // C == 1
            o[0].x += coef * delta;
        }
    }
}

// half (rot) <-> full complex spectrum (reference layout (B*C, n0, n1[, n2])), used for u_0_fft input,
// return_in_fourier and recorder frames. Hermitian completion of the missing half.
template <typename T>
__global__ void k_half_to_full(Geom<T> g, const cplx<T>* __restrict__ half, cplx<T>* __restrict__ full, long nfields) {
    const long ntot = (long)g.n[0] * g.n[1] * g.n[2];
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nfields * ntot) return;
    const long f = i / ntot;
    long rem = i % ntot;
    int idx[3];
    idx[2] = (int)(rem % g.n[2]); rem /= g.n[2];
    idx[1] = (int)(rem % g.n[1]); rem /= g.n[1];
    idx[0] = (int)rem;
    // physical index order (x, y, z) -> for 2-D the arrays are (n0, n1, 1)
    const int last = g.ndim - 1;
    const int nl = g.n[last];
    bool conj = idx[last] > nl / 2;
    int q[3] = {idx[0], idx[1], idx[2]};
    if (conj)
        for (int d = 0; d < g.ndim; ++d) q[d] = (g.n[d] - idx[d]) % g.n[d];
    long mode;
    if (g.ndim == 1) mode = q[0];
    else if (g.ndim == 2) mode = (long)q[1] * g.n[0] + q[0];
    else mode = ((long)q[1] * g.nh + q[2]) * g.n[0] + q[0];
    cplx<T> v = half[f * g.nmodes + mode];
    if (conj) v.y = -v.y;
    full[i] = v;
}

template <typename T>
__global__ void k_full_to_half(Geom<T> g, const cplx<T>* __restrict__ full, cplx<T>* __restrict__ half, long nfields) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nfields * g.nmodes) return;
    const long f = i / g.nmodes;
    long mode = i % g.nmodes;
    int q[3] = {0, 0, 0};
    if (g.ndim == 1) q[0] = (int)mode;
    else if (g.ndim == 2) { q[0] = (int)(mode % g.n[0]); q[1] = (int)(mode / g.n[0]); }
    else { q[0] = (int)(mode % g.n[0]); mode /= g.n[0]; q[2] = (int)(mode % g.nh); q[1] = (int)(mode / g.nh); }
    // Hermitian projection of the real-field spectrum: (U(k) + conj U(-k)) / 2
    const long ntot = (long)g.n[0] * g.n[1] * g.n[2];
    const long a = ((long)q[0] * g.n[1] + q[1]) * g.n[2] + q[2];
    const int m0 = (g.n[0] - q[0]) % g.n[0], m1 = (g.n[1] - q[1]) % g.n[1], m2 = (g.n[2] - q[2]) % g.n[2];
    const long bidx = ((long)m0 * g.n[1] + m1) * g.n[2] + m2;
    const cplx<T> u = full[f * ntot + a], w = full[f * ntot + bidx];
    half[i] = mk<T>(T(0.5) * (u.x + w.x), T(0.5) * (u.y - w.y));
}

}  // namespace fsm

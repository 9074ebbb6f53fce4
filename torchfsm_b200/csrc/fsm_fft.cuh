// In-register radix butterflies and the per-line multi-stage Stockham FFT.
//
// Data distribution of one line of length N handled by TL = N/EPT threads:
//   thread tau owns positions  p = tau + m*TL,  m = 0..EPT-1   (register v[m])
// both on entry (natural-order input) and on exit (natural-order output). With that
// ownership a warp's global loads/stores of consecutive tau are fully coalesced and
// point-wise products between several transformed lines need no data movement.
// Stages exchange data through one padded shared-memory line buffer.
#pragma once
#include "fsm_compat.h"
#include <type_traits>

namespace fsm {

template <typename T>
struct alignas(2 * sizeof(T)) cplx {
    T x, y;
};

template <typename T> FSM_HD __forceinline__ cplx<T> mk(T x, T y) { cplx<T> r; r.x = x; r.y = y; return r; }
template <typename T> FSM_HD __forceinline__ cplx<T> operator+(cplx<T> a, cplx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> FSM_HD __forceinline__ cplx<T> operator-(cplx<T> a, cplx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T> FSM_HD __forceinline__ cplx<T> cmul(cplx<T> a, cplx<T> b) {
    return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
template <typename T> FSM_HD __forceinline__ cplx<T> cmulc(cplx<T> a, cplx<T> b) {
    return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
template <typename T> FSM_HD __forceinline__ cplx<T> cscale(cplx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
template <typename T> FSM_HD __forceinline__ cplx<T> cconj(cplx<T> a) { return mk<T>(a.x, -a.y); }
// multiply by i*s (s real)
template <typename T> FSM_HD __forceinline__ cplx<T> cmul_i(cplx<T> a, T s) { return mk<T>(-a.y * s, a.x * s); }

// ------------------------------------------------------------------------------------
// Blackwell packed fp32 pairs: add/mul/fma.rn.f32x2 (SASS FADD2 / FMUL2 / FFMA2) work on a 64-bit
// register pair, i.e. on one complex number; ptxas folds the half-swap, per-half sign flips and scalar
// broadcasts into operand modifiers. The passes are issue-bound, so a complex add is 1 instruction
// instead of 2, a complex multiply 2 instead of 4, a twiddled radix-2 butterfly 4 instead of 8.
// ------------------------------------------------------------------------------------
#if defined(__CUDACC__) && !defined(FSM_EMU) && !defined(FSM_NO_F32X2)
#define FSM_PACKED 1
typedef unsigned long long fsm_u64;
__device__ __forceinline__ fsm_u64 pk(float lo, float hi) { fsm_u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ fsm_u64 pk(cplx<float> a) { return pk(a.x, a.y); }
__device__ __forceinline__ cplx<float> upk(fsm_u64 v) { cplx<float> r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ fsm_u64 add2(fsm_u64 a, fsm_u64 b) { fsm_u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ fsm_u64 sub2(fsm_u64 a, fsm_u64 b) { fsm_u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ fsm_u64 mul2(fsm_u64 a, fsm_u64 b) { fsm_u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ fsm_u64 fma2(fsm_u64 a, fsm_u64 b, fsm_u64 c) { fsm_u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// non-template overloads win over the generic templates for cplx<float>
__device__ __forceinline__ cplx<float> operator+(cplx<float> a, cplx<float> b) { return upk(add2(pk(a), pk(b))); }
__device__ __forceinline__ cplx<float> operator-(cplx<float> a, cplx<float> b) { return upk(sub2(pk(a), pk(b))); }
__device__ __forceinline__ cplx<float> cmul(cplx<float> a, cplx<float> b) {
    return upk(fma2(pk(a.y, a.x), pk(-b.y, b.y), mul2(pk(a), pk(b.x, b.x))));
}
__device__ __forceinline__ cplx<float> cmulc(cplx<float> a, cplx<float> b) {
    return upk(fma2(pk(a.y, a.x), pk(b.y, -b.y), mul2(pk(a), pk(b.x, b.x))));
}
__device__ __forceinline__ cplx<float> cscale(cplx<float> a, float s) { return upk(mul2(pk(a), pk(s, s))); }
__device__ __forceinline__ cplx<float> cmul_i(cplx<float> a, float s) { return upk(mul2(pk(a.y, a.x), pk(-s, s))); }
#endif

// Radix-2 combine  p = e + W o,  m = e - W o  for W = (c, s) given at compile time.
template <typename T>
struct Bfly {
    static FSM_HD __forceinline__ void plain(cplx<T> e, cplx<T> o, cplx<T>& p, cplx<T>& m) { p = e + o; m = e - o; }
    template <int D>  // W = (0, D)
    static FSM_HD __forceinline__ void rot(cplx<T> e, cplx<T> o, cplx<T>& p, cplx<T>& m) {
        const cplx<T> t = mk<T>(-T(D) * o.y, T(D) * o.x);
        p = e + t; m = e - t;
    }
    static FSM_HD __forceinline__ void gen(cplx<T> e, cplx<T> o, T c, T s, cplx<T>& p, cplx<T>& m) {
        const cplx<T> t = mk<T>(c * o.x - s * o.y, c * o.y + s * o.x);
        p = e + t; m = e - t;
    }
};
#ifdef FSM_PACKED
template <>
struct Bfly<float> {
    static __device__ __forceinline__ void plain(cplx<float> e, cplx<float> o, cplx<float>& p, cplx<float>& m) {
        p = upk(add2(pk(e), pk(o))); m = upk(sub2(pk(e), pk(o)));
    }
    template <int D>
    static __device__ __forceinline__ void rot(cplx<float> e, cplx<float> o, cplx<float>& p, cplx<float>& m) {
        const fsm_u64 so = pk(o.y, o.x), ee = pk(e);
        p = upk(fma2(so, pk(-float(D), float(D)), ee));
        m = upk(fma2(so, pk(float(D), -float(D)), ee));
    }
    static __device__ __forceinline__ void gen(cplx<float> e, cplx<float> o, float c, float s, cplx<float>& p, cplx<float>& m) {
        const fsm_u64 so = pk(o.y, o.x), oo = pk(o), ee = pk(e);
        p = upk(fma2(so, pk(-s, s), fma2(oo, pk(c, c), ee)));
        m = upk(fma2(so, pk(s, -s), fma2(oo, pk(-c, -c), ee)));
    }
};
#endif

// acc + coef * x (real coefficient)
template <typename T> FSM_HD __forceinline__ cplx<T> cfma_s(T coef, cplx<T> x, cplx<T> acc) {
    return mk<T>(coef * x.x + acc.x, coef * x.y + acc.y);
}
#ifdef FSM_PACKED
__device__ __forceinline__ cplx<float> cfma_s(float coef, cplx<float> x, cplx<float> acc) {
    return upk(fma2(pk(coef, coef), pk(x), pk(acc)));
}
#endif

// compile-time loop with a constexpr index
template <int I, int E, class F>
FSM_HD __forceinline__ void static_for(F&& f) {
    if constexpr (I < E) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, E>(f);
    }
}

// cos / sin of 2*pi*j/32, j = 0..16
FSM_HD constexpr double cos32(int j) {
    constexpr double t[17] = {1.0,
                              0.98078528040323044913,
                              0.92387953251128675613,
                              0.83146961230254523708,
                              0.70710678118654752440,
                              0.55557023301960222474,
                              0.38268343236508977173,
                              0.19509032201612826785,
                              0.0,
                              -0.19509032201612826785,
                              -0.38268343236508977173,
                              -0.55557023301960222474,
                              -0.70710678118654752440,
                              -0.83146961230254523708,
                              -0.92387953251128675613,
                              -0.98078528040323044913,
                              -1.0};
    return t[j];
}
FSM_HD constexpr double sin32(int j) { return j <= 8 ? cos32(8 - j) : cos32(j - 8); }

// DIR = -1: forward (exp(-2 pi i jk/R)); DIR = +1: inverse (unnormalised).
template <int R, int DIR, typename T>
struct Dft {
    static_assert(R == 4 || R == 8 || R == 16 || R == 32, "radix");
    static FSM_HD __forceinline__ void run(cplx<T>* a) {
        cplx<T> e[R / 2], o[R / 2];
        static_for<0, R / 2>([&](auto kc) FSM_INLINE_LAMBDA {
            constexpr int k = decltype(kc)::value;
            e[k] = a[2 * k];
            o[k] = a[2 * k + 1];
        });
        Dft<R / 2, DIR, T>::run(e);
        Dft<R / 2, DIR, T>::run(o);
        static_for<0, R / 2>([&](auto kc) FSM_INLINE_LAMBDA {
            constexpr int k = decltype(kc)::value;
            constexpr int j = k * (32 / R);  // W_R^k = W_32^j, 0 <= j < 16
            if constexpr (j == 0) {
                Bfly<T>::plain(e[k], o[k], a[k], a[k + R / 2]);
            } else if constexpr (j == 8) {  // W = (0, DIR)
                Bfly<T>::template rot<DIR>(e[k], o[k], a[k], a[k + R / 2]);
            } else {
                constexpr T c = T(cos32(j));
                constexpr T s = T(DIR) * T(sin32(j));
                Bfly<T>::gen(e[k], o[k], c, s, a[k], a[k + R / 2]);
            }
        });
    }
};
template <int DIR, typename T>
struct Dft<2, DIR, T> {
    static FSM_HD __forceinline__ void run(cplx<T>* a) {
        cplx<T> p, m;
        Bfly<T>::plain(a[0], a[1], p, m);
        a[0] = p;
        a[1] = m;
    }
};
template <int DIR, typename T>
struct Dft<1, DIR, T> {
    static FSM_HD __forceinline__ void run(cplx<T>*) {}
};

// ------------------------------------------------------------------------------------
// FFT configuration: N = R0*R1*R2, every radix divides EPT, TL = N/EPT threads per line.
// ------------------------------------------------------------------------------------
// Stage twiddles W^(j*t), t = 1..R-1: with FSM_TW_RECUR only the t = 1 row is stored and read (one LDS per
// butterfly group); the other rows are its powers, formed in registers by repeated multiplication
// p[t] = p[t/2] * p[t - t/2] (depth <= log2 R, a few 1e-8 of extra rounding in fp32). The passes are bound by the
// shared-memory pipe (ncu: 41-70 % of the wavefront peak with 28 twiddle LDS per 1024-point transform), and the
// packed-FMA pipe has room.
#ifndef FSM_TW_RECUR
#define FSM_TW_RECUR 1
#endif
template <int N_, int EPT_, int R0_, int R1_ = 1, int R2_ = 1>
struct FftCfg {
    static constexpr int N = N_, EPT = EPT_, TL = N_ / EPT_;
    static constexpr int R0 = R0_, R1 = R1_, R2 = R2_;
    static constexpr int NST = (R2_ > 1) ? 3 : ((R1_ > 1) ? 2 : 1);
    static_assert(R0_ * R1_ * R2_ == N_, "radix product");
    static_assert(EPT_ % R0_ == 0 && EPT_ % R1_ == 0 && EPT_ % R2_ == 0, "radix must divide EPT");
    static constexpr int PADSHIFT = (R0_ >= 32) ? 5 : 4;
    // padded index inside a line buffer
    static FSM_HD constexpr int pad(int i) { return i + (i >> PADSHIFT); }
    // pad(a + c) == pad(a) + pad(c) when the low PADSHIFT bits cannot carry. Every buffer index of the
    // stages is (thread part) + (compile-time part); where the split is exact the thread part is padded
    // once and the compile-time part becomes the immediate offset of the LDS/STS.
    static constexpr int PADG = 1 << PADSHIFT;
    static constexpr int TLG = (N_ / EPT_ < PADG) ? N_ / EPT_ : PADG;   // granularity of tau's low bits
    // complex elements per line buffer, rounded so that LINE_PITCH % 16 == 2: eight lines
    // read "column-wise" by consecutive lanes (transposed stores) hit distinct banks.
    // Layout of the LAST stage's input when the middle stage writes it (NST == 3). With R0 = R1 = 8 the middle stage's
    // half-warp stores two 8-slot runs (lanes 0-7 and 8-15: butterfly groups g and g+1) whose bases are pad(64) = 68
    // slots apart: 4 of 16 bank pairs collide (ncu: every excess shared-memory wavefront of the 3-D last-axis pass
    // sat on these STS.64, profiles/r2f_phys3d_shared_wavefronts.txt). Four extra slots in front of every odd group
    // of 64 (and, to stay injective, of everything behind it: 4 * ceil(g / 2) slots before group g) put the two runs
    // 72 = 8 (mod 16) apart; the loads of the last stage (consecutive lanes, consecutive slots) do not care. The
    // offset depends on the group index i >> 6 only, which comes from the thread part in the stores and from the
    // compile-time part in the loads, so the affine splits below stay exact.
    static constexpr int XTRA = (R2_ > 1 && R0_ == 8 && R1_ == 8 && PADSHIFT == 4) ? 4 : 0;
    static_assert(XTRA == 0 || EPT_ == R1_, "pad_last: one butterfly per thread in the middle stage");
    static FSM_HD constexpr int pad_last(int i) { return pad(i) + XTRA * (((i >> 6) + 1) >> 1); }
    static constexpr int RAWLEN = N_ + (N_ >> PADSHIFT) + 1 + XTRA * (((N_ >> 6) + 1) >> 1);
    static constexpr int LINE_PITCH = RAWLEN + ((2 - RAWLEN % 16) + 16) % 16;
    // twiddle tables (complex entries): stage 1 uses (R1-1)*R0, stage 2 uses (R2-1)*R0*R1 -- or only their
    // first rows (R0, R0*R1 entries) when the other rows are formed by recurrence
    static constexpr bool RECUR = (FSM_TW_RECUR != 0);
    static constexpr int TWROWS1 = RECUR ? 1 : (R1_ - 1), TWROWS2 = RECUR ? 1 : (R2_ - 1);
    static constexpr int TW1 = (R1_ > 1) ? TWROWS1 * R0_ : 0;
    static constexpr int TW2 = (R2_ > 1) ? TWROWS2 * R0_ * R1_ : 0;
    static constexpr int TW_TOTAL = TW1 + TW2;
};

// Synchronisation among the TL threads of one line.
template <int TL>
struct LineSync {
    int bar_id;  // named barrier id when the line spans several warps
    __device__ __forceinline__ void operator()() const {
        if constexpr (TL <= 32) {
            __syncwarp();
        } else {
            FSM_NAMED_BARRIER(bar_id, TL);
        }
    }
};

// p[t] = b^t for t = 1..R-1 (p[0] unused)
template <int R, typename T>
FSM_HD __forceinline__ void twiddle_powers(cplx<T> b, cplx<T>* p) {
    p[1] = b;
    static_for<2, R>([&](auto tc) FSM_INLINE_LAMBDA {
        constexpr int t = decltype(tc)::value;
        p[t] = cmul(p[t >> 1], p[t - (t >> 1)]);
    });
}

// ---- building blocks -----------------------------------------------------------------------
// line_fft_head: all stages but the last. On exit the input of the last stage sits in `buf`
// (padded layout); the caller synchronises before the last stage reads it.
template <class Cfg, int DIR, typename T>
__device__ __forceinline__ void line_fft_head(cplx<T>* v, cplx<T>* buf, const cplx<T>* tw, int tau,
                                              const LineSync<Cfg::TL>& sync) {
    constexpr int N = Cfg::N, EPT = Cfg::EPT, TL = Cfg::TL;
    constexpr int R0 = Cfg::R0, R1 = Cfg::R1;
    static_assert(Cfg::NST >= 2, "head/last split needs at least two stages");
    // ---- stage 0: Ns = 1, no twiddles
    static_for<0, EPT / R0>([&](auto qc) FSM_INLINE_LAMBDA {
        constexpr int q = decltype(qc)::value;
        cplx<T> a[R0];
        static_for<0, R0>([&](auto tc) FSM_INLINE_LAMBDA {
            constexpr int t = decltype(tc)::value;
            a[t] = v[q + t * (EPT / R0)];
        });
        Dft<R0, DIR, T>::run(a);
        const int w = tau + q * TL;
        static_for<0, R0>([&](auto tc) FSM_INLINE_LAMBDA {
            constexpr int t = decltype(tc)::value;
            buf[Cfg::pad(w * R0 + t)] = a[t];
        });
    });
    if constexpr (Cfg::NST == 3) {
        sync();
        // ---- stage 1: Ns = R0 (middle stage)
        constexpr int Ns = R0;
        constexpr bool kAffLd = ((N / R1) % Cfg::TLG == 0);
        const cplx<T>* ldb = buf + Cfg::pad(tau);
        static_for<0, EPT / R1>([&](auto qc) FSM_INLINE_LAMBDA {
            constexpr int q = decltype(qc)::value;
            const int w = tau + q * TL;
            static_for<0, R1>([&](auto tc) FSM_INLINE_LAMBDA {
                constexpr int t = decltype(tc)::value;
                if constexpr (kAffLd) v[q * R1 + t] = ldb[Cfg::pad(q * TL + t * (N / R1))];
                else v[q * R1 + t] = buf[Cfg::pad(w + t * (N / R1))];
            });
        });
        sync();  // everyone has read before anyone overwrites
        // w = tau + q*TL: with TL a multiple of Ns the twiddle row j and the output base split the same way
        constexpr bool kAffSt = (TL % Ns == 0) && (Cfg::PADG % Ns == 0) && ((Ns * R1) % Cfg::PADG == 0) &&
                                ((TL * R1) % Cfg::PADG == 0);
        const int jt = tau & (Ns - 1);
        cplx<T>* stb = buf + Cfg::pad_last((tau / Ns) * Ns * R1 + jt);   // XTRA: the thread part carries bit 6
        cplx<T> pw[R1];   // twiddle row of this thread (same for every q when the split is affine)
        if constexpr (Cfg::RECUR && kAffSt) twiddle_powers<R1, T>(tw[jt], pw);
        static_for<0, EPT / R1>([&](auto qc) FSM_INLINE_LAMBDA {
            constexpr int q = decltype(qc)::value;
            const int w = tau + q * TL;
            const int j = kAffSt ? jt : (w & (Ns - 1));
            if constexpr (Cfg::RECUR && !kAffSt) twiddle_powers<R1, T>(tw[j], pw);
            cplx<T> a[R1];
            a[0] = v[q * R1];
            static_for<1, R1>([&](auto tc) FSM_INLINE_LAMBDA {
                constexpr int t = decltype(tc)::value;
                cplx<T> wv;
                if constexpr (Cfg::RECUR) wv = pw[t];
                else wv = tw[(t - 1) * Ns + j];
                a[t] = (DIR < 0) ? cmul(v[q * R1 + t], wv) : cmulc(v[q * R1 + t], wv);
            });
            Dft<R1, DIR, T>::run(a);
            const int base = (w / Ns) * Ns * R1 + j;
            static_for<0, R1>([&](auto tc) FSM_INLINE_LAMBDA {
                constexpr int t = decltype(tc)::value;
                if constexpr (kAffSt) stb[Cfg::pad(q * TL * R1 + t * Ns)] = a[t];
                else buf[Cfg::pad_last(base + t * Ns)] = a[t];
            });
        });
    }
}

// Last stage for work item w (0 <= w < N/RL): a[t] = X[w + t * N/RL], t < RL.
template <class Cfg>
struct LastStage {
    static constexpr int RL = (Cfg::NST == 3) ? Cfg::R2 : Cfg::R1;
    static constexpr int NS = Cfg::N / RL;   // stride between the outputs of one work item
};
// The work item is w = wt + WC with WC known at compile time (a multiple of TL).
template <class Cfg, int DIR, typename T, int WC = 0>
__device__ __forceinline__ void fft_last_item(const cplx<T>* buf, const cplx<T>* tw, int wt, cplx<T>* a) {
    constexpr int RL = LastStage<Cfg>::RL, NS = LastStage<Cfg>::NS;
    const cplx<T>* twl = ((Cfg::NST == 3) ? tw + Cfg::TW1 : tw) + wt;
    // NST == 2: Ns = R0 = NS and j = w; NST == 3: Ns = R0*R1 = NS and j = w
    constexpr bool kAff = (WC % Cfg::TLG == 0) && (NS % Cfg::TLG == 0);
    const int w = wt + WC;
    // NST == 3: the buffer was written by the middle stage in the pad_last layout; wt + (WC mod 64) < 64 whenever the
    // affine split is taken with XTRA != 0 (TL <= 64, WC a multiple of TL), so bit 6 comes from the compile-time part
    constexpr bool kLast = (Cfg::NST == 3);
    static_assert(Cfg::XTRA == 0 || (Cfg::TL <= 64 && (64 % Cfg::TL) == 0), "pad_last split");
    const cplx<T>* ldb = buf + Cfg::pad(wt);
    if constexpr (kAff) a[0] = ldb[kLast ? Cfg::pad_last(WC) : Cfg::pad(WC)];
    else a[0] = buf[kLast ? Cfg::pad_last(w) : Cfg::pad(w)];
    cplx<T> pw[RL];
    if constexpr (Cfg::RECUR) twiddle_powers<RL, T>(twl[WC], pw);
    static_for<1, RL>([&](auto tc) FSM_INLINE_LAMBDA {
        constexpr int t = decltype(tc)::value;
        cplx<T> x;
        if constexpr (kAff) x = ldb[kLast ? Cfg::pad_last(WC + t * NS) : Cfg::pad(WC + t * NS)];
        else x = buf[kLast ? Cfg::pad_last(w + t * NS) : Cfg::pad(w + t * NS)];
        cplx<T> wv;
        if constexpr (Cfg::RECUR) wv = pw[t];
        else wv = twl[(t - 1) * NS + WC];
        a[t] = (DIR < 0) ? cmul(x, wv) : cmulc(x, wv);
    });
    Dft<RL, DIR, T>::run(a);
}

// v[m] holds x[tau + m*TL] on entry and X[tau + m*TL] on exit. `buf` is this line's
// shared-memory buffer (Cfg::LINE_PITCH complex); `tw` the forward twiddle tables.
// The caller must make sure every thread of the line is done with `buf` before calling.
template <class Cfg, int DIR, typename T>
__device__ __forceinline__ void line_fft(cplx<T>* v, cplx<T>* buf, const cplx<T>* tw, int tau,
                                         const LineSync<Cfg::TL>& sync) {
    constexpr int EPT = Cfg::EPT, TL = Cfg::TL;
    if constexpr (Cfg::NST == 1) {
        Dft<Cfg::R0, DIR, T>::run(v);
    } else {
        constexpr int RL = LastStage<Cfg>::RL;
        line_fft_head<Cfg, DIR, T>(v, buf, tw, tau, sync);
        sync();
        cplx<T> out[EPT];
        static_for<0, EPT / RL>([&](auto qc) FSM_INLINE_LAMBDA {
            constexpr int q = decltype(qc)::value;
            cplx<T> a[RL];
            fft_last_item<Cfg, DIR, T, q * TL>(buf, tw, tau, a);
            static_for<0, RL>([&](auto tc) FSM_INLINE_LAMBDA {
                constexpr int t = decltype(tc)::value;
                out[q + t * (EPT / RL)] = a[t];
            });
        });
        static_for<0, EPT>([&](auto mc) FSM_INLINE_LAMBDA {
            constexpr int m = decltype(mc)::value;
            v[m] = out[m];
        });
    }
}

// Host-side twiddle table for a configuration (forward sign), layout described above.
template <class Cfg, typename T>
inline void fill_twiddles(cplx<T>* tw) {
    const double two_pi = 6.283185307179586476925286766559;
    if (Cfg::R1 > 1) {
        const int Ns = Cfg::R0;
        for (int t = 1; t <= Cfg::TWROWS1; ++t)
            for (int j = 0; j < Ns; ++j) {
                double ang = -two_pi * (double)(j * t) / (double)(Ns * Cfg::R1);
                tw[(t - 1) * Ns + j] = mk<T>((T)__builtin_cos(ang), (T)__builtin_sin(ang));
            }
    }
    if (Cfg::R2 > 1) {
        const int Ns = Cfg::R0 * Cfg::R1;
        cplx<T>* tw2 = tw + Cfg::TW1;
        for (int t = 1; t <= Cfg::TWROWS2; ++t)
            for (int j = 0; j < Ns; ++j) {
                double ang = -two_pi * (double)(j * t) / (double)(Ns * Cfg::R2);
                tw2[(t - 1) * Ns + j] = mk<T>((T)__builtin_cos(ang), (T)__builtin_sin(ang));
            }
    }
}

}  // namespace fsm

// Point-wise spectral maps in the rot-half layout: out_c(k) = sum_t coef_t * prod_a (i k_a)^p_ta * (1/lap)^q_t * in_c'(k).
// One thread per (sample, mode); every operand is read once and every result written once (HBM bound by
// construction: (c_in + c_out) * 8 B per mode). Included by fsm_plan.cu only.
//
// Reference cores this evaluates without a transform in between (they are products with symbol tensors there):
//   _GradCore (generic/_grad.py:6-15), _DivCore (_div.py:9-23), _Curl2DCore/_Curl3DCore (_curl.py:9-55),
//   _Vorticity2VelocityCore (dedicated/_navier_stokes.py:75-92), the pressure solve of
//   _Velocity2PressureCore/_Vorticity2PressureCore (:158-163, :213-217), and the linear cores when they appear
//   in the same sum (_laplacian.py:7-15, _biharmonic.py:8-16, _spatial_derivative.py:7-20, _source.py:9-17).
//
// Hermitian projection: the reference multiplies its FULL spectrum and keeps .real of the inverse
// transform. For a real input field that equals the half-spectrum product with S_sym(k) = (S(k) + conj S(-k)) / 2;
// for a monomial symbol S = prod (i k_a)^p_a this is S itself unless the powers on the axes that sit on their
// Nyquist index add up to an odd number, where it is zero (the Nyquist wavenumber has no negative partner).
#pragma once
#include "fsm_passes.cuh"

namespace fsm {

#define FSM_MAP_MAX_TERMS 32
#define FSM_MAP_MAX_CH 6

template <typename T>
struct MapTerms {
    int n_terms, c_in, c_out, dealias;
    signed char out_ch[FSM_MAP_MAX_TERMS], in_ch[FSM_MAP_MAX_TERMS];
    signed char pw[FSM_MAP_MAX_TERMS][3];
    signed char inv_lap[FSM_MAP_MAX_TERMS];
    T coef[FSM_MAP_MAX_TERMS];
};

template <typename T>
__device__ __forceinline__ T ipow(T x, int p) {
    T r = T(1);
    for (int i = 0; i < p; ++i) r *= x;
    return r;
}

template <typename T>
__global__ void __launch_bounds__(256) k_spectral_map(Geom<T> g, MapTerms<T> mt, const cplx<T>* __restrict__ in,
                                                       cplx<T>* __restrict__ out, long total /* B * nmodes */) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long b = i / g.nmodes;
    long mode = i % g.nmodes;
    int q[3] = {0, 0, 0};
    if (g.ndim == 1) {
        q[0] = (int)mode;
    } else if (g.ndim == 2) {
        q[0] = (int)(mode % g.n[0]);
        q[1] = (int)(mode / g.n[0]);
    } else {
        q[0] = (int)(mode % g.n[0]);
        long r = mode / g.n[0];
        q[2] = (int)(r % g.nh);
        q[1] = g.ky0 + g.kys * (int)(r / g.nh);   // slab decomposition: cyclic ky ownership
    }
    T k[3] = {T(0), T(0), T(0)};
    bool nyq[3] = {false, false, false};
    bool kept = true;
    T k2 = T(0);
    for (int a = 0; a < g.ndim; ++a) {
        k[a] = g.dkraw[a][q[a]];
        nyq[a] = (g.n[a] % 2 == 0) && (q[a] == g.n[a] / 2);
        const int m = (q[a] <= g.n[a] / 2) ? q[a] : g.n[a] - q[a];
        kept = kept && (m <= g.kmax[a]);
        k2 += k[a] * k[a];
    }
    const T lap = -k2;                                   // sum_a (i k_a)^2, mesh.py:406-410
    const T inv = (lap == T(0)) ? T(1) : T(1) / lap;     // mesh.py:412-418
    cplx<T> u[FSM_MAP_MAX_CH], acc[FSM_MAP_MAX_CH];
    const bool zero_in = mt.dealias && !kept;
    FSM_UNROLL
    for (int c = 0; c < FSM_MAP_MAX_CH; ++c) {
        acc[c] = mk<T>(T(0), T(0));
        u[c] = mk<T>(T(0), T(0));
        if (c < mt.c_in && !zero_in) u[c] = in[(b * mt.c_in + c) * g.nmodes + mode];
    }
    for (int t = 0; t < mt.n_terms; ++t) {
        int odd = 0, tot = 0;
        T mag = mt.coef[t];
        for (int a = 0; a < g.ndim; ++a) {
            const int p = mt.pw[t][a];
            tot += p;
            if (nyq[a]) odd += p;
            mag *= ipow<T>(k[a], p);
        }
        if (odd & 1) continue;
        for (int j = 0; j < mt.inv_lap[t]; ++j) mag *= inv;
        cplx<T> v = mk<T>(T(0), T(0));
        FSM_UNROLL
        for (int c = 0; c < FSM_MAP_MAX_CH; ++c)
            if (c == mt.in_ch[t]) v = u[c];
        // i^tot * mag * v
        cplx<T> w;
        switch (tot & 3) {
            case 0: w = mk<T>(mag * v.x, mag * v.y); break;
            case 1: w = mk<T>(-mag * v.y, mag * v.x); break;
            case 2: w = mk<T>(-mag * v.x, -mag * v.y); break;
            default: w = mk<T>(mag * v.y, -mag * v.x); break;
        }
        FSM_UNROLL
        for (int c = 0; c < FSM_MAP_MAX_CH; ++c)
            if (c == mt.out_ch[t]) { acc[c].x += w.x; acc[c].y += w.y; }
    }
    FSM_UNROLL
    for (int c = 0; c < FSM_MAP_MAX_CH; ++c)
        if (c < mt.c_out) out[(b * mt.c_out + c) * g.nmodes + mode] = acc[c];
}

// Zero the modes of a rot-half state outside the dealiasing box, in place (no read). The reference does this to the
// integrator's own stage state when NSPressureConvection carries an external force (`u_fft *= low_pass_filter()` on the
// caller's tensor, dedicated/_navier_stokes.py:241).
template <typename T>
__global__ void __launch_bounds__(256) k_mask_state(Geom<T> g, cplx<T>* __restrict__ state, long total /* fields * nmodes */) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    long mode = i % g.nmodes;
    int q[3] = {0, 0, 0};
    if (g.ndim == 1) {
        q[0] = (int)mode;
    } else if (g.ndim == 2) {
        q[0] = (int)(mode % g.n[0]);
        q[1] = (int)(mode / g.n[0]);
    } else {
        q[0] = (int)(mode % g.n[0]);
        long r = mode / g.n[0];
        q[2] = (int)(r % g.nh);
        q[1] = g.ky0 + g.kys * (int)(r / g.nh);
    }
    bool kept = true;
    for (int a = 0; a < g.ndim; ++a) {
        const int m = (q[a] <= g.n[a] / 2) ? q[a] : g.n[a] - q[a];
        kept = kept && (m <= g.kmax[a]);
    }
    if (!kept) state[i] = mk<T>(T(0), T(0));
}

// Symmetric products of the channels of a physical field: out[b][idx(i,j)] = u[b][i] * u[b][j] for i <= j, row-major
// pairs ((0,0), (0,1), .., (1,1), ..). The point-wise part of _ConservativeConvectionCore
// (generic/_conservative_convection.py:23-25), which forms all C*C products; the transforms and the divergence run
// on the fused passes and the spectral map.
template <typename T>
__global__ void __launch_bounds__(256) k_sym_outer(const T* __restrict__ u, T* __restrict__ out, int C, long npts, long total /* B * npts */) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long b = i / npts, x = i % npts;
    T v[FSM_MAP_MAX_CH];
    FSM_UNROLL
    for (int c = 0; c < FSM_MAP_MAX_CH; ++c) v[c] = (c < C) ? u[(b * C + c) * npts + x] : T(0);
    const int npair = C * (C + 1) / 2;
    int idx = 0;
    for (int a = 0; a < C; ++a)
        for (int c = a; c < C; ++c, ++idx) out[(b * npair + idx) * npts + x] = v[a] * v[c];
}

// out = base + sum_j coef_j * term_j over complex arrays: the state combinations of the explicit Runge-Kutta family
// (`x_t + dt * sum([a_i * k for ...])`, integrator/_rk.py:43-58) without a chain of separate axpy kernels.
#define FSM_LINCOMB_MAX 8
template <typename T>
struct LinComb {
    int n;
    const cplx<T>* term[FSM_LINCOMB_MAX];
    T coef[FSM_LINCOMB_MAX];
};
template <typename T>
__global__ void __launch_bounds__(256) k_lincomb(cplx<T>* __restrict__ out, const cplx<T>* __restrict__ base, LinComb<T> lc, long total) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    cplx<T> acc = base[i];
    FSM_UNROLL
    for (int j = 0; j < FSM_LINCOMB_MAX; ++j)
        if (j < lc.n) acc = cfma_s(lc.coef[j], lc.term[j][i], acc);
    out[i] = acc;
}

}  // namespace fsm

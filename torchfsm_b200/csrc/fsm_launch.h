// Launch layer between the plan (fsm_plan.cu) and the per-size kernel instantiations
// (fsm_kernels.cu, compiled once per line length N with -DFSM_N=<N>).
#pragma once
#include "fsm_passes.cuh"

namespace fsm {

template <typename T>
struct IxArgs {
    Geom<T> g;
    const cplx<T>* state;
    cplx<T>* w1;
    long state_bstride, w1_fstride, in_t_stride, in_o_stride, out_o_stride, out_e_stride;
    int n_t, n_outer, nbc;
    Blk eb = {30, 0, 30, 0, 0, 0, 0};
    Peers pe = {{nullptr}, 0, 0};
};
template <typename T>
struct MidArgs {
    Geom<T> g;
    const cplx<T>* in;
    cplx<T>* out;
    long in_fstride, out_fstride, in_t_stride, in_o_stride, out_o_stride, out_e_stride;
    int nfi, n_t, n_outer, nb;
    MidSpec spec;
    Blk ib = {30, 0, 30, 0, 0, 0, 0}, eb = {30, 0, 30, 0, 0, 0, 0};
    Peers pe = {{nullptr}, 0, 0};
};
template <typename T>
struct PhysArgs {
    Geom<T> g;
    const cplx<T>* win;
    cplx<T>* wout;
    const T* phys_in;
    T* phys_out;
    long win_fstride, wout_fstride, in_t_stride, in_o_stride, out_o_stride, out_e_stride;
    int n_t, n_outer, nb;
};
template <typename T>
struct FxArgs {
    Geom<T> g;
    const cplx<T>* win;
    long win_fstride;
    Combine<T> cb;
    FxEpilogue<T> ep;
    int nlines, b0, nb;
    Blk ib = {30, 0, 30, 0, 0, 0, 0};
    long line_stride = 0;  // 0 = N (contiguous lines)
};

template <typename T>
struct Step1dArgs {
    Geom<T> g;
    StageList<T> sl;
    FxEpilogue<T> ep;
    int n_steps, nb;
};

// One table per supported line length; entries return 0 or a negative errno.
template <typename T>
struct LaunchTable {
    int N;
    int (*ix)(int prog, const IxArgs<T>&, cudaStream_t);
    int (*mid)(int dir, const MidArgs<T>&, cudaStream_t);
    int (*phys)(int prog, int ndim, const PhysArgs<T>&, cudaStream_t);
    int (*fx)(int C, const FxArgs<T>&, cudaStream_t);
    int (*step1d)(const Step1dArgs<T>&, cudaStream_t);
    int (*line1d)(int mode, const void* in, void* out, long nfields, cudaStream_t);
    int (*prepare)();   // once per device: fill the static twiddle tables of this line length (synchronous)
};

template <typename T>
const LaunchTable<T>* launch_table(int N);  // nullptr when N is unsupported

}  // namespace fsm

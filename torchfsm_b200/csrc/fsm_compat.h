// Compatibility layer: the kernels in this directory compile either with nvcc for
// sm_100a (the product, libfsm_b200.so) or with g++ under -DFSM_EMU (a fiber-based
// host emulator used ONLY by the CPU test-suite to validate kernel logic without a
// GPU; tests/emu/). The product never links or loads the emulator build.
#pragma once
#include <cstddef>
#include <cstdint>

// Every lambda of the kernels is force-inlined: left to the inliner's cost model, a lambda body shared by several
// kernel instantiations of one translation unit may stay a call, its captured register arrays then live in local
// memory and its shared-memory pointers turn generic (measured: the same last-axis kernel 0.75 ms or 1.9 ms
// depending on what else the file instantiates).
#define FSM_INLINE_LAMBDA __attribute__((always_inline))

#ifdef FSM_EMU
// ----------------------------------------------------------------------------------
// Host emulation: one CUDA thread == one ucontext fiber; __syncthreads / __syncwarp /
// named barriers are cooperative yields handled by the scheduler in fsm_emu.cpp.
// ----------------------------------------------------------------------------------
#include <cmath>
#include <cstring>
#include <functional>
#include <algorithm>

struct fsm_dim3 {
    unsigned x, y, z;
    fsm_dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef fsm_dim3 dim3;
typedef void* cudaStream_t;

namespace fsm_emu {
struct Ctx {
    fsm_dim3 tid, bid, bdim, gdim;
    char* smem;
};
extern thread_local Ctx g_ctx;
void barrier_block();
void barrier_warp();
void barrier_named(int id, int count);
void launch(fsm_dim3 grid, fsm_dim3 block, size_t smem_bytes, const std::function<void()>& body);
}  // namespace fsm_emu

#define threadIdx (fsm_emu::g_ctx.tid)
#define blockIdx (fsm_emu::g_ctx.bid)
#define blockDim (fsm_emu::g_ctx.bdim)
#define gridDim (fsm_emu::g_ctx.gdim)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __syncthreads() fsm_emu::barrier_block()
#define __syncwarp() fsm_emu::barrier_warp()
#define FSM_NAMED_BARRIER(id, count) fsm_emu::barrier_named((id), (count))
#define FSM_DYN_SMEM(name) char* name = fsm_emu::g_ctx.smem
#define FSM_LAUNCH(kernel, grid, block, smem, stream, ...) \
    fsm_emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define FSM_UNROLL
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float fsm_fma(float a, float b, float c) { return std::fma(a, b, c); }
static inline double fsm_fma(double a, double b, double c) { return std::fma(a, b, c); }
#define FSM_HD
#define FSM_PIN(ptr) (void)0
#define FSM_CUDA_CHECK_LAUNCH() 0

#else
// ----------------------------------------------------------------------------------
// Real CUDA (sm_100a)
// ----------------------------------------------------------------------------------
#include <cuda_runtime.h>
#define FSM_NAMED_BARRIER(id, count) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory")
#define FSM_DYN_SMEM(name) extern __shared__ __align__(16) char name[]
#define FSM_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define FSM_UNROLL _Pragma("unroll")
static __device__ __forceinline__ float fsm_fma(float a, float b, float c) { return fmaf(a, b, c); }
static __device__ __forceinline__ double fsm_fma(double a, double b, double c) { return fma(a, b, c); }
#define FSM_HD __host__ __device__
// Materialise a base pointer in a register pair here and now. Without it the compiler sinks the 64-bit
// address arithmetic (strides are runtime longs) into every predicated element access.
#define FSM_PIN(ptr)                    \
    do {                                \
        asm volatile("" : "+l"(ptr));   \
        __builtin_assume(__isGlobal(ptr)); /* keep LDG/STG: the asm hides the address space */ \
    } while (0)
#endif

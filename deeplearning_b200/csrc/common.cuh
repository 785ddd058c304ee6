// common.cuh — shared device/host helpers of libdlb200 (sm_100a).
#pragma once

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#ifndef DLB_CUDA_EMU
#include <cuda_runtime.h>
#endif

#include "../../include/dlb200_spec.h"

namespace dlb {

// CUDA failures become exceptions carrying cudaGetErrorString + file + line, like the reference's
// gpuErrchk (DeepLearningCuda/Math/CudaUtils.cpp:24-30); the C ABI turns them into status codes.
struct cuda_error : std::runtime_error { using std::runtime_error::runtime_error; };

inline void cuda_check(cudaError_t e, const char* file, int line) {
    if (e != cudaSuccess)
        throw cuda_error(std::string(cudaGetErrorString(e)) + "; " + file + ":" + std::to_string(line));
}
#define DLB_CUDA(x) ::dlb::cuda_check((x), __FILE__, __LINE__)

// Kernel launch.  `kernel` is passed in parentheses so that template argument lists with commas survive the
// macro; every launch is counted (bench.py reports the counter as gpu_launches).  When per-kernel timing is
// switched on (dlb_timing_enable, used by bench.py's roofline pass only) each launch is bracketed by CUDA events on
// the launching stream and attributed to the current tag ("<layer>:<operator>").
extern std::atomic<long long> g_launch_count;
extern bool g_timing_on;
extern thread_local const char* g_tag;
extern thread_local int g_tag_layer;
void timing_begin(cudaStream_t stream, const char* kernel_text);
void timing_end(cudaStream_t stream);

// Opt a kernel in to more than 48 KB of dynamic shared memory (once per kernel and size, not per launch).
void allow_big_smem(const void* kernel, size_t bytes);

struct TagScope {
    const char* prev;
    explicit TagScope(const char* t) : prev(g_tag) { g_tag = t; }
    ~TagScope() { g_tag = prev; }
};

#define DLB_UNPAREN(...) __VA_ARGS__
#ifdef DLB_CUDA_EMU
#define DLB_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...)                         \
    do {                                                                                \
        ::dlb::g_launch_count.fetch_add(1, std::memory_order_relaxed);                  \
        ::cuemu::launch_smem(DLB_UNPAREN kernel, (grid), (block), (smem), __VA_ARGS__); \
    } while (0)
// dynamic shared memory of the running block, typed
#define DLB_DYNAMIC_SMEM(type, name) type* name = reinterpret_cast<type*>(::cuemu::dyn_smem())
#else
#define DLB_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...)                         \
    do {                                                                                \
        ::dlb::g_launch_count.fetch_add(1, std::memory_order_relaxed);                  \
        if (::dlb::g_timing_on) ::dlb::timing_begin((stream), #kernel);                 \
        auto dlb_kernel_ = DLB_UNPAREN kernel;                                          \
        if ((smem) > 48 * 1024) ::dlb::allow_big_smem((const void*)dlb_kernel_, (size_t)(smem));                   \
        dlb_kernel_<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                \
        DLB_CUDA(cudaGetLastError());                                                   \
        if (::dlb::g_timing_on) ::dlb::timing_end((stream));                            \
    } while (0)
#define DLB_DYNAMIC_SMEM(type, name)                                  \
    extern __shared__ __align__(16) unsigned char dlb_dyn_smem_raw[]; \
    type* name = reinterpret_cast<type*>(dlb_dyn_smem_raw)
#endif
#define DLB_LAUNCH(kernel, grid, block, stream, ...) DLB_LAUNCH_SMEM(kernel, grid, block, 0, stream, __VA_ARGS__)

// Packed fp32 pairs for the FFMA2 pipe (`fma.rn.f32x2`, sm_100+): one instruction performs two IEEE fused multiply-adds
// on a 64-bit register pair; the scalar operand is broadcast to both halves (SASS: FFMA2 Rd, Ra.F32x2.HI_LO, Rb.F32, Rc).
// The emulator build carries the pair as a struct.
#ifdef DLB_CUDA_EMU
struct f32x2 { float lo, hi; };
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { return f32x2{lo, hi}; }
__device__ __forceinline__ float lo_of(f32x2 v) { return v.lo; }
__device__ __forceinline__ float hi_of(f32x2 v) { return v.hi; }
__device__ __forceinline__ f32x2 ffma2_bcast(f32x2 a, float w, f32x2 c) { return f32x2{fmaf(a.lo, w, c.lo), fmaf(a.hi, w, c.hi)}; }
#else
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo_of(f32x2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo; }
__device__ __forceinline__ float hi_of(f32x2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return hi; }
__device__ __forceinline__ f32x2 ffma2_bcast(f32x2 a, float w, f32x2 c) {
    f32x2 d;
    asm("{ .reg .b64 ww; mov.b64 ww, {%2, %2}; fma.rn.f32x2 %0, %1, ww, %3; }" : "=l"(d) : "l"(a), "f"(w), "l"(c));
    return d;
}
#endif
struct alignas(16) f32x2_quad { f32x2 a, b; };   // two pairs = one 16-byte shared-memory load

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may start while
// its predecessor in the stream is still running; it must not touch the predecessor's results before pdl_wait().  A
// predecessor calls pdl_trigger() as soon as the successor's prologue may begin (here: at once - the successor's CTAs only
// get an SM when this grid's CTAs leave it).  Both are no-ops in launches without the attribute.
#ifndef DLB_CUDA_EMU
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_trigger() {}
#endif

template <class T> __host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200

// -----------------------------------------------------------------------------------------------------------------
// Activations (value + auxiliary derivative), the reference's formulas:
//   sigmoid  DeepLearning/Math/ActivationFunctionHelper.cpp:84-95     tanh :103-114     relu :64-76
//   linear   ActivationFunctionFactory.h:38 through the dual-number path (value x, derivative 1)
// Softmax is a row operation and lives in its own kernels (layers.cuh).
// -----------------------------------------------------------------------------------------------------------------
template <class R> __device__ __forceinline__ R dl_exp(R x);
template <> __device__ __forceinline__ float dl_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double dl_exp<double>(double x) { return exp(x); }
template <class R> __device__ __forceinline__ R dl_log(R x);
template <> __device__ __forceinline__ float dl_log<float>(float x) { return logf(x); }
template <> __device__ __forceinline__ double dl_log<double>(double x) { return log(x); }
template <class R> __device__ __forceinline__ R dl_tanh(R x);
template <> __device__ __forceinline__ float dl_tanh<float>(float x) { return tanhf(x); }
template <> __device__ __forceinline__ double dl_tanh<double>(double x) { return tanh(x); }

template <class R>
__device__ __forceinline__ R apply_activation(int act, R z, R& deriv) {
    switch (act) {
    case DLB_ACT_SIGMOID: { const R v = R(1) / (R(1) + dl_exp<R>(-z)); deriv = v * (R(1) - v); return v; }
    case DLB_ACT_TANH: { const R v = dl_tanh<R>(z); deriv = R(1) - v * v; return v; }
    case DLB_ACT_RELU: { deriv = z > R(0) ? R(1) : R(0); return z > R(0) ? z : R(0); }
    default: deriv = R(1); return z;  // LINEAR; SOFTMAX rows are finished by softmax_forward_kernel
    }
}

template <class R> struct real_limits;
template <> struct real_limits<float> { __host__ __device__ static constexpr float max() { return 3.402823466e+38f; } };
template <> struct real_limits<double> { __host__ __device__ static constexpr double max() { return 1.7976931348623157e+308; } };

}  // namespace dlb

// tools/cudaemu/cuda_emu.h — DEVELOPMENT-TIME LOGIC CHECKER, NOT A BACKEND.
//
// The build container has nvcc but no GPU, and GPU time is rationed.  This header lets the plain-CUDA-C++ kernels
// of deeplearning_b200/csrc (no inline PTX) be compiled with g++ and executed on the CPU — one fiber per CUDA
// thread, blocks run one after another — so that indexing, barriers and host-side plumbing can be debugged before
// a GPU run.  It is compiled only by tools/cudaemu/build_emu.sh into tools/cudaemu/_build/libdlb200_emu.so, which
// the product never loads: deeplearning_b200/_capi.py opens libdlb200.so and nothing else, __graft_entry__.build()
// does not build this, and no test marked `gpu` or `not gpu` uses it unless the developer passes `--emu` to pytest.
// Results from it are never reported as measurements or as parity evidence.
#pragma once

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#ifndef __align__
#define __align__(n) __attribute__((aligned(n)))
#endif

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };

inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)0x1; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }

namespace cuemu {

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;      // allocated once, reused by every block
    uint3 tid{};
    bool done = false;
    int wait_kind = 0;          // 0 runnable, 1 block barrier, 2 warp barrier
    unsigned long long wait_gen = 0;
};

struct WarpState {
    int alive = 0, arrived = 0;
    unsigned long long gen = 0;
    double slot[2][32];
};

constexpr size_t kStackBytes = 64 * 1024;
inline std::vector<Fiber>& fiber_pool() { static std::vector<Fiber> pool(1024); return pool; }

struct BlockState {
    Fiber* fibers = nullptr;
    std::vector<WarpState> warps;
    ucontext_t sched;
    Fiber* cur = nullptr;
    int alive = 0, arrived = 0;
    unsigned long long gen = 0;
    void (*body)(void*) = nullptr;
    void* arg = nullptr;
};

inline BlockState*& block() { static BlockState* b = nullptr; return b; }
inline uint3& block_idx() { static uint3 v{}; return v; }
inline dim3& block_dim() { static dim3 v; return v; }
inline dim3& grid_dim() { static dim3 v; return v; }

inline void yield_to_scheduler() { BlockState* b = block(); swapcontext(&b->cur->ctx, &b->sched); }

inline void sync_block() {
    BlockState* b = block();
    Fiber* f = b->cur;
    f->wait_kind = 1; f->wait_gen = b->gen;
    if (++b->arrived == b->alive) { b->arrived = 0; ++b->gen; }
    yield_to_scheduler();
}

inline int linear_tid(const Fiber* f) { const dim3& d = block_dim(); return (int)(f->tid.x + d.x * (f->tid.y + d.y * f->tid.z)); }

template <class T>
inline T warp_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= sizeof(double), "shuffle payload too large");
    BlockState* b = block();
    Fiber* f = b->cur;
    const int lt = linear_tid(f), lane = lt % 32;
    WarpState& w = b->warps[(size_t)lt / 32];
    const unsigned long long my_gen = w.gen;
    std::memcpy(&w.slot[my_gen & 1][lane], &v, sizeof(T));
    f->wait_kind = 2; f->wait_gen = my_gen;
    if (++w.arrived == w.alive) { w.arrived = 0; ++w.gen; }
    yield_to_scheduler();
    T r;
    std::memcpy(&r, &w.slot[my_gen & 1][src_lane & 31], sizeof(T));
    return r;
}

inline void fiber_entry() {
    BlockState* b = block();
    b->body(b->arg);
    Fiber* f = b->cur;
    f->done = true;
    const int lt = linear_tid(f);
    WarpState& w = b->warps[(size_t)lt / 32];
    --b->alive; --w.alive;
    // a thread that exits releases barriers the others are waiting on
    if (b->alive > 0 && b->arrived == b->alive) { b->arrived = 0; ++b->gen; }
    if (w.alive > 0 && w.arrived == w.alive) { w.arrived = 0; ++w.gen; }
    swapcontext(&f->ctx, &b->sched);
}

inline void run_block(void (*body)(void*), void* arg) {
    const dim3 bd = block_dim();
    const int nt = (int)(bd.x * bd.y * bd.z);
    if (nt > 1024) { std::fprintf(stderr, "cuemu: block too large\n"); std::abort(); }
    BlockState bs;
    bs.fibers = fiber_pool().data();
    bs.warps.resize((size_t)(nt + 31) / 32);
    bs.body = body; bs.arg = arg; bs.alive = nt;
    block() = &bs;
    for (int t = 0; t < nt; ++t) {
        Fiber& f = bs.fibers[(size_t)t];
        if (!f.stack) f.stack = static_cast<char*>(std::malloc(kStackBytes));
        f.done = false; f.wait_kind = 0; f.wait_gen = 0;
        f.tid = uint3{(unsigned)t % bd.x, ((unsigned)t / bd.x) % bd.y, (unsigned)t / (bd.x * bd.y)};
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStackBytes;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        bs.warps[(size_t)t / 32].alive++;
    }
    while (bs.alive > 0) {
        bool progressed = false;
        for (int t = 0; t < nt; ++t) {
            Fiber& f = bs.fibers[(size_t)t];
            if (f.done) continue;
            if (f.wait_kind == 1 && bs.gen == f.wait_gen) continue;
            if (f.wait_kind == 2 && bs.warps[(size_t)t / 32].gen == f.wait_gen) continue;
            f.wait_kind = 0;
            bs.cur = &f;
            swapcontext(&bs.sched, &f.ctx);
            progressed = true;
        }
        if (!progressed) { std::fprintf(stderr, "cuemu: deadlock (divergent barrier)\n"); std::abort(); }
    }
    block() = nullptr;
}

inline std::vector<double>& dyn_smem_storage() { static std::vector<double> v; return v; }
inline void* dyn_smem() { return dyn_smem_storage().data(); }

template <class K, class... A>
void launch(K kernel, dim3 grid, dim3 blk, A... args);

template <class K, class... A>
void launch_smem(K kernel, dim3 grid, dim3 blk, size_t smem, A... args) {
    dyn_smem_storage().assign(smem / sizeof(double) + 2, 0.0);
    launch(kernel, grid, blk, args...);
}

template <class K, class... A>
void launch(K kernel, dim3 grid, dim3 blk, A... args) {
    grid_dim() = grid; block_dim() = blk;
    auto body = [&]() { kernel(args...); };
    using Body = decltype(body);
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                block_idx() = uint3{x, y, z};
                run_block([](void* p) { (*static_cast<Body*>(p))(); }, &body);
            }
}

}  // namespace cuemu

#define threadIdx (::cuemu::block()->cur->tid)
#define blockIdx (::cuemu::block_idx())
#define blockDim (::cuemu::block_dim())
#define gridDim (::cuemu::grid_dim())

inline void __syncthreads() { ::cuemu::sync_block(); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    const int lane = ::cuemu::linear_tid(::cuemu::block()->cur) % 32;
    return ::cuemu::warp_exchange(v, lane ^ lane_mask);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, int delta) {
    const int lane = ::cuemu::linear_tid(::cuemu::block()->cur) % 32;
    return ::cuemu::warp_exchange(v, lane + delta < 32 ? lane + delta : lane);
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return ::cuemu::warp_exchange(v, src); }
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }

using std::isinf;
using std::isnan;
using std::max;
using std::min;
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) float2 { float x, y; };
struct alignas(8) int2 { int x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline void __threadfence_system() {}
inline void __threadfence() {}
inline void __nanosleep(unsigned) {}
inline void __trap() { std::abort(); }

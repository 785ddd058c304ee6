// tools/ffma2_probe.cu — issue rate of the FFMA2 forms the convolution kernels use (sm_100a): the accumulator pair and a per-thread
// scalar (broadcast) times a pair that is either a regular register pair or, when the compiler proves it warp-uniform (taps read
// from shared memory at a uniform address), a UNIFORM register pair.  12 warps per SM like the kernels, 32 independent accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/ffma2_probe tools/ffma2_probe.cu && tools/_build/ffma2_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 ffma2_bcast(f32x2 a, float w, f32x2 c) {
    f32x2 d; asm("{ .reg .b64 ww; mov.b64 ww, {%2, %2}; fma.rn.f32x2 %0, %1, ww, %3; }" : "=l"(d) : "l"(a), "f"(w), "l"(c)); return d;
}
template <bool UNIFORM_TAPS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(const float* __restrict__ taps_g, const float* __restrict__ xs_g, float* sink, int iters) {
    __shared__ float wsm[64];
    if (threadIdx.x < 50) wsm[threadIdx.x] = taps_g[threadIdx.x];
    __syncthreads();
    f32x2 acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = pack2(0.f, (float)i);
    float x[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = xs_g[(threadIdx.x * 12 + i) & 1023];
    for (int it = 0; it < iters; ++it) {
        f32x2 w[25];   // re-read per pass like the kernels' filter-pair loop (uniform address: ptxas moves the pairs to uniform registers)
#pragma unroll
        for (int t = 0; t < 25; ++t) {
            w[t] = pack2(wsm[(t + it) & 31], wsm[(25 + t + it) & 63]);
            if (!UNIFORM_TAPS) asm volatile("mov.b64 %0, %0;" : "+l"(w[t]));   // opaque: keeps the pair in regular registers
        }
#pragma unroll
        for (int t = 0; t < 25; ++t)
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = ffma2_bcast(w[t], x[(i + t) % 12], acc[i]);
    }
    f32x2 s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s ^= acc[i];
    if (s == 0x123456789ull) sink[0] = 1.f;
}
template <class F> static double time_ms(F&& launch) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    return best;
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, iters = 400;
    float *taps, *xs, *sink; CK(cudaMalloc(&taps, 256)); CK(cudaMalloc(&xs, 4096)); CK(cudaMalloc(&sink, 16));
    CK(cudaMemset(taps, 0, 256)); CK(cudaMemset(xs, 0, 4096));
    auto report = [&](const char* name, double ms, int warps) {
        const double ffma2 = 25.0 * 32 * iters * warps * sms;   // warp-level FFMA2 instructions
        const double clk = prop.clockRate * 1e3;                // (nominal; the ratio between the variants is what matters)
        printf("%-44s %8.3f ms  %6.1f TFLOP/s  %.3f FFMA2 / clk / scheduler (at %.0f MHz nominal)\n", name, ms, ffma2 * 32 * 4 / (ms * 1e-3) / 1e12,
               ffma2 / (sms * 4) / (ms * 1e-3 * clk), clk / 1e6);
    };
    report("taps in uniform registers, 12 warps / SM", time_ms([&] { probe<true, 12><<<sms, 384>>>(taps, xs, sink, iters); }), 12);
    report("taps in regular registers, 12 warps / SM", time_ms([&] { probe<false, 12><<<sms, 384>>>(taps, xs, sink, iters); }), 12);
    report("taps in uniform registers,  8 warps / SM", time_ms([&] { probe<true, 8><<<sms, 256>>>(taps, xs, sink, iters); }), 8);
    report("taps in regular registers,  8 warps / SM", time_ms([&] { probe<false, 8><<<sms, 256>>>(taps, xs, sink, iters); }), 8);
    report("taps in uniform registers,  4 warps / SM", time_ms([&] { probe<true, 4><<<sms, 128>>>(taps, xs, sink, iters); }), 4);
    report("taps in regular registers,  4 warps / SM", time_ms([&] { probe<false, 4><<<sms, 128>>>(taps, xs, sink, iters); }), 4);
    return 0;
}

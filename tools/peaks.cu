// tools/peaks.cu — measured arithmetic peaks of this B200 for the roofline denominators MEASURED_PEAKS.json does not have
// (it holds the HBM copy bandwidth and the bf16 GEMM throughput): dense TF32 on the 5th-generation tensor cores the way the
// library issues it (tcgen05.mma kind::tf32, cta_group::1, M = 128, operands in shared memory), packed fp32 (FFMA2), scalar
// fp32 (FFMA) and fp64 (DFMA).  Writes one JSON object to stdout; profiles/peaks_r2.json is a copy of a run.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/peaks tools/peaks.cu && tools/_build/peaks
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {   // K-major SWIZZLE_128B, SBO = 1024 B
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// ---- TF32 tensor peak: every CTA (one per SM) issues `iters` x 4 MMAs of M = 128, N = n, K = 8 on operands that stay in shared memory,
// alternating between two accumulators, and waits for the last commit.
template <int N>
__global__ void __launch_bounds__(128, 1) tf32_peak_kernel(int iters, float* sink) {
    extern __shared__ unsigned char raw[];
    unsigned char* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* f = reinterpret_cast<float*>(sm);
    for (int i = tid; i < (128 + N) * 32; i += 128) f[i] = 1.0f + (float)((i * 2654435761u) >> 20) * 1e-4f;   // A tile [128][32] | B tile [N][32]
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(sm), b0 = a0 + 128 * 128;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = make_desc(a0 + k * 32), db = make_desc(b0 + k * 32);
                const uint32_t d = tm + (uint32_t)((it & 1) * 256);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(it > 1 ? 1u : 0u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tm + ((uint32_t)(warp * 32) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (v == 0x12345678u) sink[0] = 1.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// ---- CUDA-core peaks: 16 independent accumulators per thread
__global__ void __launch_bounds__(512, 2) ffma_peak_kernel(int iters, float* sink) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0f + threadIdx.x * 1e-6f + i;
    const float x = 1.0000001f, y = 1e-7f;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
}
__global__ void __launch_bounds__(512, 2) ffma2_peak_kernel(int iters, float* sink) {
    unsigned long long a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float v = 1.0f + threadIdx.x * 1e-6f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(v)); }
    unsigned long long x, y;
    { const float fx = 1.0000001f, fy = 1e-7f; asm("mov.b64 %0, {%1, %1};" : "=l"(x) : "f"(fx)); asm("mov.b64 %0, {%1, %1};" : "=l"(y) : "f"(fy)); }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(x), "l"(y));
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= a[i];
    if (s == 0x123456789ull) sink[0] = 1.f;
}
__global__ void __launch_bounds__(512, 2) dfma_peak_kernel(int iters, double* sink) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    const double x = 1.00000000001, y = 1e-12;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
}

// fp64 tensor path available on sm_100a: mma.sync.m8n8k4.f64 (DMMA), 8 independent accumulator tiles per warp
__global__ void __launch_bounds__(512, 2) dmma_peak_kernel(int iters, double* sink) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9 + i; c[i][1] = 1.0 + i; }
    const double a = 1.00000000001, b = 0.99999999999;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

template <class F>
static double time_ms(F&& launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float* sink; double* dsink;
    CK(cudaMalloc(&sink, 16)); CK(cudaMalloc(&dsink, 16));
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    {
        const int iters = 20000;
        const size_t smem = (128 + 256) * 128 + 2048;
        CK(cudaFuncSetAttribute(tf32_peak_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(tf32_peak_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(tf32_peak_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const double ms256 = time_ms([&] { tf32_peak_kernel<256><<<sms, 128, smem>>>(iters, sink); });
        const double ms128 = time_ms([&] { tf32_peak_kernel<128><<<sms, 128, smem>>>(iters, sink); });
        const double ms48 = time_ms([&] { tf32_peak_kernel<48><<<sms, 128, smem>>>(iters, sink); });
        auto tf = [&](int n, double ms) { return 2.0 * 128 * n * 8 * 4.0 * iters * sms / (ms * 1e-3) / 1e12; };
        printf(", \"tf32_tcgen05_cta1_m128_n256_tflops\": %.1f, \"tf32_tcgen05_cta1_m128_n128_tflops\": %.1f, \"tf32_tcgen05_cta1_m128_n48_tflops\": %.1f",
               tf(256, ms256), tf(128, ms128), tf(48, ms48));
        printf(", \"tf32_3x_equivalent_fp32_tflops\": %.1f", tf(256, ms256) / 3.0);
    }
    {
        const int iters = 4000;
        const int blocks = sms * 2;
        const double total = 2.0 * 16.0 * iters * 512.0 * blocks;
        const double ms1 = time_ms([&] { ffma_peak_kernel<<<blocks, 512>>>(iters, sink); });
        const double ms2 = time_ms([&] { ffma2_peak_kernel<<<blocks, 512>>>(iters, sink); });
        const double ms3 = time_ms([&] { dfma_peak_kernel<<<blocks, 512>>>(iters, dsink); });
        printf(", \"ffma_tflops\": %.1f, \"ffma2_tflops\": %.1f, \"dfma_tflops\": %.2f", total / (ms1 * 1e-3) / 1e12, 2.0 * total / (ms2 * 1e-3) / 1e12, total / (ms3 * 1e-3) / 1e12);
        const double ms4 = time_ms([&] { dmma_peak_kernel<<<blocks, 512>>>(iters, dsink); });
        printf(", \"dmma_m8n8k4_tflops\": %.2f", 2.0 * 8 * 8 * 4 * 8.0 * iters * 16.0 * blocks / (ms4 * 1e-3) / 1e12);
    }
    printf(", \"how\": \"tools/peaks.cu: best of 5 launches, CUDA events; tensor: one CTA per SM issuing tcgen05.mma kind::tf32 cta_group::1 back to back on shared-memory operands; CUDA cores: 2 x 512 threads per SM, 16 independent accumulators per thread\"}\n");
    return 0;
}

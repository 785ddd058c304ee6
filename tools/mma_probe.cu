// tools/mma_probe.cu — development probe (not part of the product): checks on a B200 that tcgen05.mma kind::tf32 reads
// shared-memory operands the way the megakernel's descriptors say, for the three orientations it uses on ONE set of
// tiles of [rows][32 fp32] whose intra-row swizzle depends on the use: SWIZZLE_128B for K-major reads, SWIZZLE_128B_BASE32B
// for MN-major reads (the only MN-major layout 32-bit operands have: MN-major reads of SWIZZLE_128B and of un-swizzled
// tiles both returned zeros in earlier versions of this probe):
//   test 0  A K-major  (W1 [n][k] tiles),  B K-major  (pooled [b][k] tiles)       D[n][b]  = sum_k  W[n][k] X[b][k]
//   test 1  A MN-major (W1 tiles read as W^T: M = k, K = n),  B K-major (g [b][n]) D[k][b]  = sum_n  W[n][k] G[b][n]
//   test 2  A MN-major (g tiles read as G^T: M = n, K = b),   B MN-major (pooled tiles read as X^T: N = k, K = b)
//                                                                                  D[n][k]  = sum_b  G[b][n] X[b][k]
// Values are small integers (exact in TF32), so the device result must equal the host result exactly.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/mma_probe tools/mma_probe.cu && gpurun_out/mma_probe
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// start address, leading byte offset, stride byte offset (all >> 4), descriptor version 1, layout type in bits [61, 64):
// 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (the only MN-major layout of 32-bit operands)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

constexpr int RW = 104;   // rows of a W tile (n), 13 groups of 8
constexpr int RX = 40;    // rows of an X / G tile (b)
constexpr int TW = RW * 128, TX = RX * 128;   // tile bytes: [rows][32 fp32]
constexpr int NKT = 5;    // k tiles of the slice (160 k)
constexpr int NNT = 4;    // n tiles of G (128 n)

// Every operand is kept as tiles of [rows][32 fp32] (128-byte rows), in one of two intra-row swizzles:
//   K  (K-major use: MN = row, K = column)      16-byte chunk c of row r at chunk c ^ (r & 7)     SWIZZLE_128B
//   MN (MN-major use: K = row, MN = column)     32-byte chunk c of row r at chunk c ^ (r & 3)     SWIZZLE_128B_BASE32B
// smem: W_K [NKT] | W_MN [NKT] | X_K [NKT] | X_MN [NKT] | G_K [NNT] | G_MN [NNT]
constexpr int OFF_WK = 0, OFF_WM = OFF_WK + NKT * TW, OFF_XK = OFF_WM + NKT * TW, OFF_XM = OFF_XK + NKT * TX, OFF_GK = OFF_XM + NKT * TX,
              OFF_GM = OFF_GK + NNT * TX, SMEM = OFF_GM + NNT * TX + 2 * 1024 /*slack read by over-wide MMAs*/;

__global__ void __launch_bounds__(128, 1) probe(const float* W, const float* X, const float* G, float* D0, float* D1, float* D2) {
    extern __shared__ unsigned char raw[];
    unsigned char* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto put_k = [&](int off, int tile_bytes, int t, int r, int c, float v) {
        *reinterpret_cast<float*>(sm + off + t * tile_bytes + r * 128 + ((((c >> 2) ^ (r & 7))) << 4) + (c & 3) * 4) = v;
    };
    auto put_mn = [&](int off, int tile_bytes, int t, int r, int c, float v) {
        *reinterpret_cast<float*>(sm + off + t * tile_bytes + r * 128 + ((((c >> 3) ^ (r & 3))) << 5) + (c & 7) * 4) = v;
    };
    for (int i = tid; i < (SMEM - 1024) / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < RW * NKT * 32; i += 128) {
        const int n = i / (NKT * 32), k = i % (NKT * 32); const float v = n < 100 ? W[n * 160 + k] : 0.f;
        put_k(OFF_WK, TW, k / 32, n, k % 32, v); put_mn(OFF_WM, TW, k / 32, n, k % 32, v);
    }
    for (int i = tid; i < RX * NKT * 32; i += 128) {
        const int b = i / (NKT * 32), k = i % (NKT * 32); const float v = X[b * 160 + k];
        put_k(OFF_XK, TX, k / 32, b, k % 32, v); put_mn(OFF_XM, TX, k / 32, b, k % 32, v);
    }
    for (int i = tid; i < RX * NNT * 32; i += 128) {
        const int b = i / (NNT * 32), n = i % (NNT * 32); const float v = n < 100 ? G[b * 128 + n] : 0.f;
        put_k(OFF_GK, TX, n / 32, b, n % 32, v); put_mn(OFF_GM, TX, n / 32, b, n % 32, v);
    }
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
    const uint32_t base = smem_u32(sm);
    if (tid == 0) {
        // ---- test 0: D[n][b], columns [0, 48): K-major both, 18 k-steps (144 k)
        {
            const uint32_t id = make_idesc(128, 48, 0, 0);
            uint32_t acc = 0;
            for (int ks = 0; ks < 18; ++ks) {
                const int t = ks >> 2, j = ks & 3;
                const uint64_t a = make_desc(base + OFF_WK + t * TW + j * 32, 16, 1024, 2);
                const uint64_t b = make_desc(base + OFF_XK + t * TX + j * 32, 16, 1024, 2);
                umma(tm + 0, a, b, id, acc); acc = 1;
            }
        }
        // ---- test 1: D[k][b], two M tiles at columns [64, 112) and [128, 176): A = W^T MN-major (MN atoms = k tiles, LBO = tile stride;
        //      K atoms = groups of 4 n rows, SBO = 512), B = G K-major, 13 k-steps (n)
        for (int mt = 0; mt < 2; ++mt) {
            const uint32_t id = make_idesc(128, 48, 1, 0);
            uint32_t acc = 0;
            for (int ks = 0; ks < 13; ++ks) {
                const uint64_t a = make_desc(base + OFF_WM + (4 * mt) * TW + ks * 1024, TW, 512, 1);
                const uint64_t b = make_desc(base + OFF_GK + (ks >> 2) * TX + (ks & 3) * 32, 16, 1024, 2);
                umma(tm + 64 + 64 * mt, a, b, id, acc); acc = 1;
            }
        }
        // ---- test 3 (diagnostic): test 1, M tile 0, with LBO / SBO the other way round, columns [192, 240)
        {
            const uint32_t id = make_idesc(128, 48, 1, 0);
            uint32_t acc = 0;
            for (int ks = 0; ks < 13; ++ks) {
                const uint64_t a = make_desc(base + OFF_WM + ks * 1024, 512, TW, 1);
                const uint64_t b = make_desc(base + OFF_GK + (ks >> 2) * TX + (ks & 3) * 32, 16, 1024, 2);
                umma(tm + 192, a, b, id, acc); acc = 1;
            }
        }
        // ---- test 2: D[n][k], columns [256, 416): A = G^T MN-major, B = X^T MN-major, 5 k-steps (b)
        {
            const uint32_t id = make_idesc(128, 160, 1, 1);
            uint32_t acc = 0;
            for (int ks = 0; ks < 5; ++ks) {
                const uint64_t a = make_desc(base + OFF_GM + ks * 1024, TX, 512, 1);
                const uint64_t b = make_desc(base + OFF_XM + ks * 1024, TX, 512, 1);
                umma(tm + 256, a, b, id, acc); acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs
    {
        uint32_t ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // read back: warp w reads TMEM lanes 32 w .. 32 w + 31, 16 columns at a time
    auto read16 = [&](int col, uint32_t (&v)[16]) {
        const uint32_t ta = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)col;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(ta) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    };
    const int row = warp * 32 + lane;
    uint32_t v[16];
    for (int c = 0; c < 48; c += 16) { read16(c, v); for (int q = 0; q < 16; ++q) D0[row * 48 + c + q] = __uint_as_float(v[q]); }
    for (int mt = 0; mt < 2; ++mt)
        for (int c = 0; c < 48; c += 16) { read16(64 + 64 * mt + c, v); for (int q = 0; q < 16; ++q) D1[(mt * 128 + row) * 48 + c + q] = __uint_as_float(v[q]); }
    for (int c = 0; c < 48; c += 16) { read16(192 + c, v); for (int q = 0; q < 16; ++q) D1[(256 + row) * 48 + c + q] = __uint_as_float(v[q]); }
    for (int c = 0; c < 160; c += 16) { read16(256 + c, v); for (int q = 0; q < 16; ++q) D2[row * 160 + c + q] = __uint_as_float(v[q]); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
    std::vector<float> W(100 * 160), X(40 * 160), G(40 * 128);
    srand(1);
    for (auto& v : W) v = (float)(rand() % 15 - 7);
    for (auto& v : X) v = (float)(rand() % 15 - 7);
    for (auto& v : G) v = (float)(rand() % 15 - 7);
    for (int b = 0; b < 40; ++b) for (int n = 100; n < 128; ++n) G[b * 128 + n] = 0.f;
    float *dW, *dX, *dG, *d0, *d1, *d2;
    CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dG, G.size() * 4));
    CK(cudaMalloc(&d0, 128 * 48 * 4)); CK(cudaMalloc(&d1, 384 * 48 * 4)); CK(cudaMalloc(&d2, 128 * 160 * 4));
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM + 1024));
    probe<<<1, 128, SMEM + 1024>>>(dW, dX, dG, d0, d1, d2);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> h0(128 * 48), h1(384 * 48), h2(128 * 160);
    CK(cudaMemcpy(h0.data(), d0, h0.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h1.data(), d1, h1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h2.data(), d2, h2.size() * 4, cudaMemcpyDeviceToHost));
    int bad0 = 0, bad1 = 0, bad2 = 0;
    for (int n = 0; n < 100; ++n) for (int b = 0; b < 40; ++b) {
        float s = 0; for (int k = 0; k < 144; ++k) s += W[n * 160 + k] * X[b * 160 + k];
        if (s != h0[n * 48 + b]) { if (bad0 < 4) printf("test0 D[%d][%d] = %g want %g\n", n, b, h0[n * 48 + b], s); ++bad0; }
    }
    for (int k = 0; k < 160; ++k) for (int b = 0; b < 40; ++b) {
        float s = 0; for (int n = 0; n < 100; ++n) s += W[n * 160 + k] * G[b * 128 + n];
        if (s != h1[k * 48 + b]) { if (bad1 < 4) printf("test1 D[%d][%d] = %g want %g\n", k, b, h1[k * 48 + b], s); ++bad1; }
    }
    int bad3 = 0;
    for (int k = 0; k < 128; ++k) for (int b = 0; b < 40; ++b) {
        float s = 0; for (int n = 0; n < 100; ++n) s += W[n * 160 + k] * G[b * 128 + n];
        if (s != h1[(256 + k) * 48 + b]) ++bad3;
    }
    printf("diagnostic test3 (test 1 with LBO/SBO not swapped): %d bad\n", bad3);
    for (int n = 0; n < 100; ++n) for (int k = 0; k < 160; ++k) {
        float s = 0; for (int b = 0; b < 40; ++b) s += G[b * 128 + n] * X[b * 160 + k];
        if (s != h2[n * 160 + k]) { if (bad2 < 4) printf("test2 D[%d][%d] = %g want %g\n", n, k, h2[n * 160 + k], s); ++bad2; }
    }
    printf("MMA_PROBE test0 (K-major x K-major) %s (%d bad), test1 (MN-major A) %s (%d bad), test2 (MN-major A and B) %s (%d bad)\n",
           bad0 ? "FAIL" : "ok", bad0, bad1 ? "FAIL" : "ok", bad1, bad2 ? "FAIL" : "ok", bad2);
    return (bad0 || bad1 || bad2) ? 1 : 0;
}

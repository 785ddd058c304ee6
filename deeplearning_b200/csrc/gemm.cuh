// gemm.cuh — generic shared-memory tiled contraction  C[i][j] = sum_r A(i, r) * B(j, r)  in the reference's Real
// precision (fp32 / fp64 FMA), with functor operand loaders and functor epilogues.
//
// Every contraction of the training path is expressed through this one kernel:
// the operand functors implement dense row access, the layer-0 sample gather, and the implicit im2col /
// transposed-im2col index maps, so no receptive-field matrix is ever materialised in HBM (the reference's CPU
// path builds one per sample, Tensor.cpp:40-72).  A reduction that is long but has a small output (filter and
// weight gradients) is split over blockIdx.z into a partial-sum workspace; the finalize kernel adds the partials
// in a fixed order (deterministic, unlike the reference's atomicAdd kernel CudaTensor.cu:453-480) and runs the
// epilogue, which is where the gradient scale / SGD update is fused.
#pragma once

#include "common.cuh"

namespace dlb {

// Operand loader concept:
//   struct L { static constexpr bool kRedFast;  __device__ R operator()(int outer, int red) const; };
// kRedFast says which index is contiguous in memory, i.e. which one consecutive threads should walk when the
// tile is staged (coalescing).  The kernel guarantees outer < extent and red < K when it calls the functor.
//
// Epilogue concept:   struct E { __device__ void operator()(int i, int j, R acc) const; };

template <class R, int BM, int BN, int BK, int TM, int TN, class AL, class BL, class EP>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_kernel(int M, int N, int K, int k_per_split, AL a, BL b, EP ep, R* __restrict__ partials) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int PAD = 4;
    __shared__ R As[BK][BM + PAD];
    __shared__ R Bs[BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int kb = blockIdx.z * k_per_split;
    const int ke = min(K, kb + k_per_split);

    R acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = R(0);

    for (int k0 = kb; k0 < ke; k0 += BK) {
        for (int e = tid; e < BM * BK; e += NT) {
            int oo, rr;
            if (AL::kRedFast) { rr = e % BK; oo = e / BK; } else { oo = e % BM; rr = e / BM; }
            const int i = i0 + oo, r = k0 + rr;
            As[rr][oo] = (i < M && r < ke) ? a(i, r) : R(0);
        }
        for (int e = tid; e < BN * BK; e += NT) {
            int oo, rr;
            if (BL::kRedFast) { rr = e % BK; oo = e / BK; } else { oo = e % BN; rr = e / BN; }
            const int j = j0 + oo, r = k0 + rr;
            Bs[rr][oo] = (j < N && r < ke) ? b(j, r) : R(0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            R ar[TM], br[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) ar[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) br[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] += ar[i] * br[j];
        }
        __syncthreads();
    }

    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gi = i0 + ty * TM + i;
        if (gi >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gj = j0 + tx * TN + j;
            if (gj >= N) continue;
            if (split) partials[((size_t)blockIdx.z * M + gi) * N + gj] = acc[i][j];
            else ep(gi, gj, acc[i][j]);
        }
    }
}

// Adds the split partials in split order and runs the epilogue once per output element.
template <class R, class EP>
__global__ void gemm_finalize_kernel(const R* __restrict__ partials, int splits, int M, int N, EP ep) {
    const size_t total = (size_t)M * N;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    // fixed association (four interleaved running sums, then a fixed combine): deterministic, four loads in flight
    R s0 = R(0), s1 = R(0), s2 = R(0), s3 = R(0);
    int z = 0;
    for (; z + 3 < splits; z += 4) {
        s0 += partials[(size_t)z * total + e];
        s1 += partials[(size_t)(z + 1) * total + e];
        s2 += partials[(size_t)(z + 2) * total + e];
        s3 += partials[(size_t)(z + 3) * total + e];
    }
    for (; z < splits; ++z) s0 += partials[(size_t)z * total + e];
    ep((int)(e / N), (int)(e % N), (s0 + s1) + (s2 + s3));
}

struct GemmWorkspace {
    void* ptr = nullptr;      // device scratch for split partials
    size_t bytes = 0;
};

// Host launcher: picks the tile (64x64 / 32x32) and the number of reduction splits so that even the skinny
// contractions of small mini-batches put at least ~2 CTAs on each of the 148 SMs.
template <class R, class AL, class BL, class EP>
void gemm(cudaStream_t st, int M, int N, int K, AL a, BL b, EP ep, const GemmWorkspace& ws) {
    if (M <= 0 || N <= 0) return;
    constexpr int BK = 16;
    const long tiles64 = (long)ceil_div(M, 64) * ceil_div(N, 64);
    const bool big = tiles64 >= 2 * kNumSMs;
    const int BM = big ? 64 : 32, BN = BM;
    const long tiles = (long)ceil_div(M, BM) * ceil_div(N, BN);
    int splits = 1;
    if (tiles < 2 * kNumSMs && K > 4 * BK) {
        splits = (int)ceil_div<long>(2 * kNumSMs, tiles);
        const int max_by_k = K / (2 * BK) > 0 ? K / (2 * BK) : 1;   // at least two k-tiles per split
        if (splits > max_by_k) splits = max_by_k;
        const size_t per_split = (size_t)M * N * sizeof(R);
        const size_t max_by_ws = per_split ? ws.bytes / per_split : 0;
        if ((size_t)splits > max_by_ws) splits = (int)max_by_ws;
        if (splits < 1) splits = 1;
    }
    int k_per_split = ceil_div(ceil_div(K > 0 ? K : 1, splits), BK) * BK;
    splits = ceil_div(K > 0 ? K : 1, k_per_split);
    const dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits);
    R* partials = static_cast<R*>(ws.ptr);
    if (big) {
        DLB_LAUNCH((gemm_kernel<R, 64, 64, BK, 4, 4, AL, BL, EP>), grid, dim3(256), st, M, N, K, k_per_split, a, b, ep, partials);
    } else {
        DLB_LAUNCH((gemm_kernel<R, 32, 32, BK, 2, 2, AL, BL, EP>), grid, dim3(256), st, M, N, K, k_per_split, a, b, ep, partials);
    }
    if (splits > 1) {
        const size_t total = (size_t)M * N;
        DLB_LAUNCH((gemm_finalize_kernel<R, EP>), dim3((unsigned)ceil_div<size_t>(total, 256)), dim3(256), st,
                   (const R*)partials, splits, M, N, ep);
    }
}

}  // namespace dlb

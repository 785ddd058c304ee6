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

#include <type_traits>

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

// 64 x 64 output tile, 256 threads, 4 x 4 outputs per thread, BK reduction steps per stage.  The operand functors of the
// next stage are evaluated into registers while the current stage is multiplied out of shared memory (two stages, one
// barrier per stage).  Real = float runs the micro-kernel on the packed fp32 pipe: the 4 consecutive columns of a thread
// are two register pairs fetched by one 16-byte load, the row operand is the broadcast scalar of FFMA2.
template <class R> struct Gemm64 { static constexpr int BK = sizeof(R) == 4 ? 32 : 16; };

template <class R, class AL, class BL, class EP>
__global__ void __launch_bounds__(256)
gemm64_kernel(int M, int N, int K, int k_per_split, AL a, BL b, EP ep, R* __restrict__ partials) {
    constexpr int BM = 64, BN = 64, BK = Gemm64<R>::BK, NT = 256, PAD = 4;
    constexpr int EPT = BM * BK / NT;               // operand elements per thread and stage (8 / 4)
    __shared__ __align__(16) R As[2][BK][BM + PAD];
    __shared__ __align__(16) R Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int kb = blockIdx.z * k_per_split;
    const int ke = min(K, kb + k_per_split);

    R ra[EPT], rb[EPT];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            const int e = tid + q * NT;
            int oo, rr;
            if (AL::kRedFast) { rr = e % BK; oo = e / BK; } else { oo = e % BM; rr = e / BM; }
            ra[q] = (i0 + oo < M && k0 + rr < ke) ? a(i0 + oo, k0 + rr) : R(0);
            if (BL::kRedFast) { rr = e % BK; oo = e / BK; } else { oo = e % BN; rr = e / BN; }
            rb[q] = (j0 + oo < N && k0 + rr < ke) ? b(j0 + oo, k0 + rr) : R(0);
        }
    };
    auto store = [&](int s) {
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            const int e = tid + q * NT;
            int oo, rr;
            if (AL::kRedFast) { rr = e % BK; oo = e / BK; } else { oo = e % BM; rr = e / BM; }
            As[s][rr][oo] = ra[q];
            if (BL::kRedFast) { rr = e % BK; oo = e / BK; } else { oo = e % BN; rr = e / BN; }
            Bs[s][rr][oo] = rb[q];
        }
    };

    constexpr bool kPacked = sizeof(R) == 4;
    f32x2 acc2[4][2];
    R acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc2[i][0] = pack2(0.f, 0.f); acc2[i][1] = pack2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = R(0);
    }

    fetch(kb);
    store(0);
    __syncthreads();
    int s = 0;
    for (int k0 = kb; k0 < ke; k0 += BK, s ^= 1) {
        const bool more = k0 + BK < ke;
        if (more) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            if constexpr (kPacked) {
                const float4 av = *reinterpret_cast<const float4*>(&As[s][kk][ty * 4]);
                const f32x2_quad bq = *reinterpret_cast<const f32x2_quad*>(&Bs[s][kk][tx * 4]);
                const float ar[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc2[i][0] = ffma2_bcast(bq.a, ar[i], acc2[i][0]);
                    acc2[i][1] = ffma2_bcast(bq.b, ar[i], acc2[i][1]);
                }
            } else {
                R ar[4], br[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) ar[i] = As[s][kk][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) br[j] = Bs[s][kk][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] += ar[i] * br[j];
            }
        }
        if (more) store(s ^ 1);
        __syncthreads();
    }
    if constexpr (kPacked) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[i][0] = (R)lo_of(acc2[i][0]); acc[i][1] = (R)hi_of(acc2[i][0]);
            acc[i][2] = (R)lo_of(acc2[i][1]); acc[i][3] = (R)hi_of(acc2[i][1]);
        }
    }

    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = i0 + ty * 4 + i;
        if (gi >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = j0 + tx * 4 + j;
            if (gj >= N) continue;
            if (split) partials[((size_t)blockIdx.z * M + gi) * N + gj] = acc[i][j];
            else ep(gi, gj, acc[i][j]);
        }
    }
}

// Optional "strided" protocol of an operand loader: strided() (host-evaluable) promises that for outer < outer_limit() the value is
// origin()[outer * outer_step() + red * red_step()] and 1 from outer_limit() on (the bias column of ones).  The double kernel below then
// keeps ONE running pointer per operand and two bits of state per element instead of evaluating the functor per element and k-tile:
// the functor path spends ~30 integer instructions per fetched element (64-bit index products, the gather test and its branch), a
// third of everything the 128 x 64 kernel issued (ncu source page, profiles/r2_full_gemm_f64_B8192.txt).
template <class L, class = void> struct IsStrided { static constexpr bool value = false; };
template <class L> struct IsStrided<L, std::void_t<decltype(&L::red_step)>> { static constexpr bool value = true; };

// 8-byte asynchronous global -> shared copy (LDGSTS), zero fill when !valid (src-size 0: nothing is read)
__device__ __forceinline__ void gemm_cp_async_f64(double* dst_smem, const double* src, bool valid) {
#ifdef DLB_CUDA_EMU
    *dst_smem = valid ? *src : 0.0;
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int n = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
#endif
}
__device__ __forceinline__ void gemm_cp_async_commit() {
#ifndef DLB_CUDA_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N> __device__ __forceinline__ void gemm_cp_async_wait() {
#ifndef DLB_CUDA_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

template <class L, int EXT, int PAD, int BK, int NT>
struct StridedFetch {
    static constexpr int E = EXT * BK / NT;          // elements per thread and k-tile
    const double* p; const double* safe; size_t qstep, kstep; unsigned state; int r0, rq, soff, sstep;
    __device__ __forceinline__ void init(const L& l, int tid, int o_base, int extent, int kb) {
        int oo, rr, dq_o, dq_r;
        if (L::kRedFast) { rr = tid % BK; oo = tid / BK; dq_o = NT / BK; dq_r = 0; }
        else { oo = tid % EXT; rr = tid / EXT; dq_o = 0; dq_r = NT / EXT; }
        safe = l.origin();
        p = safe + (size_t)(o_base + oo) * l.outer_step() + (size_t)(kb + rr) * l.red_step();
        qstep = (size_t)dq_o * l.outer_step() + (size_t)dq_r * l.red_step();
        kstep = (size_t)BK * l.red_step();
        r0 = rr; rq = dq_r; state = 0;
        soff = rr * (EXT + PAD) + oo; sstep = dq_r * (EXT + PAD) + dq_o;   // element q lands at stage[soff + q * sstep]
        const int limit = l.outer_limit();
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int o = o_base + oo + q * dq_o;
            state |= (o >= extent ? 0u : (o >= limit ? 2u : 1u)) << (2 * q);   // 0: outside the matrix, 1: load, 2: the column of ones
        }
    }
    // k-tiles are issued in order, so the pointer just advances; `stage` = first element of the [BK][EXT + PAD] stage
    __device__ __forceinline__ void issue(double* stage, int k0, int ke) {
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const unsigned st = (state >> (2 * q)) & 3u;
            const bool in_k = k0 + r0 + q * rq < ke;
            double* dst = stage + soff + q * sstep;
            if (st == 2u) *dst = in_k ? 1.0 : 0.0;
            else { const bool ok = st == 1u && in_k; gemm_cp_async_f64(dst, ok ? p + q * qstep : safe, ok); }
        }
        p += kstep;
    }
};

// Real = double: 128 x 64 output tile, 256 threads, 8 x 4 outputs per thread, 8 reduction steps per stage.  The 64-wide micro-kernel
// above is bound by shared-memory wavefronts in double precision (4 LDS.128 per 16 DFMA, the column operand at a 32-byte lane stride =
// half-used wavefronts: ~47 % of the DFMA peak in steady state, profiles/r2_bench_f64.json); here a thread owns 8 rows (one broadcast
// 64-byte run) and two column PAIRS 32 columns apart, so that the 16 lanes of a row read 256 contiguous bytes per load: 6 LDS.128 per
// 32 DFMA and at most 12 wavefronts per 64 pipe cycles - the DFMA pipe (36.3 TFLOP/s measured, the same as DMMA) is the bound.
template <bool FAST, class AL, class BL, class EP>
__global__ void __launch_bounds__(256, 2)
gemm_f64_kernel(int M, int N, int K, int k_per_split, AL a, BL b, EP ep, double* __restrict__ partials) {
    constexpr int BM = 128, BN = 64, BK = 8, NT = 256, PAD = 4;
    constexpr int EA = BM * BK / NT, EB = BN * BK / NT;     // operand elements per thread and stage (4 / 2)
    constexpr int NS = FAST ? 3 : 2;                          // stages: cp.async two k-tiles ahead / register-staged double buffer
    __shared__ __align__(16) double As[NS][BK][BM + PAD];
    __shared__ __align__(16) double Bs[NS][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int kb = blockIdx.z * k_per_split;
    const int ke = min(K, kb + k_per_split);

    double ra[EA], rb[EB];
    auto fetch = [&](int k0) {
        {
#pragma unroll
            for (int q = 0; q < EA; ++q) {
                const int e = tid + q * NT;
                int oo, rr;
                if (AL::kRedFast) { rr = e % BK; oo = e / BK; } else { oo = e % BM; rr = e / BM; }
                ra[q] = (i0 + oo < M && k0 + rr < ke) ? a(i0 + oo, k0 + rr) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < EB; ++q) {
                const int e = tid + q * NT;
                int oo, rr;
                if (BL::kRedFast) { rr = e % BK; oo = e / BK; } else { oo = e % BN; rr = e / BN; }
                rb[q] = (j0 + oo < N && k0 + rr < ke) ? b(j0 + oo, k0 + rr) : 0.0;
            }
        }
    };
    auto store = [&](int s) {
#pragma unroll
        for (int q = 0; q < EA; ++q) {
            const int e = tid + q * NT;
            int oo, rr;
            if (AL::kRedFast) { rr = e % BK; oo = e / BK; } else { oo = e % BM; rr = e / BM; }
            As[s][rr][oo] = ra[q];
        }
#pragma unroll
        for (int q = 0; q < EB; ++q) {
            const int e = tid + q * NT;
            int oo, rr;
            if (BL::kRedFast) { rr = e % BK; oo = e / BK; } else { oo = e % BN; rr = e / BN; }
            Bs[s][rr][oo] = rb[q];
        }
    };

    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    auto multiply = [&](int s) {
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            double ar[8], br[4];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const double2 v = *reinterpret_cast<const double2*>(&As[s][kk][ty * 8 + i]);
                ar[i] = v.x; ar[i + 1] = v.y;
            }
            const double2 b0 = *reinterpret_cast<const double2*>(&Bs[s][kk][tx * 2]);
            const double2 b1 = *reinterpret_cast<const double2*>(&Bs[s][kk][32 + tx * 2]);
            br[0] = b0.x; br[1] = b0.y; br[2] = b1.x; br[3] = b1.y;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += ar[i] * br[j];
        }
    };
    if constexpr (FAST) {
        // three stages, asynchronous copies two k-tiles ahead, ONE barrier per k-tile: the stage refilled in iteration `it` was
        // multiplied in iteration it - 1, which every thread has left when it passes this iteration's barrier
        StridedFetch<AL, BM, PAD, BK, NT> fa; StridedFetch<BL, BN, PAD, BK, NT> fb;
        fa.init(a, tid, i0, M, kb); fb.init(b, tid, j0, N, kb);
        int k_next = kb, s_next = 0;
        auto issue = [&]() {
            if (k_next < ke) { fa.issue(&As[s_next][0][0], k_next, ke); fb.issue(&Bs[s_next][0][0], k_next, ke); }
            gemm_cp_async_commit();                       // (an empty group at the tail keeps the wait count uniform)
            k_next += BK; s_next = s_next == 2 ? 0 : s_next + 1;
        };
        issue(); issue();
        int s = 0;
        for (int k0 = kb; k0 < ke; k0 += BK, s = s == 2 ? 0 : s + 1) {
            gemm_cp_async_wait<1>();
            __syncthreads();
            issue();
            multiply(s);
        }
        gemm_cp_async_wait<0>();
    } else {
        fetch(kb);
        store(0);
        __syncthreads();
        int s = 0;
        for (int k0 = kb; k0 < ke; k0 += BK, s ^= 1) {
            const bool more = k0 + BK < ke;
            if (more) fetch(k0 + BK);
            multiply(s);
            if (more) store(s ^ 1);
            __syncthreads();
        }
    }

    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gi = i0 + ty * 8 + i;
        if (gi >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = j0 + (j >> 1) * 32 + tx * 2 + (j & 1);
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
    // fixed association (eight interleaved running sums, then a fixed combine): deterministic, eight loads in flight - each
    // load is a separate L2 round trip (the partials of one element lie M * N elements apart), and at mini-batch 256 this
    // kernel sits on the step's critical path
    R s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = R(0);
    for (int z = 0; z < splits; z += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (z + j < splits) s[j] += partials[(size_t)(z + j) * total + e];
    }
    ep((int)(e / N), (int)(e % N), ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7])));
}

struct GemmWorkspace {
    void* ptr = nullptr;      // device scratch for split partials
    size_t bytes = 0;
};

// Host launcher: 64 x 64 tiles (32 x 32 for outputs smaller than one such tile) and as many reduction splits as it takes
// for even the skinny contractions of small mini-batches to put about two CTAs on each of the 148 SMs.
template <class R, class AL, class BL, class EP>
void gemm(cudaStream_t st, int M, int N, int K, AL a, BL b, EP ep, const GemmWorkspace& ws) {
    if (M <= 0 || N <= 0) return;
    const bool small = M <= 32 || N <= 32;
    if constexpr (sizeof(R) == 8) {
        if (!small) {
            // 128 x 64 tiles, two CTAs per SM.  The reduction is split so that the grid fills whole waves of 2 x 148 CTAs: the smallest
            // number of splits (at least two k-tiles each) whose last wave is at least 85 % full, else the fullest one.
            constexpr int BM = 128, BN = 64, BK = 8;
            const long tiles = (long)ceil_div(M, BM) * ceil_div(N, BN);
            const long cap = 2L * kNumSMs;
            const int ktiles = ceil_div(K > 0 ? K : 1, BK);
            const size_t per_split = (size_t)M * N * sizeof(R);
            long max_s = std::min<long>(ktiles / 2 > 0 ? ktiles / 2 : 1, per_split ? (long)(ws.bytes / per_split) : 1);
            if (max_s > 64) max_s = 64;
            if (max_s < 1) max_s = 1;
            int splits = 1;
            if (tiles < 4 * cap) {
                double best = -1.0;
                for (long sp = 1; sp <= max_s; ++sp) {
                    const long ctas = tiles * sp;
                    const double eff = (double)ctas / (double)(ceil_div<long>(ctas, cap) * cap);
                    if (eff > best + 1e-9) { best = eff; splits = (int)sp; }
                    if (eff >= 0.85) { splits = (int)sp; break; }
                }
            }
            int k_per_split = ceil_div(ceil_div(K > 0 ? K : 1, splits), BK) * BK;
            splits = ceil_div(K > 0 ? K : 1, k_per_split);
            const dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits);
            R* partials = static_cast<R*>(ws.ptr);
            bool fast = false;
            if constexpr (IsStrided<AL>::value && IsStrided<BL>::value) fast = a.strided() && b.strided();
            if constexpr (IsStrided<AL>::value && IsStrided<BL>::value) {
                if (fast) DLB_LAUNCH((gemm_f64_kernel<true, AL, BL, EP>), grid, dim3(256), st, M, N, K, k_per_split, a, b, ep, partials);
            }
            if (!fast) DLB_LAUNCH((gemm_f64_kernel<false, AL, BL, EP>), grid, dim3(256), st, M, N, K, k_per_split, a, b, ep, partials);
            if (splits > 1) {
                const size_t total = (size_t)M * N;
                DLB_LAUNCH((gemm_finalize_kernel<R, EP>), dim3((unsigned)ceil_div<size_t>(total, 256)), dim3(256), st,
                           (const R*)partials, splits, M, N, ep);
            }
            return;
        }
    }
    const int BM = small ? 32 : 64, BN = BM;
    const int BK = small ? 16 : Gemm64<R>::BK;
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
    if (small) {
        DLB_LAUNCH((gemm_kernel<R, 32, 32, 16, 2, 2, AL, BL, EP>), grid, dim3(256), st, M, N, K, k_per_split, a, b, ep, partials);
    } else {
        DLB_LAUNCH((gemm64_kernel<R, AL, BL, EP>), grid, dim3(256), st, M, N, K, k_per_split, a, b, ep, partials);
    }
    if (splits > 1) {
        const size_t total = (size_t)M * N;
        DLB_LAUNCH((gemm_finalize_kernel<R, EP>), dim3((unsigned)ceil_div<size_t>(total, 256)), dim3(256), st,
                   (const R*)partials, splits, M, N, ep);
    }
}

}  // namespace dlb

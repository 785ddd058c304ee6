// fused.cuh — the training step's fast path ("train mode"): kernels that fuse what the reference runs as separate
// per-sample tensor operations, for nets of the shape  [CLayer -> PLayer(max|min)]  NLayer*  NLayer(head).
//
//   conv_pool_forward_kernel    CLayer::act + PLayer::act in one pass: convolution, untied bias, activation, window
//                               max/min and its index map; the un-pooled activations never leave the SM
//                               (CLayer.inl:108-137, PLayer.inl:115-138, Tensor.cpp:673-716).
//   conv_pool_backward_kernel   PLayer::backpropagate + CLayer::backpropagate for a first layer: routes every pooled
//                               gradient to its arg-max position, multiplies by the activation derivative (recomputed
//                               from the pooled value), accumulates the untied bias gradient and the filter gradient
//                               over the CTA's samples in registers (PLayer.inl:140-168, CLayer.inl:139-197,
//                               Tensor.cpp:466-498, 721-734).  Zero gradients are skipped exactly like the reference.
//   full_head_kernel            last NLayer forward + softmax/activation + cost derivative + activation gradient +
//                               the input gradient (W^T g), optionally already multiplied by the previous layer's
//                               activation derivative (NLayer.inl:131-171, CostFunction.inl:47-51).
//   batch_outer_partial_kernel  weight + bias gradient of a narrow NLayer as per-chunk partial sums.
//   reduce_update_kernel        sums the per-CTA / per-chunk partial gradients in a fixed order into the flat gradient
//                               buffer and applies the SGD update in the same pass (ALayer.inl:97-122) — the weight
//                               update fused into the gradient reduction.
//
// The trace mode (dlb_net_forward_backward_host + dlb_net_read, the op-level C ABI) keeps using the unfused operators
// of layers.cuh, which materialise every Trace buffer of the reference.
#pragma once

#include "layers.cuh"

namespace dlb {

// Activation derivative recomputed from the activation VALUE; the same expressions the forward pass stores as
// "Derivatives" (ActivationFunctionHelper.cpp:64-114), so the result is bit-identical to the stored trace.
template <class R>
__device__ __forceinline__ R deriv_from_value(int act, R v) {
    switch (act) {
    case DLB_ACT_SIGMOID: return v * (R(1) - v);
    case DLB_ACT_TANH: return R(1) - v * v;
    case DLB_ACT_RELU: return v > R(0) ? R(1) : R(0);
    default: return R(1);
    }
}

struct ConvPoolGeom {
    int C, H, W, F, OH, OW, ph, pw, POH, POW, Hp, Wp, K, NPOS;
    int bRS, bHO;   // backward kernel's shared-memory image: row stride and offset of the odd-column half (see there)
};

__device__ __forceinline__ const int* resolve_gather(const int* gather_base, const int* cursor) {
    return gather_base ? gather_base + (cursor ? *cursor : 0) : nullptr;
}

// ---------------------------------------------------------------------------------------------------------------
// Convolution (stride 1) + untied bias + activation + 2 x 2 max/min pooling.
// One CTA walks over groups of `spc` samples.  A thread owns PP horizontally adjacent pooled positions and FPT filters of
// one sample: it keeps the (KH+1) x (KW+2*PP-1) input patch in registers, reads every filter tap once (shared-memory
// broadcast) for its 4*PP convolution outputs, and reads the untied bias through the read-only path.  The activation and
// the pooling operator are compile-time in the epilogue (one uniform switch per sample, no per-element branches).
// smem: filters [F][K] | `spc` zero-padded input images [C][Hp][Wp]
// ---------------------------------------------------------------------------------------------------------------
template <class R, int ACT> __device__ __forceinline__ R activation_ct(R z) {
    if (ACT == DLB_ACT_SIGMOID) return R(1) / (R(1) + dl_exp<R>(-z));
    if (ACT == DLB_ACT_TANH) return dl_tanh<R>(z);
    if (ACT == DLB_ACT_RELU) return z > R(0) ? z : R(0);
    return z;
}

template <class R, int FPT, int PP, int ACT, bool IS_MAX>
__device__ __forceinline__ void conv_pool_epilogue(const R (&acc)[FPT][2][2 * PP], const R* __restrict__ bias, R* __restrict__ pooled,
                                                   int* __restrict__ map, const ConvPoolGeom& g, int fs, int py, int pxg, size_t b) {
#pragma unroll
    for (int fi = 0; fi < FPT; ++fi) {
        const int f = fs * FPT + fi;
        if (f >= g.F) continue;
#pragma unroll
        for (int pp = 0; pp < PP; ++pp) {
            const int px = pxg * PP + pp;
            if (px >= g.POW) continue;
            R best = IS_MAX ? -real_limits<R>::max() : real_limits<R>::max();
            int best_id = -1;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int id = (f * g.OH + py * 2 + dy) * g.OW + px * 2 + dx;
                    const R v = activation_ct<R, ACT>(acc[fi][dy][pp * 2 + dx] + __ldg(bias + id));
                    if (IS_MAX ? (best < v) : (best > v)) { best = v; best_id = id; }   // strict: the first wins ties
                }
            const size_t o = (b * g.F + f) * g.NPOS + py * g.POW + px;
            pooled[o] = best;
            map[o] = best_id;
        }
    }
}

template <class R, int KH, int KW, int FPT, int PP, int TMAX>
__global__ void __launch_bounds__(TMAX, 1) conv_pool_forward_kernel(const R* __restrict__ x, const int* __restrict__ gather_base,
                                                                    const int* __restrict__ cursor, const R* __restrict__ bias,
                                                                    const R* __restrict__ filt, R* __restrict__ pooled, int* __restrict__ map,
                                                                    int B, ConvPoolGeom g, int act, int is_max, int spc, int tps) {
    DLB_DYNAMIC_SMEM(R, smem);
    R* filt_s = smem;
    const int img = g.C * g.Hp * g.Wp;
    R* x_all = filt_s + g.F * g.K;
    const int* gather = resolve_gather(gather_base, cursor);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < g.F * g.K; i += nt) filt_s[i] = filt[i];
    for (int i = tid; i < spc * img; i += nt) x_all[i] = R(0);   // the borders stay zero (padding)
    __syncthreads();

    const int slot = tid / tps, lt = tid - slot * tps;
    const int PGW = (g.POW + PP - 1) / PP, NPG = g.POH * PGW;
    const int fs = lt / NPG, pg = lt - fs * NPG;
    const int py = pg / PGW, pxg = pg - py * PGW;
    const bool active = slot < spc && fs * FPT < g.F;
    const R* x_s = x_all + (slot < spc ? slot : 0) * img;
    const int HW = g.H * g.W, n_in = g.C * HW;
    constexpr int PR = KH + 1, PC = KW + 2 * PP - 1;
    // (16-byte vector loads of the patch rows were tried: conflict-free, but the kernel then spills at the 96-register cap
    // of 576 threads and is slower - 325 vs 284 us at B=8192 - so the scalar rows stay.)

    // The next group's inputs are fetched into registers while the current group is computed.
    constexpr int XPT = 4;
    const int n_all = spc * n_in;
    const bool prefetch = n_all <= XPT * nt;
    R xr[XPT];
    auto fetch = [&](int b0) {
#pragma unroll
        for (int j = 0; j < XPT; ++j) {
            const int i = tid + j * nt;
            const int sm = i / n_in, e = i - sm * n_in;
            const int b = b0 + sm;
            xr[j] = (i < n_all && b < B) ? __ldg(x + (size_t)(gather ? gather[b] : b) * n_in + e) : R(0);
        }
    };
    auto stage_one = [&](int i, R v) {
        const int sm = i / n_in, e = i - sm * n_in;
        const int c = e / HW, r = e - c * HW;
        const int yy = r / g.W, xx = r - yy * g.W;
        x_all[sm * img + (c * g.Hp + yy + g.ph) * g.Wp + xx + g.pw] = v;
    };
    const int stride = gridDim.x * spc;
    if (prefetch && (int)blockIdx.x * spc < B) fetch(blockIdx.x * spc);
    for (int b0 = blockIdx.x * spc; b0 < B; b0 += stride) {
        if (prefetch) {
#pragma unroll
            for (int j = 0; j < XPT; ++j) { const int i = tid + j * nt; if (i < n_all) stage_one(i, xr[j]); }
            if (b0 + stride < B) fetch(b0 + stride);
        } else {
            for (int i = tid; i < n_all; i += nt) {
                const int b = b0 + i / n_in;
                stage_one(i, b < B ? x[(size_t)(gather ? gather[b] : b) * n_in + (i % n_in)] : R(0));
            }
        }
        __syncthreads();
        const int b = b0 + slot;
        if (active && b < B) {
            R acc[FPT][2][2 * PP];
#pragma unroll
            for (int fi = 0; fi < FPT; ++fi)
#pragma unroll
                for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                    for (int q = 0; q < 2 * PP; ++q) acc[fi][dy][q] = R(0);
            for (int c = 0; c < g.C; ++c) {
                R p[PR][PC];
#pragma unroll
                for (int r = 0; r < PR; ++r)
#pragma unroll
                    for (int q = 0; q < PC; ++q) {
                        const int col = pxg * (2 * PP) + q;
                        p[r][q] = col < g.Wp ? x_s[(c * g.Hp + py * 2 + r) * g.Wp + col] : R(0);
                    }
#pragma unroll
                for (int fi = 0; fi < FPT; ++fi) {
                    const int f = fs * FPT + fi;
                    const R* w = filt_s + (f < g.F ? f : 0) * g.K + c * (KH * KW);
#pragma unroll
                    for (int ky = 0; ky < KH; ++ky)
#pragma unroll
                        for (int kx = 0; kx < KW; ++kx) {
                            const R wv = w[ky * KW + kx];
#pragma unroll
                            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                                for (int q = 0; q < 2 * PP; ++q) acc[fi][dy][q] += p[dy + ky][q + kx] * wv;
                        }
                }
            }
#define DLB_EPI(ACT)                                                                                               \
            do {                                                                                                       \
                if (is_max) conv_pool_epilogue<R, FPT, PP, ACT, true>(acc, bias, pooled, map, g, fs, py, pxg, (size_t)b);  \
                else conv_pool_epilogue<R, FPT, PP, ACT, false>(acc, bias, pooled, map, g, fs, py, pxg, (size_t)b);        \
            } while (0)
            switch (act) {
            case DLB_ACT_RELU: DLB_EPI(DLB_ACT_RELU); break;
            case DLB_ACT_SIGMOID: DLB_EPI(DLB_ACT_SIGMOID); break;
            case DLB_ACT_TANH: DLB_EPI(DLB_ACT_TANH); break;
            default: DLB_EPI(DLB_ACT_LINEAR); break;
            }
#undef DLB_EPI
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Real = float variant of the forward kernel on the packed fp32 pipe (FFMA2, `fma.rn.f32x2`, sm_100+): one instruction
// performs two fused multiply-adds on a 64-bit register pair, with the filter tap broadcast to both halves - half the
// issue slots of the scalar FFMA loop for the same IEEE fma results.  A thread owns 2 horizontally adjacent pooled
// positions (4 convolution columns q = 0..3) x 2 rows x 4 filters; accumulators are kept as pairs (q, q+2) and the input
// patch as pairs (column j, column j+2), so every pair an FFMA2 needs is an aligned register pair for every tap.
// Filters sit in shared memory tap-major [K][Fp] (Fp = filters padded to a multiple of 4): the 4 taps of a thread's
// filters are one 16-byte broadcast load.  One sample per CTA iteration; filter groups beyond `fgp` are looped.
// ---------------------------------------------------------------------------------------------------------------
// Pair arithmetic of the kernel below: float = one FFMA2 on a 64-bit register pair, double = two DFMA on a {lo, hi} struct
// (same data layout and control flow, so Real = double gets the staging / rolling-window structure as well).
template <class R> struct PairOf;
template <> struct PairOf<float> {
    using type = f32x2;
    static __device__ __forceinline__ type zero() { return pack2(0.f, 0.f); }
    static __device__ __forceinline__ type fma(type a, float w, type c) { return ffma2_bcast(a, w, c); }
    static __device__ __forceinline__ float lo(type v) { return lo_of(v); }
    static __device__ __forceinline__ float hi(type v) { return hi_of(v); }
};
struct alignas(16) f64x2 { double lo, hi; };
template <> struct PairOf<double> {
    using type = f64x2;
    static __device__ __forceinline__ type zero() { return f64x2{0.0, 0.0}; }
    static __device__ __forceinline__ type fma(type a, double w, type c) { return f64x2{::fma(a.lo, w, c.lo), ::fma(a.hi, w, c.hi)}; }
    static __device__ __forceinline__ double lo(type v) { return v.lo; }
    static __device__ __forceinline__ double hi(type v) { return v.hi; }
};
template <class R> struct alignas(16) PairQuad { typename PairOf<R>::type a, b; };   // two pair slots
template <class R> struct alignas(16) Vec4 { R v[4]; };                              // the 4 taps of a thread's filters

#ifdef DLB_CUDA_EMU
template <class T> __device__ __forceinline__ void stage_copy_elem(T* dst, const T* src) { *dst = *src; }
#else
template <class T> __device__ __forceinline__ void stage_copy_elem(T* dst, const T* src) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "cp.async element size");
    if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
#endif

constexpr int kX2Threads = 384;   // launch bound of the packed kernel; float: two CTAs per SM (80 registers per thread)
constexpr int kX2VecBias = 2, kX2VecOut = 4;

template <class R, int ACT, bool IS_MAX>
__device__ __forceinline__ void conv_pool_epilogue_x2(const typename PairOf<R>::type (&acc)[4][2][2], const R* __restrict__ bias, R* __restrict__ pooled,
                                                      int* __restrict__ map, const ConvPoolGeom& g, int fs, int py, int pxg, size_t b, int flags) {
    using P = PairOf<R>;
    const int col0 = pxg * 4;
    R bzs[4][2][4];   // all untied biases first (one batch of loads in flight), then the arithmetic
#pragma unroll
    for (int fi = 0; fi < 4; ++fi) {
        const int f = min(fs * 4 + fi, g.F - 1);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const R* src = bias + (f * g.OH + py * 2 + dy) * g.OW + col0;
            bool vec = false;
            if constexpr (sizeof(R) == 4) {
                if (flags & kX2VecBias) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(src));
                    bzs[fi][dy][0] = t.x; bzs[fi][dy][1] = t.y; bzs[fi][dy][2] = t.z; bzs[fi][dy][3] = t.w;
                    vec = true;
                }
            }
            if (!vec) {
#pragma unroll
                for (int q = 0; q < 4; ++q) bzs[fi][dy][q] = col0 + q < g.OW ? __ldg(src + q) : R(0);
            }
        }
    }
#pragma unroll
    for (int fi = 0; fi < 4; ++fi) {
        const int f = fs * 4 + fi;
        if (f >= g.F) continue;
        R v[2][4];
        const int id0 = (f * g.OH + py * 2) * g.OW + col0;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const R* bz = bzs[fi][dy];
            v[dy][0] = activation_ct<R, ACT>(P::lo(acc[fi][dy][0]) + bz[0]);
            v[dy][1] = activation_ct<R, ACT>(P::lo(acc[fi][dy][1]) + bz[1]);
            v[dy][2] = activation_ct<R, ACT>(P::hi(acc[fi][dy][0]) + bz[2]);
            v[dy][3] = activation_ct<R, ACT>(P::hi(acc[fi][dy][1]) + bz[3]);
        }
        R best[2]; int best_id[2];
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
            best[pp] = IS_MAX ? -real_limits<R>::max() : real_limits<R>::max();
            best_id[pp] = -1;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const R c = v[dy][pp * 2 + dx];
                    if (IS_MAX ? (best[pp] < c) : (best[pp] > c)) { best[pp] = c; best_id[pp] = id0 + dy * g.OW + pp * 2 + dx; }   // strict: the first wins ties
                }
        }
        const size_t o = (b * g.F + f) * g.NPOS + py * g.POW + pxg * 2;
        bool stored = false;
        if constexpr (sizeof(R) == 4) {
            if (flags & kX2VecOut) {
                *reinterpret_cast<float2*>(pooled + o) = make_float2(best[0], best[1]);
                *reinterpret_cast<int2*>(map + o) = make_int2(best_id[0], best_id[1]);
                stored = true;
            }
        }
        if (!stored) {
            pooled[o] = best[0]; map[o] = best_id[0];
            if (pxg * 2 + 1 < g.POW) { pooled[o + 1] = best[1]; map[o + 1] = best_id[1]; }
        }
    }
}

// Asynchronous global -> shared copies for the image staging (the emulator build copies synchronously).
#ifdef DLB_CUDA_EMU
__device__ __forceinline__ void stage_copy16(float* dst, const float* src) { for (int q = 0; q < 4; ++q) dst[q] = src[q]; }
__device__ __forceinline__ void stage_copy4(float* dst, const float* src) { *dst = *src; }
__device__ __forceinline__ void stage_commit() {}
template <int N> __device__ __forceinline__ void stage_wait() {}
#else
__device__ __forceinline__ void stage_copy16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void stage_copy4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// Shared-memory image of the packed kernel: every element slot holds the PAIR (x[col], x[col + 2]) of the zero-padded row
// (8 bytes, row stride RS slots), so the pairs an FFMA2 consumes arrive as aligned 64-bit registers straight from 16-byte
// loads - ptxas otherwise keeps the patch as single registers and re-packs a pair with two MOVs in front of every FFMA2
// (616 moves per 400 FFMA2 measured).  The image is double buffered and filled by cp.async one sample ahead (the gather
// index two ahead): a sample costs ONE block barrier and the compute phase never waits on global memory.
template <class R, int KH, int KW>
__global__ void __launch_bounds__(kX2Threads, sizeof(R) == 4 ? 2 : 1) conv_pool_forward_x2_kernel(const R* __restrict__ x, const int* __restrict__ gather_base,
                                                                               const int* __restrict__ cursor, const R* __restrict__ bias,
                                                                               const R* __restrict__ filt, R* __restrict__ pooled,
                                                                               int* __restrict__ map, int B, ConvPoolGeom g, int act, int is_max,
                                                                               int nfg, int fgp, int RS, int flags) {
    using P = PairOf<R>;
    using pair_t = typename P::type;
    DLB_DYNAMIC_SMEM(R, smem);
    pdl_trigger();   // the NLayer contraction that follows may set itself up (barriers, tensor memory) while this grid runs
    const int Fp = nfg * 4;
    const int img_stride = g.C * g.Hp * RS * 2;   // floats per image buffer (a multiple of 4: RS is even)
    R* filt_s = smem;                 // [K][Fp], zero padded filters
    R* x_buf = smem + g.K * Fp;       // two pair images [C][Hp][RS][2]
    const int* gather = resolve_gather(gather_base, cursor);
    const int tid = threadIdx.x, nt = blockDim.x;
    // the gather indices of the first two samples are requested before anything else: their latency overlaps the set-up
    auto row_of = [&](int b) -> int { return b < B ? (gather ? __ldg(gather + b) : b) : -1; };   // int on purpose: a widening here would wait for the load
    const int G = gridDim.x;
    const int row0 = row_of(blockIdx.x);
    int row_next = row_of((int)blockIdx.x + G);        // sample it + 1
    for (int i = tid; i < g.K * Fp; i += nt) {
        const int k = i / Fp, f = i - k * Fp;
        filt_s[i] = f < g.F ? filt[f * g.K + k] : R(0);
    }
    for (int i = tid; i < 2 * img_stride; i += nt) x_buf[i] = R(0);   // borders (padding) and unused slots stay zero
    __syncthreads();

    const int PGW = (g.POW + 1) / 2, NPG = g.POH * PGW;
    const int fsl = tid / NPG, pg = tid - fsl * NPG;
    const int py = pg / PGW, pxg = pg - py * PGW;
    const bool active = fsl < fgp;
    const int n_in = g.C * g.H * g.W;
    constexpr int NP = KW + 1;            // pairs per patch row (even)

    // staging: a warp copies whole image rows, lane = column: element (row, col) is the low half of slot col and the high
    // half of slot col - 2
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    auto stage = [&](int row, R* dst) {
        const R* xs = x + (size_t)row * n_in;
        for (int r = warp; r < g.C * g.H; r += nw) {
            const int c = g.C == 1 ? 0 : r / g.H;
            const int yy = r - c * g.H;
            R* drow = dst + (size_t)((c * g.Hp + yy + g.ph) * RS) * 2;
            for (int xx = lane; xx < g.W; xx += 32) {
                const int cc = xx + g.pw;
                const R* src = xs + r * g.W + xx;
                stage_copy_elem(drow + cc * 2, src);
                if (cc >= 2) stage_copy_elem(drow + (cc - 2) * 2 + 1, src);
            }
        }
    };
    if (row0 >= 0) stage(row0, x_buf);
    stage_commit();
    int it = 0;
    for (int b = blockIdx.x; b < B; b += G, ++it) {
        stage_wait<0>();        // this thread's copies of sample `it` have landed
        __syncthreads();        // ... everyone's have, and everyone is done computing on the other buffer
        const R* x_s = x_buf + (it & 1) * img_stride;
        if (row_next >= 0) stage(row_next, x_buf + ((it + 1) & 1) * img_stride);
        stage_commit();
        row_next = row_of(b + 2 * G);
        if (active) {
            for (int fs = fsl; fs < nfg; fs += fgp) {
                pair_t acc[4][2][2];
#pragma unroll
                for (int fi = 0; fi < 4; ++fi)
#pragma unroll
                    for (int dy = 0; dy < 2; ++dy) { acc[fi][dy][0] = P::zero(); acc[fi][dy][1] = P::zero(); }
                for (int c = 0; c < g.C; ++c) {
                    // rolling window over the patch rows: row r serves tap row ky = r (upper conv row) and ky = r - 1 (lower one),
                    // so only two rows of pairs are live at a time
                    auto load_row = [&](int r, pair_t (&pr)[NP]) {
                        const PairQuad<R>* q = reinterpret_cast<const PairQuad<R>*>(x_s + (size_t)((c * g.Hp + py * 2 + r) * RS + pxg * 4) * 2);
#pragma unroll
                        for (int j = 0; j < NP / 2; ++j) { const PairQuad<R> t = q[j]; pr[2 * j] = t.a; pr[2 * j + 1] = t.b; }
                    };
                    const R* wbase = filt_s + c * (KH * KW) * Fp + fs * 4;
                    pair_t rows[2][NP];
                    load_row(0, rows[0]);
#pragma unroll
                    for (int ky = 0; ky < KH; ++ky) {
                        load_row(ky + 1, rows[(ky + 1) & 1]);
#pragma unroll
                        for (int kx = 0; kx < KW; ++kx) {
                            const Vec4<R> w4 = *reinterpret_cast<const Vec4<R>*>(wbase + (ky * KW + kx) * Fp);
#pragma unroll
                            for (int fi = 0; fi < 4; ++fi)
#pragma unroll
                                for (int dy = 0; dy < 2; ++dy) {
                                    acc[fi][dy][0] = P::fma(rows[(ky + dy) & 1][kx], w4.v[fi], acc[fi][dy][0]);
                                    acc[fi][dy][1] = P::fma(rows[(ky + dy) & 1][kx + 1], w4.v[fi], acc[fi][dy][1]);
                                }
                        }
                    }
                }
#define DLB_EPI(ACT)                                                                                                     \
                do {                                                                                                      \
                    if (is_max) conv_pool_epilogue_x2<R, ACT, true>(acc, bias, pooled, map, g, fs, py, pxg, (size_t)b, flags);   \
                    else conv_pool_epilogue_x2<R, ACT, false>(acc, bias, pooled, map, g, fs, py, pxg, (size_t)b, flags);         \
                } while (0)
                switch (act) {
                case DLB_ACT_RELU: DLB_EPI(DLB_ACT_RELU); break;
                case DLB_ACT_SIGMOID: DLB_EPI(DLB_ACT_SIGMOID); break;
                case DLB_ACT_TANH: DLB_EPI(DLB_ACT_TANH); break;
                default: DLB_EPI(DLB_ACT_LINEAR); break;
                }
#undef DLB_EPI
            }
        }
    }
    stage_wait<0>();
}

// ---------------------------------------------------------------------------------------------------------------
// Backward of the fused first stage.  One warp per filter; the lanes split the filter's pooled windows.  Per lane:
// the K = C*KH*KW filter-gradient accumulators and the PH*PW bias-gradient accumulators of each of its windows, all in
// registers across the CTA's samples.  partial layout per CTA = the layer's flat parameter layout [bias | filters].
// ---------------------------------------------------------------------------------------------------------------
template <class R, int C, int KH, int KW, int PH, int PW, int WPL, int FMAX>
__global__ void __launch_bounds__(32 * FMAX, 1) conv_pool_backward_kernel(const R* __restrict__ x, const int* __restrict__ gather_base, const int* __restrict__ cursor,
                                          const R* __restrict__ dpool, const int* __restrict__ map, const R* __restrict__ pooled,
                                          R* __restrict__ partial, int B, ConvPoolGeom g, int act, int n_par) {
    DLB_DYNAMIC_SMEM(R, x_s);
    const int* gather = resolve_gather(gather_base, cursor);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int f = tid / 32, lane = tid % 32;
    const int HW = g.H * g.W;
    constexpr int K = C * KH * KW;
    // Shared-memory image layout: every padded row stores its even columns first, then (at +bHO) its odd columns, row stride
    // bRS.  The lanes of a warp walk consecutive pooled windows, i.e. conv positions two columns apart: in this layout they
    // read consecutive words (one bank each) instead of every other word, and bRS is chosen on the host so that the window
    // rows a warp spans fall into disjoint bank ranges.  (ncu before: 62 % of the shared-memory wavefronts were conflicts.)
    for (int i = tid; i < C * g.Hp * g.bRS; i += nt) x_s[i] = R(0);
    __syncthreads();
    auto stage_one = [&](int i, R v) {
        const int c = i / HW, r = i - c * HW;
        const int yy = r / g.W, xx = r - yy * g.W;
        const int col = xx + g.pw;
        x_s[(c * g.Hp + yy + g.ph) * g.bRS + (col & 1) * g.bHO + (col >> 1)] = v;
    };

    R dk[K];
    R db[WPL][PH * PW];
#pragma unroll
    for (int k = 0; k < K; ++k) dk[k] = R(0);
#pragma unroll
    for (int wi = 0; wi < WPL; ++wi)
#pragma unroll
        for (int q = 0; q < PH * PW; ++q) db[wi][q] = R(0);

    // Per-sample operands are fetched one sample ahead into registers: the lane's WPL windows (index map, pooled value,
    // pooled gradient) and the thread's share of the input image.
    constexpr int XPT = 2;
    const int n_in = C * HW;
    const bool prefetch = n_in <= XPT * nt;
    R xr[XPT];
    int wm[WPL]; R wp[WPL], wd[WPL];
    auto fetch = [&](int b) {
        const R* xs = x + (size_t)(gather ? gather[b] : b) * n_in;
        if (prefetch) {
#pragma unroll
            for (int j = 0; j < XPT; ++j) { const int i = tid + j * nt; xr[j] = i < n_in ? __ldg(xs + i) : R(0); }
        }
#pragma unroll
        for (int wi = 0; wi < WPL; ++wi) {
            const int win = lane + 32 * wi;
            wm[wi] = -1; wp[wi] = R(0); wd[wi] = R(0);
            if (win < g.NPOS) {
                const size_t o = ((size_t)b * g.F + f) * g.NPOS + win;
                wm[wi] = __ldg(map + o); wp[wi] = __ldg(pooled + o); wd[wi] = __ldg(dpool + o);
            }
        }
    };
    // (hoisting the per-window geometry out of the sample loop was tried: 15 more live registers, spills, 345 vs 310 us)
    if ((int)blockIdx.x < B) fetch(blockIdx.x);
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        if (prefetch) {
#pragma unroll
            for (int j = 0; j < XPT; ++j) { const int i = tid + j * nt; if (i < n_in) stage_one(i, xr[j]); }
        } else {
            const R* xs = x + (size_t)(gather ? gather[b] : b) * n_in;
            for (int i = tid; i < n_in; i += nt) stage_one(i, xs[i]);
        }
        int cm[WPL]; R cg[WPL];
#pragma unroll
        for (int wi = 0; wi < WPL; ++wi) { cm[wi] = wm[wi]; cg[wi] = deriv_from_value<R>(act, wp[wi]) * wd[wi]; }
        if (b + (int)gridDim.x < B) fetch(b + gridDim.x);
        __syncthreads();
#pragma unroll
        for (int wi = 0; wi < WPL; ++wi) {
            const int win = lane + 32 * wi;
            const int m = cm[wi];
            if (win >= g.NPOS || m < 0) continue;
            const R gval = cg[wi];
            const int rem = m - f * g.OH * g.OW;
            const int oy = rem / g.OW, ox = rem - oy * g.OW;
            const int py = win / g.POW, px = win - py * g.POW;
            const int sel = (oy - py * PH) * PW + (ox - px * PW);
#pragma unroll
            for (int q = 0; q < PH * PW; ++q) db[wi][q] += (q == sel) ? gval : R(0);
            if (gval != R(0)) {   // the reference skips zero factors as well (Tensor.cpp:484-486)
                // taps of one row alternate between the two column halves: even taps from pa, odd taps from pb
                const int par = ox & 1, oxh = ox >> 1;
                const R* pa = x_s + oy * g.bRS + par * g.bHO + oxh;
                const R* pb = x_s + oy * g.bRS + (par ? oxh + 1 : g.bHO + oxh);
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int ky = 0; ky < KH; ++ky) {
                        const int ro = (c * g.Hp + ky) * g.bRS;
#pragma unroll
                        for (int kx = 0; kx < KW; ++kx)
                            dk[(c * KH + ky) * KW + kx] += ((kx & 1) ? pb[ro + (kx >> 1)] : pa[ro + (kx >> 1)]) * gval;
                    }
            }
        }
        __syncthreads();
    }

    R* pb = partial + (size_t)blockIdx.x * n_par;
#pragma unroll
    for (int wi = 0; wi < WPL; ++wi) {
        const int win = lane + 32 * wi;
        if (win >= g.NPOS) continue;
        const int py = win / g.POW, px = win - py * g.POW;
#pragma unroll
        for (int dy = 0; dy < PH; ++dy)
#pragma unroll
            for (int dx = 0; dx < PW; ++dx) pb[(f * g.OH + py * PH + dy) * g.OW + px * PW + dx] = db[wi][dy * PW + dx];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const R v = warp_sum<R>(dk[k]);
        if (lane == 0) pb[g.F * g.OH * g.OW + f * K + k] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Staged variant of the backward kernel (same work split, same partial layout, same arithmetic order).  Everything a
// sample needs - the image AND the index map / pooled values / pooled gradients of all F * NPOS windows - is copied into a
// double buffer in shared memory by cp.async one sample ahead, so the compute phase never waits on global memory and a
// sample costs one block barrier; the window geometry of a lane is fixed for the whole kernel (registers), and the
// arg-max position is decoded from the index map by one subtraction and one comparison against the window's first flat
// index - no integer division in the sample loop.  (ncu of the register-prefetch kernel: 75 % of the issued instructions
// were address arithmetic, 15 % of the stall samples waits on the prefetch, 8 % barriers.)  2 x 2 windows only.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_copy_16B(void* dst, const void* src) {
#ifdef DLB_CUDA_EMU
    for (int q = 0; q < 16; ++q) static_cast<char*>(dst)[q] = static_cast<const char*>(src)[q];
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#endif
}

template <class R, int C, int KH, int KW, int WPL, int FMAX>
__global__ void __launch_bounds__(32 * FMAX, 1) conv_pool_backward_staged_kernel(const R* __restrict__ x, const int* __restrict__ gather_base,
                                                                                   const int* __restrict__ cursor, const R* __restrict__ dpool,
                                                                                   const int* __restrict__ map, const R* __restrict__ pooled,
                                                                                   R* __restrict__ partial, int B, ConvPoolGeom g, int act, int n_par,
                                                                                   int buf_bytes, int vec_windows) {
    DLB_DYNAMIC_SMEM(unsigned char, staged_smem);
    const int* gather = resolve_gather(gather_base, cursor);
    const int tid = threadIdx.x, nt = blockDim.x;
    // the gather indices of the first two samples are requested before anything else: their latency overlaps the set-up
    auto row_of = [&](int b) -> int { return b < B ? (gather ? __ldg(gather + b) : b) : -1; };
    const int G = gridDim.x;
    const int row0 = row_of(blockIdx.x);
    int row_next = row_of((int)blockIdx.x + G);
    const int f = tid / 32, lane = tid % 32;
    constexpr int K = C * KH * KW;
    const int NW = g.F * g.NPOS;                       // windows per sample
    const int img_elems = C * g.Hp * g.bRS;
    // buffer layout: image [C][Hp][bRS] (even columns, then odd columns at +bHO) | pooled [NW] | dpool [NW] | map [NW]
    const int off_wp = (img_elems * (int)sizeof(R) + 15) & ~15;
    const int off_wd = off_wp + ((NW * (int)sizeof(R) + 15) & ~15);
    const int off_wm = off_wd + ((NW * (int)sizeof(R) + 15) & ~15);
    for (int i = tid; i < 2 * buf_bytes / 4; i += nt) reinterpret_cast<int*>(staged_smem)[i] = 0;   // image borders stay zero
    __syncthreads();

    const int warp = tid >> 5, nw = nt >> 5;
    const int n_in = C * g.H * g.W;
    auto stage = [&](int b, int row, unsigned char* buf) {
        R* img = reinterpret_cast<R*>(buf);
        const R* xs = x + (size_t)row * n_in;
        for (int r = warp; r < C * g.H; r += nw) {      // a warp copies whole image rows, lane = column
            const int c = C == 1 ? 0 : r / g.H;
            const int yy = r - c * g.H;
            R* drow = img + (c * g.Hp + yy + g.ph) * g.bRS;
            for (int xx = lane; xx < g.W; xx += 32) {
                const int col = xx + g.pw;
                stage_copy_elem(drow + (col & 1) * g.bHO + (col >> 1), xs + r * g.W + xx);
            }
        }
        const size_t o = (size_t)b * NW;
        if (vec_windows) {
            const int nr = NW * (int)sizeof(R) / 16, ni = NW * 4 / 16;
            for (int i = tid; i < nr; i += nt) {
                stage_copy_16B(buf + off_wp + 16 * i, reinterpret_cast<const unsigned char*>(pooled + o) + 16 * i);
                stage_copy_16B(buf + off_wd + 16 * i, reinterpret_cast<const unsigned char*>(dpool + o) + 16 * i);
            }
            for (int i = tid; i < ni; i += nt) stage_copy_16B(buf + off_wm + 16 * i, reinterpret_cast<const unsigned char*>(map + o) + 16 * i);
        } else {
            for (int i = tid; i < NW; i += nt) {
                stage_copy_elem(reinterpret_cast<R*>(buf + off_wp) + i, pooled + o + i);
                stage_copy_elem(reinterpret_cast<R*>(buf + off_wd) + i, dpool + o + i);
                stage_copy_elem(reinterpret_cast<int*>(buf + off_wm) + i, map + o + i);
            }
        }
    };
    R dk[K];
    R db[WPL][4];
    int base_id[WPL], addr0[WPL];     // first flat index of the lane's window, shared-memory offset of its first element
#pragma unroll
    for (int k = 0; k < K; ++k) dk[k] = R(0);
#pragma unroll
    for (int wi = 0; wi < WPL; ++wi) {
#pragma unroll
        for (int q = 0; q < 4; ++q) db[wi][q] = R(0);
        const int win = min(lane + 32 * wi, g.NPOS - 1);
        const int py = win / g.POW, px = win - py * g.POW;
        base_id[wi] = (f * g.OH + 2 * py) * g.OW + 2 * px;
        addr0[wi] = 2 * py * g.bRS + px;
    }

    if (row0 >= 0) stage(blockIdx.x, row0, staged_smem);
    stage_commit();
    int it = 0;
    for (int b = blockIdx.x; b < B; b += G, ++it) {
        stage_wait<0>();
        __syncthreads();
        const unsigned char* buf = staged_smem + (it & 1) * buf_bytes;
        if (row_next >= 0) stage(b + G, row_next, staged_smem + ((it + 1) & 1) * buf_bytes);
        stage_commit();
        row_next = row_of(b + 2 * G);
        const R* x_s = reinterpret_cast<const R*>(buf);
        const R* wp_s = reinterpret_cast<const R*>(buf + off_wp) + f * g.NPOS;
        const R* wd_s = reinterpret_cast<const R*>(buf + off_wd) + f * g.NPOS;
        const int* wm_s = reinterpret_cast<const int*>(buf + off_wm) + f * g.NPOS;
#pragma unroll
        for (int wi = 0; wi < WPL; ++wi) {
            const int win = lane + 32 * wi;
            if (win >= g.NPOS) continue;
            const int m = wm_s[win];
            if (m < 0) continue;
            const R gval = deriv_from_value<R>(act, wp_s[win]) * wd_s[win];
            const int diff = m - base_id[wi];
            const int dy = diff >= g.OW ? 1 : 0;
            const int dx = diff - dy * g.OW;
            const int sel = dy * 2 + dx;
#pragma unroll
            for (int q = 0; q < 4; ++q) db[wi][q] += (q == sel) ? gval : R(0);
            if (gval != R(0)) {   // the reference skips zero factors as well (Tensor.cpp:484-486)
                // column 2 px + dx: even taps come from the half with the column's parity, odd taps from the other half
                const R* pa = x_s + addr0[wi] + dy * g.bRS + dx * g.bHO;
                const R* pb = x_s + addr0[wi] + dy * g.bRS + (dx ? 1 : g.bHO);
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int ky = 0; ky < KH; ++ky) {
                        const int ro = (c * g.Hp + ky) * g.bRS;
#pragma unroll
                        for (int kx = 0; kx < KW; ++kx)
                            dk[(c * KH + ky) * KW + kx] += ((kx & 1) ? pb[ro + (kx >> 1)] : pa[ro + (kx >> 1)]) * gval;
                    }
            }
        }
    }
    stage_wait<0>();

    R* pb = partial + (size_t)blockIdx.x * n_par;
#pragma unroll
    for (int wi = 0; wi < WPL; ++wi) {
        const int win = lane + 32 * wi;
        if (win >= g.NPOS) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) pb[base_id[wi] + (q >> 1) * g.OW + (q & 1)] = db[wi][q];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const R v = warp_sum<R>(dk[k]);
        if (lane == 0) pb[g.F * g.OH * g.OW + f * K + k] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The head: last NLayer (n <= NMAX outputs) forward, activation (incl. whole-vector softmax), cost derivative,
// activation gradient g = dC/dz, and the input gradient dx = W^T g (optionally times the previous layer's activation
// derivative recomputed from x, which IS that layer's output).  One warp per sample; W lives in shared memory.
// ---------------------------------------------------------------------------------------------------------------
// Split-K partials of the previous NLayer's forward contraction, when that launch left them unfinalised (in_n <= 128): the head then
// computes its own input x = f(sum of the partials + bias) - same association as gemm_finalize_kernel, i.e. the same bits - and stores it
// for the weight gradients; one dependent kernel less on the critical path of a small mini-batch.
template <class R>
struct HeadInputPartials { const R* partials; int splits; const R* bias; int act; R* x_store; };

template <class R, int NMAX>
__global__ void full_head_kernel(const R* __restrict__ x, const R* __restrict__ bias, const R* __restrict__ W, const R* __restrict__ ref,
                                 const int* __restrict__ gather_base, const int* __restrict__ cursor, R* __restrict__ out,
                                 R* __restrict__ g_out, R* __restrict__ dx, int B, int in_n, int n, int act, int cost_id, int prev_act,
                                 HeadInputPartials<R> hp) {
    DLB_DYNAMIC_SMEM(R, W_s);
    pdl_trigger();   // the input-gradient contraction that follows may set itself up while this grid runs
    const int* gather = resolve_gather(gather_base, cursor);
    const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
    // the first sample's operands are requested before the weights are staged: their latencies overlap
    constexpr int XP = 4;
    const int b_first = blockIdx.x * wpb + threadIdx.x / 32;
    const bool pre = in_n <= 32 * XP;
    auto load_x = [&](int b, int k) -> R {
        if (!hp.partials) return x[(size_t)b * in_n + k];
        R s8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s8[j] = R(0);
        const size_t total = (size_t)B * in_n, e = (size_t)b * in_n + k;
        for (int z = 0; z < hp.splits; z += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (z + j < hp.splits) s8[j] += hp.partials[(size_t)(z + j) * total + e];
        }
        R d;
        const R v = apply_activation<R>(hp.act, (((s8[0] + s8[1]) + (s8[2] + s8[3])) + ((s8[4] + s8[5]) + (s8[6] + s8[7]))) + hp.bias[k], d);
        hp.x_store[e] = v;
        return v;
    };
    R xpre[XP];
#pragma unroll
    for (int j = 0; j < XP; ++j) {
        const int k = lane + 32 * j;
        xpre[j] = (pre && b_first < B && k < in_n) ? load_x(b_first, k) : R(0);
    }
    size_t row_pre = (b_first < B && gather) ? (size_t)gather[b_first] : (size_t)b_first;
    for (int i = threadIdx.x; i < n * in_n; i += blockDim.x) W_s[i] = W[i];
    __syncthreads();
    for (int b = b_first; b < B; b += gridDim.x * wpb) {
        const R* xb = x + (size_t)b * in_n;
        const size_t row = b == b_first ? row_pre : (gather ? (size_t)gather[b] : (size_t)b);
        const int mine = lane < n ? lane : 0;
        const R refl = ref[row * n + mine];
        R z[NMAX];
#pragma unroll
        for (int nn = 0; nn < NMAX; ++nn) z[nn] = R(0);
        R xcur[XP];
        if (pre) {
#pragma unroll
            for (int j = 0; j < XP; ++j) {
                const int k = lane + 32 * j;
                xcur[j] = R(0);
                if (k < in_n) {
                    const R xv = b == b_first ? xpre[j] : load_x(b, k);
                    xcur[j] = xv;
#pragma unroll
                    for (int nn = 0; nn < NMAX; ++nn)
                        if (nn < n) z[nn] += W_s[nn * in_n + k] * xv;
                }
            }
        } else {
            for (int k = lane; k < in_n; k += 32) {
                const R xv = xb[k];
#pragma unroll
                for (int nn = 0; nn < NMAX; ++nn)
                    if (nn < n) z[nn] += W_s[nn * in_n + k] * xv;
            }
        }
#pragma unroll
        for (int nn = 0; nn < NMAX; ++nn)
            if (nn < n) z[nn] = warp_sum<R>(z[nn]) + bias[nn];

        // From here on lane nn owns output nn (lanes >= n mirror output 0 and are discarded): the activation, the cost
        // derivative and the activation gradient are evaluated once per output instead of n times per lane; the sums over
        // the outputs run in ascending order over shuffled values, i.e. in the order of the reference's loops.
        R zl = z[0];
#pragma unroll
        for (int nn = 1; nn < NMAX; ++nn)
            if (nn < n && mine == nn) zl = z[nn];
        R yl, auxl, ggl;
        if (act == DLB_ACT_SOFTMAX) {
            R m = z[0];
#pragma unroll
            for (int nn = 1; nn < NMAX; ++nn)
                if (nn < n && m < z[nn]) m = z[nn];
            auxl = dl_exp<R>(zl - m);
            R soft_sum = R(0);
#pragma unroll
            for (int nn = 0; nn < NMAX; ++nn)
                if (nn < n) soft_sum += __shfl_sync(0xffffffffu, auxl, nn);
            yl = auxl * (R(1) / soft_sum);
            const R ood = R(1) / soft_sum;                        // ActivationFunctionHelper.cpp:50-58
            ggl = auxl * cost_derivative<R>(cost_id, yl, refl) * ood;
            R t = R(0);
#pragma unroll
            for (int nn = 0; nn < NMAX; ++nn)
                if (nn < n) t += __shfl_sync(0xffffffffu, ggl, nn);
            ggl = ggl - auxl * (t * ood);
        } else {
            yl = apply_activation<R>(act, zl, auxl);
            ggl = auxl * cost_derivative<R>(cost_id, yl, refl);
        }
        if (lane < n) {
            g_out[(size_t)b * n + lane] = ggl;
            if (out) out[(size_t)b * n + lane] = yl;
        }
        R gg[NMAX];
#pragma unroll
        for (int nn = 0; nn < NMAX; ++nn) gg[nn] = nn < n ? __shfl_sync(0xffffffffu, ggl, nn) : R(0);
        if (dx) {
            if (pre) {   // (the inputs are still in registers)
#pragma unroll
                for (int j = 0; j < XP; ++j) {
                    const int k = lane + 32 * j;
                    if (k < in_n) {
                        R s = R(0);
#pragma unroll
                        for (int nn = 0; nn < NMAX; ++nn)
                            if (nn < n) s = W_s[nn * in_n + k] * gg[nn] + s;
                        if (prev_act) s = deriv_from_value<R>(prev_act, xcur[j]) * s;
                        dx[(size_t)b * in_n + k] = s;
                    }
                }
            } else {
                for (int k = lane; k < in_n; k += 32) {
                    R s = R(0);
#pragma unroll
                    for (int nn = 0; nn < NMAX; ++nn)
                        if (nn < n) s = W_s[nn * in_n + k] * gg[nn] + s;
                    if (prev_act) s = deriv_from_value<R>(prev_act, xb[k]) * s;
                    dx[(size_t)b * in_n + k] = s;
                }
            }
        }
    }
}

// Weight and bias gradient of a narrow layer: partial[chunk][bias n | W n*in_n] = sum over the chunk's samples.
template <class R>
__global__ void batch_outer_partial_kernel(const R* __restrict__ g, const R* __restrict__ x, R* __restrict__ partial,
                                           int B, int n, int in_n, int chunk) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int cols = in_n + 1;
    if (e >= n * cols) return;
    const int nn = e / cols, k = e - nn * cols;
    const int b0 = blockIdx.y * chunk, b1 = min(B, b0 + chunk);
    R acc = R(0);
    for (int b = b0; b < b1; ++b) acc += g[(size_t)b * n + nn] * (k == in_n ? R(1) : x[(size_t)b * in_n + k]);
    partial[(size_t)blockIdx.y * (n * cols) + (k == in_n ? nn : n + nn * in_n + k)] = acc;
}

// Hidden-NLayer gradient contraction with the bias folded in as a column of ones:
//   i = output n, j = input k (j == in_n: the bias), red = sample b.
template <class R>
struct DenseColOnesLoader {  // value = j < ld ? base[row(red) * ld + j] : 1
    static constexpr bool kRedFast = false;
    const R* base; size_t ld; const int* gather_base; const int* cursor;
    __device__ __forceinline__ R operator()(int o, int r) const {
        if ((size_t)o >= ld) return R(1);
        const int* gather = resolve_gather(gather_base, cursor);
        const size_t row = gather ? (size_t)gather[r] : (size_t)r;
        return base[row * ld + o];
    }
    // strided protocol (gemm.cuh): outer indices from outer_limit() on read as 1 (the bias column)
    __host__ __device__ bool strided() const { return gather_base == nullptr; }
    __device__ __forceinline__ const R* origin() const { return base; }
    __device__ __forceinline__ size_t outer_step() const { return 1; }
    __device__ __forceinline__ size_t red_step() const { return ld; }
    __device__ __forceinline__ int outer_limit() const { return (int)ld; }
};

template <class R>
struct EpiWeightAndBias {  // dW[n][k] for k < in_n, db[n] for k == in_n; plain store (the batch sum overwrites, Net.inl:259)
    R* db; R* dW; int in_n;
    __device__ __forceinline__ void operator()(int n, int k, R acc) const {
        if (k < in_n) dW[(size_t)n * in_n + k] = acc;
        else db[n] = acc;
    }
};

template <class R>
struct EpiStoreTimesDeriv {  // input gradient times the previous layer's activation derivative (from its output values)
    R* out; size_t ld; const R* prev_y; int prev_act;
    __device__ __forceinline__ void operator()(int i, int j, R acc) const {
        const size_t o = (size_t)i * ld + j;
        out[o] = prev_act ? deriv_from_value<R>(prev_act, prev_y[o]) * acc : acc;
    }
    __device__ __forceinline__ bool row4_ok() const {
        return (ld & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) | (prev_act ? reinterpret_cast<uintptr_t>(prev_y) : 0)) & 15) == 0;
    }
    __device__ __forceinline__ void row4(int i, int j, float4 a) const {
        const size_t o = (size_t)i * ld + j;
        if (prev_act) {
            const float4 p = *reinterpret_cast<const float4*>(prev_y + o);
            a.x = deriv_from_value<R>(prev_act, p.x) * a.x; a.y = deriv_from_value<R>(prev_act, p.y) * a.y;
            a.z = deriv_from_value<R>(prev_act, p.z) * a.z; a.w = deriv_from_value<R>(prev_act, p.w) * a.w;
        }
        *reinterpret_cast<float4*>(out + o) = a;
    }
};

template <class R>
struct DenseRowCursorLoader {  // DenseRowLoader whose gather slice is addressed through the device-side cursor
    static constexpr bool kRedFast = true;
    const R* base; size_t ld; const int* gather_base; const int* cursor;
    __device__ __forceinline__ R operator()(int o, int r) const {
        const int* gather = resolve_gather(gather_base, cursor);
        const size_t row = gather ? (size_t)gather[o] : (size_t)o;
        return base[row * ld + r];
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Partial-sum reduction fused with the SGD update.
// ---------------------------------------------------------------------------------------------------------------
struct ReduceSegments {
    int count;
    long long end[kMaxSegments];           // exclusive end offset in the flat parameter buffer
    unsigned char regularised[kMaxSegments];
    const void* partial[kMaxSegments];     // nullptr: the gradient is already final in the flat gradient buffer
    int chunks[kMaxSegments];
    long long stride[kMaxSegments];        // elements between consecutive chunks
    long long local_off[kMaxSegments];     // offset of the segment's first element inside one chunk
};

// Blocks [0, partial_blocks) own the elements whose gradient arrives as per-CTA / per-chunk partial sums: 16 elements
// per block, 16 threads per element, each thread adds every 16th partial (independent loads in flight), then one
// thread adds the 16 sub-sums in a fixed order - deterministic, and ~chunks/16 dependent steps instead of `chunks`.
// The remaining blocks own 256 elements each whose gradient is already final in `grads`.
template <class R>
__global__ void reduce_update_kernel(R* __restrict__ p, R* __restrict__ grads, R* __restrict__ publish0, size_t publish_stride,
                                     const unsigned long long* __restrict__ publish_counter, size_t n, ReduceSegments segs, int partial_blocks,
                                     long long n_partial_elems, int do_update, R param_scale, R grad_scale, int nontrivial_reg,
                                     int* __restrict__ cursor, int advance) {
    __shared__ R sub[16][17];
    if (blockIdx.x == 0 && threadIdx.x == 0 && cursor) *cursor += advance;   // every reader of the cursor ran earlier in the stream
    R* publish = publish0 ? publish0 + (size_t)(*publish_counter & 1) * publish_stride : nullptr;   // the half peers read this step
    long long e; R gsum; int s = 0; bool active;
    if ((int)blockIdx.x < partial_blocks) {
        // the partial elements are enumerated segment by segment in flat order
        const int el = threadIdx.x & 15, part = threadIdx.x >> 4;
        long long idx = (long long)blockIdx.x * 16 + el;
        active = idx < n_partial_elems;
        e = 0;
        if (active) {
            long long begin = 0;
            for (s = 0; s < segs.count; ++s) {
                begin = s == 0 ? 0 : segs.end[s - 1];
                if (!segs.partial[s]) continue;
                const long long len = segs.end[s] - begin;
                if (idx < len) break;
                idx -= len;
            }
            e = begin + idx;
            const R* part_ptr = static_cast<const R*>(segs.partial[s]) + segs.local_off[s] + idx;
            R a = R(0);
            for (int c = part; c < segs.chunks[s]; c += 16) a += part_ptr[(long long)c * segs.stride[s]];
            sub[el][part] = a;
        } else {
            sub[el][part] = R(0);
        }
        __syncthreads();
        if (part != 0 || !active) return;
        gsum = R(0);
#pragma unroll
        for (int q = 0; q < 16; ++q) gsum += sub[el][q];
        grads[e] = gsum;
        if (publish) publish[e] = gsum;
    } else {
        e = (long long)(blockIdx.x - partial_blocks) * blockDim.x + threadIdx.x;
        if (e >= (long long)n) return;
        while (s < segs.count - 1 && e >= segs.end[s]) ++s;
        if (segs.partial[s]) return;     // owned by a partial block
        gsum = grads[e];
        if (publish) publish[e] = gsum;
    }
    if (do_update) {
        const bool reg = nontrivial_reg && segs.regularised[s];
        const R xv = p[e];
        p[e] = reg ? xv * param_scale + gsum * grad_scale : xv + gsum * grad_scale;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Data parallelism without NCCL on the step's critical path: ONE kernel per rank that (1) tells every peer "my local
// gradient sum for step t is published" and waits for theirs (flags in peer memory, written over NVLink), (2) reads all
// ranks' gradient buffers through peer pointers (NVLink P2P loads), adds them in rank order - so every rank computes the
// bit-identical total - and (3) applies the SGD update to its own parameters in the same pass (ALayer.inl:97-122).
// The published buffers are double-buffered by step parity: a rank can only be overwriting the buffer of step t when it
// is at step t+2, which it cannot reach before every peer has passed the barrier of step t+1, i.e. finished reading
// step t.  Replaces ncclAllReduce + sgd_update (a 1.2 MB message: latency-bound).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
struct PeerTable {
    const unsigned char* region[kMaxPeers];   // peer r's shared region: [parity 0: kMaxPeers slots | parity 1: kMaxPeers slots | flags[kMaxPeers] | local release word]
    size_t publish_bytes;                     // bytes of one slot (the flat gradient, rounded up to 256)
};

// Data-parallel finish over NVLink peer memory, ONE kernel per rank and step.  Rank r's gradient sum of its share of the mini-batch sits
// in slot r of its own region (written by the reduction kernel in front of this one).
//   pull (default)  the ranks barrier through flags in peer memory, then every rank READS slot r of peer r's region.
//   push            every rank first WRITES its slot into slot r of every peer's region (posted stores), then the barrier, then each rank
//                   sums the `world` slots of its OWN region (local reads).  DLB_P2P_PUSH=1.
// Measured equal (8 GPUs: 154.0 us per step pull, 158.3 push; 2 GPUs: 309 / 312): the kernel's time is the wait for the slowest rank at the
// barrier (block 0 passes it 15-33 us after the kernel starts, DLB_P2P_PROF=1), not the transfer of 1.2 MB per peer.
// Either way the slots are summed in rank order (bit-identical total on every rank) and the update is applied in the same pass.
// The step counter lives on the device so that a captured CUDA graph of the step stays valid from step to step: the reduction kernel and
// this kernel read it (parity = which half is in use, value + 1 = the barrier ticket); the LAST block to finish advances it.
// All CTAs of the grid are co-resident (2 x 148 x 256 threads), which the in-kernel waits rely on.
template <class R> struct alignas(16) Vec16 { R v[16 / sizeof(R)]; };

template <class R>
__global__ void __launch_bounds__(256) p2p_reduce_update_kernel(R* __restrict__ p, R* __restrict__ grads_total, size_t n, PeerTable peers, int world, int rank,
                                         unsigned long long* __restrict__ counter, unsigned int* __restrict__ blocks_done, UpdateSegments segs,
                                         R param_scale, R grad_scale, int nontrivial_reg, int* __restrict__ cursor, int advance, int push,
                                         unsigned long long* __restrict__ prof) {   // prof (development, nullable): [launches, ns to barrier passed, ns to data done] of block 0
    constexpr int V = 16 / (int)sizeof(R);
    unsigned long long t_in = 0;
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_in));
    const unsigned long long t = *counter;
    const unsigned long long step = t + 1;
    const size_t half = (size_t)(t & 1) * kMaxPeers * peers.publish_bytes;
    const size_t flags_off = 2 * (size_t)kMaxPeers * peers.publish_bytes;
    const size_t nv = n / V;
    unsigned int* push_done = blocks_done + 1;
    __shared__ int last_pusher;
    auto post_flags = [&]() {   // threads < world of ONE block: tell every peer that this rank's slot is complete
        if ((int)threadIdx.x < world) {
            __threadfence_system();
            volatile unsigned long long* remote =
                reinterpret_cast<unsigned long long*>(const_cast<unsigned char*>(peers.region[threadIdx.x]) + flags_off) + rank;
            *remote = step;
        }
    };
    if (push) {
        // (0) this rank's slot -> slot `rank` of every peer (16-byte posted stores; the tail elements as scalars)
        const unsigned char* own = peers.region[rank] + half + (size_t)rank * peers.publish_bytes;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
            const Vec16<R> v = reinterpret_cast<const Vec16<R>*>(own)[i];
            for (int r = 0; r < world; ++r)
                if (r != rank) reinterpret_cast<Vec16<R>*>(const_cast<unsigned char*>(peers.region[r]) + half + (size_t)rank * peers.publish_bytes)[i] = v;
        }
        for (size_t e = nv * V + (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
            const R v = reinterpret_cast<const R*>(own)[e];
            for (int r = 0; r < world; ++r)
                if (r != rank) reinterpret_cast<R*>(const_cast<unsigned char*>(peers.region[r]) + half + (size_t)rank * peers.publish_bytes)[e] = v;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) last_pusher = atomicAdd(push_done, 1u) == gridDim.x - 1 ? 1 : 0;
        __syncthreads();
        if (last_pusher) {   // every block's stores are out: raise the flags
            if (threadIdx.x == 0) *push_done = 0u;
            post_flags();
        }
    } else if (blockIdx.x == 0) {
        post_flags();   // the local slot was written by the previous kernel in this stream, so it is complete
    }
    // (1) cross-GPU barrier: wait until every rank's flag carries this step's ticket.  Only block 0 polls the flags the peers write
    // (tools/flag_probe.cu: a flag hop takes 3.8 us with one polling block and 8.8 us when all 296 blocks poll the same line); it then
    // releases the other blocks through a local word.
    volatile unsigned long long* go = reinterpret_cast<unsigned long long*>(const_cast<unsigned char*>(peers.region[rank]) + flags_off) + kMaxPeers;
    if (blockIdx.x == 0) {
        if ((int)threadIdx.x < world) {
            volatile unsigned long long* mine =
                reinterpret_cast<unsigned long long*>(const_cast<unsigned char*>(peers.region[rank]) + flags_off) + threadIdx.x;
            long long spin = 0;
            while (*mine < step) { if (++spin > (1LL << 31)) __trap(); }   // bounded: a lost peer becomes a launch failure, not a hang
            __threadfence_system();
        }
        __syncthreads();
        if (threadIdx.x == 0) { *go = step; __threadfence(); }
    } else {
        if (threadIdx.x == 0) {
            long long spin = 0;
            while (*go < step) { __nanosleep(64); if (++spin > (1LL << 28)) __trap(); }
            __threadfence_system();
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && cursor) *cursor += advance;
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long tb; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tb)); atomicAdd(prof, 1ULL); atomicAdd(prof + 1, tb - t_in); }
    // (2) + (3): 16 bytes per thread and rank, the loads of ALL ranks in flight together, the sum in rank order (identical on every rank), the update
    auto slot = [&](int r) -> const unsigned char* { return peers.region[push ? rank : r] + half + (size_t)r * peers.publish_bytes; };
    auto update_one = [&](size_t e, R gsum) {
        grads_total[e] = gsum;
        int s = 0;
        while (s < segs.count - 1 && (long long)e >= segs.end[s]) ++s;
        const bool reg = nontrivial_reg && segs.regularised[s];
        const R xv = p[e];
        p[e] = reg ? xv * param_scale + gsum * grad_scale : xv + gsum * grad_scale;
    };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
        Vec16<R> acc;
#pragma unroll
        for (int q = 0; q < V; ++q) acc.v[q] = R(0);
        for (int r0 = 0; r0 < world; r0 += 8) {
            Vec16<R> v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = reinterpret_cast<const Vec16<R>*>(slot(min(r0 + r, world - 1)))[i];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r0 + r < world) {
#pragma unroll
                    for (int q = 0; q < V; ++q) acc.v[q] += v[r].v[q];
                }
        }
#pragma unroll
        for (int q = 0; q < V; ++q) update_one(i * V + q, acc.v[q]);
    }
    for (size_t e = nv * V + (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        R gsum = R(0);
        for (int r = 0; r < world; ++r) gsum += reinterpret_cast<const R*>(slot(r))[e];
        update_one(e, gsum);
    }
    // (4) the last block out advances the step counter (every block read it at the top)
    __syncthreads();
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long te; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(te)); atomicAdd(prof + 2, te - t_in); }
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(blocks_done, 1u) == gridDim.x - 1) { *blocks_done = 0u; *counter = step; }
    }
}

}  // namespace dlb

// =================================================================================================================
// Host-side dispatch of the fused first stage.
// =================================================================================================================
namespace dlb {

struct ConvPoolPlan {
    bool ok = false;
    ConvPoolGeom g{};
    int KH = 0, KW = 0, fpt = 1, wpl = 1, act = 0, is_max = 1;
    int pp = 1, spc = 1, tps = 0;     // forward: pooled positions per thread, samples per CTA iteration, threads per sample
    int fwd_threads = 0, bwd_threads = 0;
    size_t fwd_smem_elems = 0, bwd_smem_elems = 0;
    int n_par = 0;
};

// Which CLayer + PLayer pairs the fused kernels cover: stride 1, 5x5 (1 channel) or 3x3 (1 or 3 channels) filters,
// at most 32 filters, 2x2 max/min pooling, element-wise activation.  Everything else takes the unfused operators.
inline ConvPoolPlan plan_conv_pool(const dlb_layer_spec& c, const dlb_layer_spec& p, size_t real_bytes) {
    ConvPoolPlan pl;
    if (c.type != DLB_LAYER_CONVOLUTION || p.type != DLB_LAYER_PULL) return pl;
    if (c.stride_h != 1 || c.stride_w != 1 || c.activation == DLB_ACT_SOFTMAX) return pl;
    if (p.pool_type != DLB_POOL_MAX && p.pool_type != DLB_POOL_MIN) return pl;
    if (p.win_h != 2 || p.win_w != 2) return pl;
    const bool k55 = c.win_h == 5 && c.win_w == 5 && c.in_c == 1;
    const bool k33 = c.win_h == 3 && c.win_w == 3 && (c.in_c == 1 || c.in_c == 3);
    if (!k55 && !k33) return pl;
    if (c.out_c > 32) return pl;
    ConvPoolGeom& g = pl.g;
    g.C = c.in_c; g.H = c.in_h; g.W = c.in_w; g.F = c.out_c; g.ph = c.pad_h; g.pw = c.pad_w;
    g.OH = conv_out_size(g.H, c.win_h, g.ph, 1); g.OW = conv_out_size(g.W, c.win_w, g.pw, 1);
    g.POH = conv_out_size(g.OH, 2, 0, 2); g.POW = conv_out_size(g.OW, 2, 0, 2);
    g.Hp = g.H + 2 * g.ph; g.Wp = g.W + 2 * g.pw; g.K = g.C * c.win_h * c.win_w; g.NPOS = g.POH * g.POW;
    if (g.POH < 1 || g.POW < 1 || g.NPOS > 256) return pl;
    pl.KH = c.win_h; pl.KW = c.win_w; pl.act = c.activation; pl.is_max = p.pool_type == DLB_POOL_MAX;
    pl.pp = g.POW >= 2 ? 2 : 1;
    const int npg = g.POH * ceil_div(g.POW, pl.pp);
    pl.fpt = 1;
    for (int fpt : {8, 5, 4, 2, 1})
        if (g.F % fpt == 0 && (long)npg * (g.F / fpt) <= 640) { pl.fpt = fpt; break; }
    pl.tps = npg * ceil_div(g.F, pl.fpt);
    if (pl.tps > 640) return pl;
    pl.spc = std::max(1, std::min(4, 640 / pl.tps));
    pl.fwd_threads = ceil_div(pl.spc * pl.tps, 32) * 32;
    pl.bwd_threads = 32 * g.F;
    const int need_wpl = ceil_div(g.NPOS, 32);
    pl.wpl = need_wpl <= 1 ? 1 : need_wpl <= 2 ? 2 : need_wpl <= 5 ? 5 : 8;
    pl.fwd_smem_elems = (size_t)g.F * g.K + (size_t)pl.spc * g.C * g.Hp * g.Wp + 8 /*vector over-read of the last patch row*/;
    {   // backward image layout: even row stride >= padded width whose doubled value lands next to POW modulo the 32 banks
        const int min_rs = 2 * ceil_div(g.Wp, 2);
        int best = min_rs, best_score = 1 << 30;
        for (int rs = min_rs; rs < min_rs + 32; rs += 2) {
            const int d = ((2 * rs) % 32) - (g.POW % 32);
            const int score = d >= 0 ? d : 64 + d;      // prefer the smallest gap that does not overlap
            if (score < best_score) { best_score = score; best = rs; }
        }
        g.bRS = best; g.bHO = best / 2;
    }
    pl.bwd_smem_elems = (size_t)g.C * g.Hp * g.bRS;
    if (pl.fwd_smem_elems * real_bytes > 200 * 1024) return pl;
    pl.n_par = g.F * g.OH * g.OW + g.F * g.K;
    pl.ok = true;
    return pl;
}

template <class R, int KH, int KW>
void conv_pool_forward_fpt(cudaStream_t st, const ConvPoolPlan& pl, int grid, const R* x, const int* gather_base, const int* cursor,
                           const R* bias, const R* filt, R* pooled, int* map, int B) {
    const size_t smem = pl.fwd_smem_elems * sizeof(R);
    const dim3 gd((unsigned)grid), bd((unsigned)pl.fwd_threads);
#define DLB_CPF3(FPT, PP, TMAX)                                                                                                  \
    DLB_LAUNCH_SMEM((conv_pool_forward_kernel<R, KH, KW, FPT, PP, TMAX>), gd, bd, smem, st, x, gather_base, cursor, bias, filt,     \
                    pooled, map, B, pl.g, pl.act, pl.is_max, pl.spc, pl.tps)
#define DLB_CPF2(FPT, PP)                                          \
    do {                                                           \
        if (pl.fwd_threads <= 320) DLB_CPF3(FPT, PP, 320);         \
        else DLB_CPF3(FPT, PP, 640);                               \
    } while (0)
#define DLB_CPF(FPT)                                 \
    do {                                             \
        if (pl.pp == 2) DLB_CPF2(FPT, 2);            \
        else DLB_CPF2(FPT, 1);                       \
    } while (0)
    switch (pl.fpt) {
    case 8: DLB_CPF(8); break;
    case 5: DLB_CPF(5); break;
    case 4: DLB_CPF(4); break;
    case 2: DLB_CPF(2); break;
    default: DLB_CPF(1); break;
    }
#undef DLB_CPF
#undef DLB_CPF2
#undef DLB_CPF3
}

// The pair kernel (one sample per CTA iteration; it sizes its own grid): FFMA2 for float, DFMA pairs for double.
template <class R, int KH, int KW>
inline bool conv_pool_forward_x2(cudaStream_t st, const ConvPoolPlan& pl, const R* x, const int* gather_base, const int* cursor,
                                 const R* bias, const R* filt, R* pooled, int* map, int B) {
    const ConvPoolGeom& g = pl.g;
    const int PGW = ceil_div(g.POW, 2), NPG = g.POH * PGW;
    if (NPG > kX2Threads) return false;
    const int nfg = ceil_div(g.F, 4);
    const int fgp = std::min(nfg, kX2Threads / NPG);
    const int threads = ceil_div(fgp * NPG, 32) * 32;
    const int RS = (std::max(g.Wp, PGW * 4 + KW + 1) + 1) & ~1;   // pair slots per image row: every 16-byte patch load stays inside its row
    const size_t smem = ((size_t)g.K * nfg * 4 + 2 * (size_t)g.C * g.Hp * RS * 2) * sizeof(R);
    if (smem > 200 * 1024) return false;
    int flags = 0;
    if (g.OW % 4 == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0) flags |= kX2VecBias;
    if (g.POW % 2 == 0 && g.NPOS % 2 == 0 && (reinterpret_cast<uintptr_t>(pooled) & 7) == 0 && (reinterpret_cast<uintptr_t>(map) & 7) == 0)
        flags |= kX2VecOut;
    const dim3 gd((unsigned)std::min(B, (sizeof(R) == 4 ? 2 : 1) * kNumSMs)), bd((unsigned)threads);   // float: two resident CTAs per SM
    DLB_LAUNCH_SMEM((conv_pool_forward_x2_kernel<R, KH, KW>), gd, bd, smem, st, x, gather_base, cursor, bias, filt, pooled, map, B, g, pl.act,
                    pl.is_max, nfg, fgp, RS, flags);
    return true;
}

template <class R>
void conv_pool_forward(cudaStream_t st, const ConvPoolPlan& pl, int grid, const R* x, const int* gather_base, const int* cursor,
                       const R* bias, const R* filt, R* pooled, int* map, int B) {
    TagScope tag_("conv_pool_fwd");
    if (pl.KH == 5 ? conv_pool_forward_x2<R, 5, 5>(st, pl, x, gather_base, cursor, bias, filt, pooled, map, B)
                   : conv_pool_forward_x2<R, 3, 3>(st, pl, x, gather_base, cursor, bias, filt, pooled, map, B))
        return;
    if (pl.KH == 5) conv_pool_forward_fpt<R, 5, 5>(st, pl, grid, x, gather_base, cursor, bias, filt, pooled, map, B);
    else conv_pool_forward_fpt<R, 3, 3>(st, pl, grid, x, gather_base, cursor, bias, filt, pooled, map, B);
}

template <class R, int C, int KH, int KW>
void conv_pool_backward_wpl(cudaStream_t st, const ConvPoolPlan& pl, int grid, const R* x, const int* gather_base, const int* cursor,
                            const R* dpool, const int* map, const R* pooled, R* partial, int B) {
    const size_t smem = pl.bwd_smem_elems * sizeof(R);
    const dim3 gd((unsigned)grid), bd((unsigned)pl.bwd_threads);
    // FMAX bounds the block size (one warp per filter) so that ptxas knows the register budget per thread.
#define DLB_CPB2(WPL, FMAX)                                                                                                     \
    DLB_LAUNCH_SMEM((conv_pool_backward_kernel<R, C, KH, KW, 2, 2, WPL, FMAX>), gd, bd, smem, st, x, gather_base, cursor, dpool, \
                    map, pooled, partial, B, pl.g, pl.act, pl.n_par)
#define DLB_CPB(WPL)                                          \
    do {                                                      \
        if (pl.g.F <= 8) DLB_CPB2(WPL, 8);                    \
        else if (pl.g.F <= 20) DLB_CPB2(WPL, 20);             \
        else DLB_CPB2(WPL, 32);                               \
    } while (0)
    switch (pl.wpl) {
    case 1: DLB_CPB(1); break;
    case 2: DLB_CPB(2); break;
    case 5: DLB_CPB(5); break;
    default: DLB_CPB(8); break;
    }
#undef DLB_CPB
#undef DLB_CPB2
}

// Staged backward kernel: launches when its double buffer fits in shared memory; returns false otherwise.
template <class R, int C, int KH, int KW>
bool conv_pool_backward_staged(cudaStream_t st, const ConvPoolPlan& pl, int grid, const R* x, const int* gather_base, const int* cursor,
                               const R* dpool, const int* map, const R* pooled, R* partial, int B) {
    const ConvPoolGeom& g = pl.g;
    const int NW = g.F * g.NPOS;
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t buf = up16((size_t)C * g.Hp * g.bRS * sizeof(R)) + 2 * up16((size_t)NW * sizeof(R)) + up16((size_t)NW * 4);
    if (2 * buf > 200 * 1024) return false;
    const bool vec = (NW * sizeof(R)) % 16 == 0 && (NW * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(dpool) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(pooled) & 15) == 0 && (reinterpret_cast<uintptr_t>(map) & 15) == 0;
    const dim3 gd((unsigned)grid), bd((unsigned)pl.bwd_threads);
#define DLB_CPBS2(WPL, FMAX)                                                                                                          \
    DLB_LAUNCH_SMEM((conv_pool_backward_staged_kernel<R, C, KH, KW, WPL, FMAX>), gd, bd, 2 * buf, st, x, gather_base, cursor, dpool, map, \
                    pooled, partial, B, g, pl.act, pl.n_par, (int)buf, vec ? 1 : 0)
#define DLB_CPBS(WPL)                                          \
    do {                                                       \
        if (g.F <= 8) DLB_CPBS2(WPL, 8);                       \
        else if (g.F <= 20) DLB_CPBS2(WPL, 20);                \
        else DLB_CPBS2(WPL, 32);                               \
    } while (0)
    switch (pl.wpl) {
    case 1: DLB_CPBS(1); break;
    case 2: DLB_CPBS(2); break;
    case 5: DLB_CPBS(5); break;
    default: DLB_CPBS(8); break;
    }
#undef DLB_CPBS
#undef DLB_CPBS2
    return true;
}

template <class R>
void conv_pool_backward(cudaStream_t st, const ConvPoolPlan& pl, int grid, const R* x, const int* gather_base, const int* cursor,
                        const R* dpool, const int* map, const R* pooled, R* partial, int B) {
    TagScope tag_("conv_pool_bwd");
    if (pl.KH == 5 ? conv_pool_backward_staged<R, 1, 5, 5>(st, pl, grid, x, gather_base, cursor, dpool, map, pooled, partial, B)
        : pl.g.C == 1 ? conv_pool_backward_staged<R, 1, 3, 3>(st, pl, grid, x, gather_base, cursor, dpool, map, pooled, partial, B)
                      : conv_pool_backward_staged<R, 3, 3, 3>(st, pl, grid, x, gather_base, cursor, dpool, map, pooled, partial, B))
        return;
    if (pl.KH == 5) conv_pool_backward_wpl<R, 1, 5, 5>(st, pl, grid, x, gather_base, cursor, dpool, map, pooled, partial, B);
    else if (pl.g.C == 1) conv_pool_backward_wpl<R, 1, 3, 3>(st, pl, grid, x, gather_base, cursor, dpool, map, pooled, partial, B);
    else conv_pool_backward_wpl<R, 3, 3, 3>(st, pl, grid, x, gather_base, cursor, dpool, map, pooled, partial, B);
}

}  // namespace dlb

// mega.cu — the LeNet training step as ONE persistent cooperative kernel (sm_100a), Real = float.
//
// What Net<D>::learn does for one mini-batch (Net.inl:226-317) - forward through CLayer / PLayer / NLayer / NLayer, cost
// derivative, backward, ordered gradient sums, regularised SGD update - is cut into six phases separated by grid barriers
// (~1 us each) instead of kernel launches (the round-1 step was 10 launches, 7 of them in a dependent chain whose launch
// and ramp latencies WERE the step time at mini-batch 256).  Each phase has the decomposition that suits it:
//
//   conv fwd   CTA = (filter PAIR, sample group).  One thread = 4 x 8 convolution outputs of one sample for both filters:
//              the two filters' taps travel as one 64-bit register pair, the input value is the broadcast scalar of
//              FFMA2 (fma.rn.f32x2), so the image sits in shared memory ONCE, in its natural layout, and the 25 tap pairs
//              stay in registers.  Untied bias (CLayer.inl:130), activation, 2 x 2 max/min with strict first-wins compare
//              (Tensor.cpp:700-709) in the epilogue; pooled value + 1-byte arg-max selector leave for HBM/L2.
//   fc1 fwd    CTA = (filter, sample group): the 144 inputs of the hidden NLayer that one filter produces are a K-slice of
//              W1; the slice (hi and lo TF32 parts, 133 KB) lives in shared memory and the CTA produces the slice's
//              partial pre-activations for its samples with tcgen05.mma kind::tf32 (3xTF32: lo*hi + hi*lo + hi*hi into one
//              fp32 accumulator in tensor memory).  20 partials per output instead of a split-K over 148 CTAs.
//   head       partial sum + bias + activation, head NLayer forward, softmax / activation, cost derivative
//              (CostFunctionFactory.h:29-55 expanded), head backward; one warp per sample, sums in the reference's order.
//   fc1 bwd    same CTAs as fc1 fwd: dpool = g1 W1 (A = W1 read MN-major, i.e. as W1^T, without a transposed copy) and
//              the slice's weight-gradient partial dW1 += g1^T pooled (both operands MN-major), accumulated over the
//              group's samples in tensor memory.
//   conv bwd   same CTAs as conv fwd: the pooled gradient is routed to its arg-max position, multiplied by the activation
//              derivative recomputed from the pooled value, and correlated with the image by the forward's loop nest with the
//              roles swapped (gradient pair x broadcast image value -> 25 tap-pair accumulators per thread, then an ordered
//              cross-thread sum).  Untied-bias gradient = the routed gradient summed over the samples.
//   update     ordered sums of the partials (and, data parallel, of the ranks' sums over NVLink peer memory) fused with
//              w <- w (1 + reg l) + g l / B (ALayer.inl:108-121).
//
// tcgen05 operand layouts (verified on B200 by tools/mma_probe.cu): every operand is kept as tiles of [rows][32 fp32]
// (128-byte rows).  Read K-major (MN = row) the tile carries the SWIZZLE_128B pattern (16-byte chunk c of row r at chunk
// c ^ (r & 7)); read MN-major (K = row) it must carry SWIZZLE_128B_BASE32B (32-byte chunk c of row r at c ^ (r & 3)) - the
// only MN-major layout 32-bit operands have.  The copies (cp.async, 16 bytes) write whichever pattern the use needs.
#include "mega.h"

#include <algorithm>
#include <stdexcept>

#include "fused.cuh"

namespace dlb {
namespace mega {

constexpr int kThreads = 384;
constexpr int kWarps = kThreads / 32;
constexpr int IH = 28, IW = 28, OHW = 24, PHW = 12, NPOS = 144, NPIX = 576, IMG = 784, KK = 25;
constexpr int kTasks = 18;                    // 6 x 3 blocks of 4 x 8 convolution outputs per sample
constexpr int kSubMax = kThreads / kTasks;    // samples per pass of the convolution phases (21)
constexpr int kRX = 40;                       // rows (samples) of a chunk tile of the NLayer phases
constexpr int kNB = 48;                       // MMA N over a chunk (multiple of 16 >= kRX; the extra rows are never stored)
constexpr int kKT = 5;                        // 32-column tiles of one filter's K-slice (160 >= 144)
constexpr int kNT = 4;                        // 32-column tiles of the hidden layer's outputs (128)
constexpr int kTXB = kRX * 128;               // bytes of a chunk tile
constexpr int kDynSmem = 224 * 1024;          // carved below (+ 1 KiB alignment slack at launch)
// tensor memory columns
constexpr int kColD1 = 0, kColD2 = 64, kColD3 = 256;

struct Smem {
    unsigned char* w_hi; unsigned char* w_lo; unsigned char* r; int tw; int r_bytes;
};

// ---------------------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void* dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(su32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }

// K-major SWIZZLE_128B descriptor: SBO = 1024 B between 8-row groups
__device__ __forceinline__ uint64_t desc_k(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B_BASE32B descriptor: LBO = bytes between 32-element groups along MN (= tile stride), SBO = 512 B
// between groups of 4 rows along K
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo_bytes) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)1 << 61);
}
__device__ __forceinline__ uint32_t idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// the three products of the 3xTF32 scheme into one accumulator
__device__ __forceinline__ void mma3(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t id, uint32_t accumulate) {
    tc::umma_tf32(d, a_lo, b_hi, id, accumulate);
    tc::umma_tf32(d, a_hi, b_lo, id, 1u);
    tc::umma_tf32(d, a_hi, b_hi, id, 1u);
}
// byte offset of 16-byte chunk c16 (0..7) of row r inside a [rows][128 B] tile
__device__ __forceinline__ int off_k(int r, int c16) { return r * 128 + ((c16 ^ (r & 7)) << 4); }
__device__ __forceinline__ int off_mn(int r, int c16) { return r * 128 + ((((c16 >> 1) ^ (r & 3)) << 5) | ((c16 & 1) << 4)); }

// lo = x - (x & 0xFFFFE000) over a region of `n16` 16-byte chunks (the layout does not matter: same offsets in both tiles)
__device__ __forceinline__ void make_lo(const unsigned char* hi, unsigned char* lo, int n16) {
    for (int i = threadIdx.x; i < n16; i += kThreads) {
        const float4 v = *reinterpret_cast<const float4*>(hi + 16 * i);
        float4 l;
        l.x = v.x - tc::tf32_hi(v.x); l.y = v.y - tc::tf32_hi(v.y); l.z = v.z - tc::tf32_hi(v.z); l.w = v.w - tc::tf32_hi(v.w);
        *reinterpret_cast<float4*>(lo + 16 * i) = l;
    }
}

__device__ __forceinline__ void read16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Grid barrier on a monotonic counter: every CTA adds one; a barrier is complete when the counter reaches the next
// multiple of the grid size.  arrive / wait are split so that work that nobody else waits for can sit between them.
__device__ __forceinline__ unsigned long long grid_arrive(unsigned long long* ctr, unsigned grid) {
    __threadfence();
    const unsigned long long old = atomicAdd(ctr, 1ULL);
    return (old / grid + 1ULL) * grid;
}
__device__ __forceinline__ void grid_wait(unsigned long long* ctr, unsigned long long target) {
    long long spin = 0;
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
        if (v >= target) break;
        if (++spin > (1LL << 28)) __trap();   // bounded: a lost CTA becomes a launch failure, not a hang
    }
    __threadfence();
}
__device__ __forceinline__ void grid_sync(unsigned long long* ctr, unsigned grid) {
    __syncthreads();
    if (threadIdx.x == 0) grid_wait(ctr, grid_arrive(ctr, grid));
    __syncthreads();
}

__device__ __forceinline__ int group_begin(int g, int groups, int B) { return (int)((long long)g * B / groups); }

// ---------------------------------------------------------------------------------------------------------------
// Phase 1: CLayer::act + PLayer::act  (CLayer.inl:108-137, PLayer.inl:115-138, Tensor.cpp:673-716)
// ---------------------------------------------------------------------------------------------------------------
template <int ACT, bool IS_MAX>
__device__ __forceinline__ void conv_fwd_epilogue(const f32x2 (&acc)[4][8], const float2* __restrict__ bias2, float* __restrict__ pooled,
                                                  unsigned char* __restrict__ sel, int NK, int f0, int b, int by, int bx) {
#pragma unroll
    for (int fi = 0; fi < 2; ++fi) {
#pragma unroll
        for (int wy = 0; wy < 2; ++wy) {
            float vals[4]; unsigned sb[4];
#pragma unroll
            for (int wx = 0; wx < 4; ++wx) {
                float best = IS_MAX ? -real_limits<float>::max() : real_limits<float>::max();
                unsigned bid = 255u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rr = 2 * wy + (q >> 1), cc = 2 * wx + (q & 1);
                    const float2 bz = bias2[(4 * by + rr) * OHW + 8 * bx + cc];
                    const float z = (fi ? hi_of(acc[rr][cc]) : lo_of(acc[rr][cc])) + (fi ? bz.y : bz.x);
                    const float v = activation_ct<float, ACT>(z);
                    if (IS_MAX ? (best < v) : (best > v)) { best = v; bid = (unsigned)q; }   // strict: the first wins ties
                }
                vals[wx] = best; sb[wx] = bid;
            }
            const size_t o = (size_t)b * NK + (size_t)(f0 + fi) * NPOS + (2 * by + wy) * PHW + 4 * bx;
            *reinterpret_cast<float4*>(pooled + o) = make_float4(vals[0], vals[1], vals[2], vals[3]);
            *reinterpret_cast<uchar4*>(sel + o) = make_uchar4((unsigned char)sb[0], (unsigned char)sb[1], (unsigned char)sb[2], (unsigned char)sb[3]);
        }
    }
}

__device__ __forceinline__ void load_images(const Args& a, const int* gather, float* img, int s0, int ns) {
    for (int i = threadIdx.x; i < ns * (IMG / 4); i += kThreads) {
        const int s = i / (IMG / 4), c = i - s * (IMG / 4);
        const int row = gather ? __ldg(gather + s0 + s) : s0 + s;
        cp16(img + s * IMG + c * 4, a.x + (size_t)row * IMG + c * 4, 16u);
    }
}

__device__ void phase_conv_fwd(const Args& a, const Smem& sm, int cta) {
    const int NFP = a.F / 2;
    if (cta >= NFP * a.groups_c) return;
    const int fp = cta % NFP, g = cta / NFP, f0 = 2 * fp;
    const int gs = group_begin(g, a.groups_c, a.B), ge = group_begin(g + 1, a.groups_c, a.B);
    const int n = ge - gs;
    if (n <= 0) return;
    float* img = reinterpret_cast<float*>(sm.r);                                   // [kSubMax][784]
    float2* bias2 = reinterpret_cast<float2*>(sm.r + kSubMax * IMG * 4);           // [576] (filter f0, filter f0 + 1)
    const int tid = threadIdx.x;
    const int* gather = resolve_gather(a.gather_base, a.cursor);
    const int NK = a.F * NPOS;
    const int n_sub = (n + kSubMax - 1) / kSubMax;
    load_images(a, gather, img, gs, (int)((long long)n / n_sub));   // first pass (passes are balanced: pass i covers [i n / n_sub, (i+1) n / n_sub))
    cp_commit();
    f32x2 w[KK];
    {
        const float* filt = a.params + a.off_k0 + (size_t)f0 * KK;
#pragma unroll
        for (int t = 0; t < KK; ++t) w[t] = pack2(__ldg(filt + t), __ldg(filt + KK + t));
        const float* bias = a.params + a.off_b0 + (size_t)f0 * NPIX;
        for (int i = tid; i < NPIX; i += kThreads) bias2[i] = make_float2(__ldg(bias + i), __ldg(bias + NPIX + i));
    }
    const int s = tid / kTasks, blk = tid - s * kTasks;
    const int by = blk / 3, bx = blk - by * 3;
    for (int sc = 0; sc < n_sub; ++sc) {
        const int s0 = gs + (int)((long long)sc * n / n_sub), s1 = gs + (int)((long long)(sc + 1) * n / n_sub);
        const int ns = s1 - s0;
        if (sc > 0) { load_images(a, gather, img, s0, ns); cp_commit(); }
        cp_wait_all();
        __syncthreads();
        if (s < ns) {
            const float* xs = img + s * IMG + (4 * by) * IW + 8 * bx;
            f32x2 acc[4][8];
#pragma unroll
            for (int dy = 0; dy < 4; ++dy)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[dy][c] = pack2(0.f, 0.f);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float xr[12];
                const float4* p = reinterpret_cast<const float4*>(xs + r * IW);
                const float4 v0 = p[0], v1 = p[1], v2 = p[2];
                xr[0] = v0.x; xr[1] = v0.y; xr[2] = v0.z; xr[3] = v0.w; xr[4] = v1.x; xr[5] = v1.y; xr[6] = v1.z; xr[7] = v1.w;
                xr[8] = v2.x; xr[9] = v2.y; xr[10] = v2.z; xr[11] = v2.w;
#pragma unroll
                for (int dy = 0; dy < 4; ++dy) {
                    const int ky = r - dy;
                    if (ky < 0 || ky > 4) continue;
#pragma unroll
                    for (int kx = 0; kx < 5; ++kx)
#pragma unroll
                        for (int c = 0; c < 8; ++c) acc[dy][c] = ffma2_bcast(w[ky * 5 + kx], xr[c + kx], acc[dy][c]);
                }
            }
            const int b = s0 + s;
#define DLB_MEGA_EPI(ACT)                                                                                         \
            do {                                                                                                   \
                if (a.is_max) conv_fwd_epilogue<ACT, true>(acc, bias2, a.pooled, a.sel, NK, f0, b, by, bx);        \
                else conv_fwd_epilogue<ACT, false>(acc, bias2, a.pooled, a.sel, NK, f0, b, by, bx);                \
            } while (0)
            switch (a.act0) {
            case DLB_ACT_RELU: DLB_MEGA_EPI(DLB_ACT_RELU); break;
            case DLB_ACT_SIGMOID: DLB_MEGA_EPI(DLB_ACT_SIGMOID); break;
            case DLB_ACT_TANH: DLB_MEGA_EPI(DLB_ACT_TANH); break;
            default: DLB_MEGA_EPI(DLB_ACT_LINEAR); break;
            }
#undef DLB_MEGA_EPI
        }
        __syncthreads();   // the next pass overwrites the images
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Copies of the NLayer phases' operands into [rows][32 fp32] tiles (tile t = columns [32 t, 32 t + 32) of the slice).
// Rows outside [0, valid_rows) are zero-filled (cp.async src-size 0), so foreign / missing samples contribute nothing.
// ---------------------------------------------------------------------------------------------------------------
template <bool MN>
__device__ __forceinline__ void copy_tiles(unsigned char* dst, int tile_bytes, int rows, int n_tiles, const float* src, size_t ld, int col0, int col_end,
                                           int valid_rows) {
    const int per_row = n_tiles * 8;
    for (int i = threadIdx.x; i < rows * per_row; i += kThreads) {
        const int r = i / per_row, c = i - r * per_row;
        const int t = c >> 3, c16 = c & 7;
        const int col = col0 + 4 * c;
        const bool ok = r < valid_rows && col < col_end;
        const float* s = ok ? src + (size_t)r * ld + col : src;
        cp16(dst + t * tile_bytes + (MN ? off_mn(r, c16) : off_k(r, c16)), s, ok ? 16u : 0u);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Phase 2: hidden NLayer forward (NLayer.inl:131-142), per-filter partial pre-activations.
//   D1[n][b] = sum_{k in slice} W1[n][k0 + k] * pooled[b][k0 + k]          A = W1 slice (K-major), B = pooled chunk (K-major)
// ---------------------------------------------------------------------------------------------------------------
struct Mbar { uint32_t addr; uint32_t phase; };

__device__ void phase_fc1_fwd(const Args& a, const Smem& sm, int cta, uint32_t tmem, Mbar& bar) {
    if (cta >= a.F * a.groups_f) return;
    const int f = cta % a.F, g = cta / a.F;
    const int gs = group_begin(g, a.groups_f, a.B), ge = group_begin(g + 1, a.groups_f, a.B);
    const int NK = a.F * NPOS, k0 = f * NPOS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int RW = sm.tw / 128;
    // W1 slice, K-major tiles [RW][32] x kKT (rows >= H1 and columns >= 144 read as zero)
    copy_tiles<false>(sm.w_hi, sm.tw, RW, kKT, a.params + a.off_w1 + k0, (size_t)NK, 0, NPOS, a.H1);
    cp_commit();
    unsigned char* p_hi = sm.r; unsigned char* p_lo = sm.r + kKT * kTXB;
    bool w_ready = false;
    for (int b0 = gs; b0 < ge; b0 += kRX) {
        const int nb = min(kRX, ge - b0);
        copy_tiles<false>(p_hi, kTXB, kRX, kKT, a.pooled + (size_t)b0 * NK + k0, (size_t)NK, 0, NPOS, nb);
        cp_commit();
        cp_wait_all();
        __syncthreads();
        if (!w_ready) { make_lo(sm.w_hi, sm.w_lo, kKT * sm.tw / 16); w_ready = true; }
        make_lo(p_hi, p_lo, kKT * kTXB / 16);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t id = idesc(128, kNB, 0, 0);
            const uint32_t wh = su32(sm.w_hi), wl = su32(sm.w_lo), ph = su32(p_hi), pl = su32(p_lo);
#pragma unroll 1
            for (int ks = 0; ks < NPOS / 8; ++ks) {
                const int t = ks >> 2, j = (ks & 3) * 32;
                mma3(tmem + kColD1, desc_k(wh + t * sm.tw + j), desc_k(wl + t * sm.tw + j), desc_k(ph + t * kTXB + j), desc_k(pl + t * kTXB + j), id, ks > 0);
            }
            tc::umma_commit(bar.addr);
        }
        if (warp < 4) {   // epilogue: TMEM lane = hidden unit n, column = sample; z1part[f][b][n] (consecutive lanes -> consecutive n)
            tc::mbar_wait(bar.addr, bar.phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int nrow = warp * 32 + lane;
            float* out = a.z1part + ((size_t)f * a.B + b0) * a.H1 + nrow;
#pragma unroll 1
            for (int c = 0; c < kNB; c += 16) {
                if (c >= nb) break;
                uint32_t v[16];
                read16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kColD1 + c), v);
                if (nrow < a.H1) {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (c + q < nb) out[(size_t)(c + q) * a.H1] = __uint_as_float(v[q]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        bar.phase ^= 1;
        __syncthreads();   // tiles and accumulator are free again
    }
    if (!w_ready) cp_wait_all();
}

// ---------------------------------------------------------------------------------------------------------------
// Phase 3: finish the hidden NLayer (ordered sum of the filter partials, bias, activation), then the head: forward,
// softmax / activation, cost derivative, activation gradient, input gradient times the hidden layer's derivative
// (NLayer.inl:131-171, ActivationFunction.inl:47-63, CostFunction.inl:47-51).  One warp per sample; per-CTA partial sums
// of the head's weight / bias gradient and of the batch cost.
// ---------------------------------------------------------------------------------------------------------------
template <int NMAX>
__device__ void phase_head(const Args& a, const Smem& sm, int cta) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H1 = a.H1, n = a.n_out;
    const int HP = n * (H1 + 1);
    const int spc = (a.B + a.grid - 1) / a.grid;
    const int bs = min(a.B, cta * spc), be = min(a.B, bs + spc);
    const int* gather = resolve_gather(a.gather_base, a.cursor);
    const float* bias1 = a.params + a.off_b1;
    const float* b2 = a.params + a.off_b2;
    const float* W2 = a.params + a.off_w2;
    float dw[NMAX][4];
    float db = 0.f;
    double cost_acc = 0.0;
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) dw[j][q] = 0.f;
    const int mine = lane < n ? lane : 0;
    for (int b = bs + warp; b < be; b += kWarps) {
        float a1[4], d1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int nn = lane + 32 * q;
            float z = 0.f;
            if (nn < H1) {
                for (int f = 0; f < a.F; ++f) z += ldcg(a.z1part + ((size_t)f * a.B + b) * H1 + nn);   // filter order: fixed
                z += __ldg(bias1 + nn);
            }
            a1[q] = nn < H1 ? apply_activation<float>(a.act1, z, d1[q]) : 0.f;
            if (nn >= H1) d1[q] = 0.f;
        }
        float z[NMAX];
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            float s = 0.f;
            if (j < n) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int nn = lane + 32 * q;
                    if (nn < H1) s += __ldg(W2 + (size_t)j * H1 + nn) * a1[q];
                }
                s = warp_sum<float>(s) + __ldg(b2 + j);
            }
            z[j] = s;
        }
        const size_t row = gather ? (size_t)__ldg(gather + b) : (size_t)b;
        const float refl = __ldg(a.refs + row * n + mine);
        // lane j owns output j (lanes >= n mirror output 0 and are discarded); sums over the outputs in ascending order
        float zl = z[0];
#pragma unroll
        for (int j = 1; j < NMAX; ++j)
            if (j < n && mine == j) zl = z[j];
        float yl, auxl, ggl;
        if (a.act2 == DLB_ACT_SOFTMAX) {
            float m = z[0];
#pragma unroll
            for (int j = 1; j < NMAX; ++j)
                if (j < n && m < z[j]) m = z[j];
            auxl = dl_exp<float>(zl - m);
            float soft_sum = 0.f;
#pragma unroll
            for (int j = 0; j < NMAX; ++j)
                if (j < n) soft_sum += __shfl_sync(0xffffffffu, auxl, j);
            yl = auxl * (1.f / soft_sum);
            const float ood = 1.f / soft_sum;                        // ActivationFunctionHelper.cpp:50-58
            ggl = auxl * cost_derivative<float>(a.cost_id, yl, refl) * ood;
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < NMAX; ++j)
                if (j < n) t += __shfl_sync(0xffffffffu, ggl, j);
            ggl = ggl - auxl * (t * ood);
        } else {
            yl = apply_activation<float>(a.act2, zl, auxl);
            ggl = auxl * cost_derivative<float>(a.cost_id, yl, refl);
        }
        if (a.cost_out) {
            const float cl = cost_value<float>(a.cost_id, yl, refl);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < NMAX; ++j)
                if (j < n) s += __shfl_sync(0xffffffffu, cl, j);
            cost_acc += (double)s;
        }
        float gg[NMAX];
#pragma unroll
        for (int j = 0; j < NMAX; ++j) gg[j] = j < n ? __shfl_sync(0xffffffffu, ggl, j) : 0.f;
        if (lane < n) db += ggl;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int nn = lane + 32 * q;
            if (nn < H1) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < NMAX; ++j)
                    if (j < n) { s = __ldg(W2 + (size_t)j * H1 + nn) * gg[j] + s; dw[j][q] += gg[j] * a1[q]; }
                a.g1[(size_t)b * 128 + nn] = d1[q] * s;
            }
        }
    }
    // ordered cross-warp sums -> this CTA's partial
    float* scratch = reinterpret_cast<float*>(sm.r);    // [kWarps][HP]
    double* cscr = reinterpret_cast<double*>(sm.r + (size_t)kWarps * HP * 4 + 16);
    {
        float* mine_s = scratch + (size_t)warp * HP;
        if (lane < n) mine_s[lane] = db;
#pragma unroll
        for (int j = 0; j < NMAX; ++j)
            if (j < n) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int nn = lane + 32 * q;
                    if (nn < H1) mine_s[n + j * H1 + nn] = dw[j][q];
                }
            }
        if (lane == 0) cscr[warp] = cost_acc;
    }
    __syncthreads();
    for (int e = tid; e < HP; e += kThreads) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < kWarps; ++wv) s += scratch[(size_t)wv * HP + e];
        a.headpart[(size_t)cta * HP + e] = s;
    }
    if (tid == 0) {
        double c = 0.0;
        for (int wv = 0; wv < kWarps; ++wv) c += cscr[wv];
        a.costpart[cta] = c;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// Phase 4: hidden NLayer backward (NLayer.inl:144-171) for this CTA's filter slice and sample group.
//   dX:  D2[k][b] = sum_n W1[n][k0 + k] * g1[b][n]          A = W1 slice read MN-major (M = k, K = n), B = g1 chunk (K-major)
//   dW:  D3[n][k] += sum_b g1[b][n] * pooled[b][k0 + k]     A = g1 chunk MN-major (M = n, K = b), B = pooled chunk MN-major
// The input gradient is what the next phase waits for: the CTA arrives at the grid barrier as soon as it is stored and
// computes its weight-gradient partial between arrive and wait.
// ---------------------------------------------------------------------------------------------------------------
__device__ void prefetch_fc1_bwd_weights(const Args& a, const Smem& sm, int cta) {   // W1 slice, MN-major tiles (rows = n); issued right after phase 2
    if (cta >= a.F * a.groups_f) return;
    const int f = cta % a.F;
    const int NK = a.F * NPOS;
    copy_tiles<true>(sm.w_hi, sm.tw, sm.tw / 128, kKT, a.params + a.off_w1 + f * NPOS, (size_t)NK, 0, NPOS, a.H1);
    cp_commit();
}

__device__ unsigned long long phase_fc1_bwd(const Args& a, const Smem& sm, int cta, uint32_t tmem, Mbar& bar, bool arrive_after_dx) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned long long target = 0;
    const bool active = cta < a.F * a.groups_f;
    const int f = active ? cta % a.F : 0, g = active ? cta / a.F : 0;
    const int gs = active ? group_begin(g, a.groups_f, a.B) : 0, ge = active ? group_begin(g + 1, a.groups_f, a.B) : 0;
    const int NK = a.F * NPOS, k0 = f * NPOS;
    unsigned char* g_hi = sm.r; unsigned char* g_lo = sm.r + kNT * kTXB;                    // g1 chunk: K-major for dX, then MN-major for dW
    unsigned char* p_hi = sm.r + 2 * kNT * kTXB; unsigned char* p_lo = p_hi + kKT * kTXB;   // pooled chunk, MN-major
    const int n_chunks = (ge - gs + kRX - 1) / kRX;
    bool w_ready = false;
    float dbias = 0.f;   // threads < 128 of the f == 0 CTAs: column sums of g1 = the hidden layer's bias gradient (NLayer.inl:156-160)
    for (int ci = 0; ci < n_chunks; ++ci) {
        const int b0 = gs + ci * kRX, nb = min(kRX, ge - b0);
        const bool last = ci == n_chunks - 1;
        copy_tiles<false>(g_hi, kTXB, kRX, kNT, a.g1 + (size_t)b0 * 128, 128, 0, 128, nb);
        copy_tiles<true>(p_hi, kTXB, kRX, kKT, a.pooled + (size_t)b0 * NK + k0, (size_t)NK, 0, NPOS, nb);
        cp_commit();
        cp_wait_all();
        __syncthreads();
        if (!w_ready) { make_lo(sm.w_hi, sm.w_lo, kKT * sm.tw / 16); w_ready = true; }
        make_lo(g_hi, g_lo, kNT * kTXB / 16);
        make_lo(p_hi, p_lo, kKT * kTXB / 16);
        if (f == 0 && tid < 128) {   // rows in order
            const int t = tid >> 5, c16 = (tid & 31) >> 2, e = tid & 3;
            for (int r = 0; r < nb; ++r) dbias += *reinterpret_cast<const float*>(g_hi + t * kTXB + off_k(r, c16) + 4 * e);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t id = idesc(128, kNB, 1, 0);
            const uint32_t wh = su32(sm.w_hi), wl = su32(sm.w_lo), gh = su32(g_hi), gl = su32(g_lo);
            const int n_ks = (a.H1 + 7) / 8;
#pragma unroll 1
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll 1
                for (int ks = 0; ks < n_ks; ++ks) {
                    const uint32_t ao = (uint32_t)(4 * mt * sm.tw + ks * 1024), bo = (uint32_t)((ks >> 2) * kTXB + (ks & 3) * 32);
                    mma3(tmem + kColD2 + 64 * mt, desc_mn(wh + ao, sm.tw), desc_mn(wl + ao, sm.tw), desc_k(gh + bo), desc_k(gl + bo), id, ks > 0);
                }
            tc::umma_commit(bar.addr);
        }
        if (warp < 4) {   // dpool[b][k0 + k]: TMEM lane = k, column = sample (consecutive lanes -> consecutive k)
            tc::mbar_wait(bar.addr, bar.phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int mt = 0; mt < 2; ++mt) {
                const int k = mt * 128 + warp * 32 + lane;
                if (mt == 1 && warp > 0) break;
#pragma unroll 1
                for (int c = 0; c < kNB; c += 16) {
                    if (c >= nb) break;
                    uint32_t v[16];
                    read16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kColD2 + 64 * mt + c), v);
                    if (k < NPOS) {
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            if (c + q < nb) a.dpool[(size_t)(b0 + c + q) * NK + k0 + k] = __uint_as_float(v[q]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        bar.phase ^= 1;
        __syncthreads();   // the input gradient of this chunk is stored; the K-major g1 tiles are free
        if (last && arrive_after_dx && tid == 0) target = grid_arrive(a.barrier, (unsigned)a.grid);
        // ---- weight gradient of the chunk
        copy_tiles<true>(g_hi, kTXB, kRX, kNT, a.g1 + (size_t)b0 * 128, 128, 0, 128, nb);
        cp_commit();
        cp_wait_all();
        __syncthreads();
        make_lo(g_hi, g_lo, kNT * kTXB / 16);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t id = idesc(128, kKT * 32, 1, 1);
            const uint32_t gh = su32(g_hi), gl = su32(g_lo), ph = su32(p_hi), pl = su32(p_lo);
#pragma unroll 1
            for (int ks = 0; ks < kRX / 8; ++ks) {
                const uint32_t o = (uint32_t)(ks * 1024);
                mma3(tmem + kColD3, desc_mn(gh + o, kTXB), desc_mn(gl + o, kTXB), desc_mn(ph + o, kTXB), desc_mn(pl + o, kTXB), id, (ci > 0 || ks > 0) ? 1u : 0u);
            }
            tc::umma_commit(bar.addr);
        }
        tc::mbar_wait(bar.addr, bar.phase);   // everybody: the tiles are refilled by the next chunk
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        bar.phase ^= 1;
        __syncthreads();
    }
    if (!w_ready) cp_wait_all();
    if (active) {
        float* out = a.dw1part + (size_t)g * a.H1 * NK;
        if (warp < 4) {   // dW1 partial: TMEM lane = n, columns = k
            const int nrow = warp * 32 + lane;
#pragma unroll 1
            for (int c = 0; c < NPOS; c += 16) {
                uint32_t v[16];
                if (n_chunks > 0) read16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kColD3 + c), v);
                else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = 0u;
                }
                if (nrow < a.H1) {
                    float4* o4 = reinterpret_cast<float4*>(out + (size_t)nrow * NK + k0 + c);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        o4[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (f == 0) a.db1part[(size_t)g * 128 + tid] = dbias;
        }
    }
    __syncthreads();
    if (arrive_after_dx && n_chunks == 0) {   // CTAs without samples (or without a role) still take part in the barrier
        if (tid == 0) target = grid_arrive(a.barrier, (unsigned)a.grid);
    }
    return target;
}

// ---------------------------------------------------------------------------------------------------------------
// Phase 5: PLayer + CLayer backward for a first layer (PLayer.inl:140-168, CLayer.inl:139-197, Tensor.cpp:466-498,721-734).
// ---------------------------------------------------------------------------------------------------------------
__device__ void phase_conv_bwd(const Args& a, const Smem& sm, int cta) {
    const int NFP = a.F / 2;
    if (cta >= NFP * a.groups_c) return;
    const int fp = cta % NFP, g = cta / NFP, f0 = 2 * fp;
    const int gs = group_begin(g, a.groups_c, a.B), ge = group_begin(g + 1, a.groups_c, a.B);
    const int n = ge - gs;
    float* img = reinterpret_cast<float*>(sm.r);                                            // [kSubMax][784]
    float* gv = reinterpret_cast<float*>(sm.r + kSubMax * IMG * 4);                         // [kSubMax][288] routed gradient values
    unsigned char* sb = sm.r + kSubMax * IMG * 4 + kSubMax * 2 * NPOS * 4;                  // [kSubMax][288] selectors
    float (*red)[2 * KK] = reinterpret_cast<float (*)[2 * KK]>(gv);                         // [kWarps][50], after the last pass
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int* gather = resolve_gather(a.gather_base, a.cursor);
    const int NK = a.F * NPOS;
    const int n_sub = n > 0 ? (n + kSubMax - 1) / kSubMax : 0;
    const int s = tid / kTasks, blk = tid - s * kTasks;
    const int by = blk / 3, bx = blk - by * 3;
    f32x2 dk[KK];
#pragma unroll
    for (int t = 0; t < KK; ++t) dk[t] = pack2(0.f, 0.f);
    float dbv[3] = {0.f, 0.f, 0.f};
    for (int sc = 0; sc < n_sub; ++sc) {
        const int s0 = gs + (int)((long long)sc * n / n_sub), s1 = gs + (int)((long long)(sc + 1) * n / n_sub);
        const int ns = s1 - s0;
        load_images(a, gather, img, s0, ns);
        cp_commit();
        for (int i = tid; i < ns * 2 * NPOS; i += kThreads) {   // g = dC/dz at the arg-max position: derivative from the pooled VALUE (bit-identical to the stored trace)
            const int ss = i / (2 * NPOS), wdx = i - ss * (2 * NPOS);
            const size_t o = (size_t)(s0 + ss) * NK + (size_t)f0 * NPOS + wdx;
            gv[i] = deriv_from_value<float>(a.act0, ldcg(a.pooled + o)) * ldcg(a.dpool + o);
            sb[i] = a.sel[o];
        }
        cp_wait_all();
        __syncthreads();
        if (s < ns) {
            // the thread's 4 x 8 positions of the routed gradient, as (filter f0, filter f0 + 1) pairs
            f32x2 G[4][8];
#pragma unroll
            for (int wy = 0; wy < 2; ++wy)
#pragma unroll
                for (int wx = 0; wx < 4; ++wx) {
                    const int wdx = (2 * by + wy) * PHW + 4 * bx + wx;
                    const float g0 = gv[s * 2 * NPOS + wdx], g1 = gv[s * 2 * NPOS + NPOS + wdx];
                    const unsigned q0 = sb[s * 2 * NPOS + wdx], q1 = sb[s * 2 * NPOS + NPOS + wdx];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        G[2 * wy + (q >> 1)][2 * wx + (q & 1)] = pack2(q0 == (unsigned)q ? g0 : 0.f, q1 == (unsigned)q ? g1 : 0.f);
                }
            const float* xs = img + s * IMG + (4 * by) * IW + 8 * bx;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float xr[12];
                const float4* p = reinterpret_cast<const float4*>(xs + r * IW);
                const float4 v0 = p[0], v1 = p[1], v2 = p[2];
                xr[0] = v0.x; xr[1] = v0.y; xr[2] = v0.z; xr[3] = v0.w; xr[4] = v1.x; xr[5] = v1.y; xr[6] = v1.z; xr[7] = v1.w;
                xr[8] = v2.x; xr[9] = v2.y; xr[10] = v2.z; xr[11] = v2.w;
#pragma unroll
                for (int dy = 0; dy < 4; ++dy) {
                    const int ky = r - dy;
                    if (ky < 0 || ky > 4) continue;
#pragma unroll
                    for (int kx = 0; kx < 5; ++kx)
#pragma unroll
                        for (int c = 0; c < 8; ++c) dk[ky * 5 + kx] = ffma2_bcast(G[dy][c], xr[c + kx], dk[ky * 5 + kx]);
                }
            }
        }
        // untied bias gradient: the routed gradient summed over the samples in order (3 of the 2 x 576 positions per thread)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int i = tid + kThreads * j;
            const int fi = i / NPIX, pos = i - fi * NPIX;
            const int oy = pos / OHW, ox = pos - oy * OHW;
            const int wdx = fi * NPOS + (oy >> 1) * PHW + (ox >> 1);
            const unsigned q = (unsigned)((oy & 1) * 2 + (ox & 1));
            float acc = dbv[j];
            for (int ss = 0; ss < ns; ++ss) acc += sb[ss * 2 * NPOS + wdx] == q ? gv[ss * 2 * NPOS + wdx] : 0.f;
            dbv[j] = acc;
        }
        __syncthreads();
    }
    // filter gradient: fixed shuffle tree inside the warp, then the warps in order
#pragma unroll
    for (int t = 0; t < KK; ++t) {
        float lo = lo_of(dk[t]), hi = hi_of(dk[t]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo += __shfl_xor_sync(0xffffffffu, lo, o); hi += __shfl_xor_sync(0xffffffffu, hi, o); }
        if (lane == 0) { red[warp][t] = lo; red[warp][KK + t] = hi; }
    }
    __syncthreads();
    float* out = a.convpart + (size_t)g * ((size_t)a.F * (NPIX + KK));
    if (tid < 2 * KK) {
        float sacc = 0.f;
#pragma unroll
        for (int wv = 0; wv < kWarps; ++wv) sacc += red[wv][tid];
        out[(size_t)a.F * NPIX + (size_t)(f0 + tid / KK) * KK + tid % KK] = sacc;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int i = tid + kThreads * j;
        out[(size_t)f0 * NPIX + i] = dbv[j];
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// Phase 6: ordered sums of the partial gradients fused with the SGD update (Net.inl:279-316, ALayer.inl:97-122).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float partial_sum(const float* base, size_t stride, int count) {
    float s = 0.f;
    for (int i = 0; i < count; ++i) s += ldcg(base + (size_t)i * stride);
    return s;
}

__device__ void phase_update(const Args& a, int cta) {
    const int NK = a.F * NPOS;
    const long long n0 = (long long)a.F * (NPIX + KK);          // layer 0: [bias F x 576 | filters F x 25]
    const long long HP = (long long)a.n_out * (a.H1 + 1);       // head:    [bias | weights]
    const long long gtid = (long long)cta * kThreads + threadIdx.x, gthreads = (long long)a.grid * kThreads;
    if (gtid == 0) {
        if (a.cursor_rw) *a.cursor_rw += a.advance;   // every reader of the cursor passed a grid barrier before this phase
        if (a.cost_out) {
            double c = 0.0;
            for (int i = 0; i < a.grid; ++i) c += a.costpart[i];
            *a.cost_out = c;
        }
    }
    const bool reg = a.nontrivial_reg != 0;
    auto apply = [&](long long e, float gsum, bool regularised) {
        a.grads[e] = gsum;
        if (a.do_update) {
            const float xv = a.params[e];
            a.params[e] = (reg && regularised) ? xv * a.param_scale + gsum * a.grad_scale : xv + gsum * a.grad_scale;
        }
    };
    // hidden NLayer weights, 4 at a time
    const long long w1n4 = (long long)a.H1 * NK / 4;
    for (long long i = gtid; i < w1n4; i += gthreads) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int g = 0; g < a.groups_f; ++g) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(a.dw1part + (size_t)g * a.H1 * NK) + i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        float4* gp = reinterpret_cast<float4*>(a.grads + a.off_w1) + i;
        *gp = s;
        if (a.do_update) {
            float4* pp = reinterpret_cast<float4*>(a.params + a.off_w1) + i;
            float4 p = *pp;
            if (reg) { p.x = p.x * a.param_scale + s.x * a.grad_scale; p.y = p.y * a.param_scale + s.y * a.grad_scale; p.z = p.z * a.param_scale + s.z * a.grad_scale; p.w = p.w * a.param_scale + s.w * a.grad_scale; }
            else { p.x = p.x + s.x * a.grad_scale; p.y = p.y + s.y * a.grad_scale; p.z = p.z + s.z * a.grad_scale; p.w = p.w + s.w * a.grad_scale; }
            *pp = p;
        }
    }
    // everything else, one element per thread (spread from the far end of the grid so that it overlaps the loop above)
    const long long rest = n0 + a.H1 + HP;
    for (long long j = gthreads - 1 - gtid; j < rest; j += gthreads) {
        if (j < n0) apply(a.off_b0 + j, partial_sum(a.convpart + j, (size_t)n0, a.groups_c), j >= (long long)a.F * NPIX);
        else if (j < n0 + a.H1) apply(a.off_b1 + (j - n0), partial_sum(a.db1part + (j - n0), 128, a.groups_f), false);
        else { const long long e = j - n0 - a.H1; apply(a.off_b2 + e, partial_sum(a.headpart + e, (size_t)HP, a.grid), e >= a.n_out); }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) lenet_step_kernel(const Args a) {
    extern __shared__ unsigned char mega_smem_raw[];
    __shared__ uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;
    unsigned char* base = mega_smem_raw + ((1024u - (su32(mega_smem_raw) & 1023u)) & 1023u);
    Smem sm;
    sm.tw = ((a.H1 + 7) / 8) * 8 * 128;   // bytes of a W1 tile: [round8(H1)][128 B]
    sm.w_hi = base; sm.w_lo = base + kKT * sm.tw; sm.r = base + 2 * kKT * sm.tw; sm.r_bytes = kDynSmem - 2 * kKT * sm.tw;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int cta = blockIdx.x;
    const int mask = a.phase_mask;
    const bool use_tc = (mask & (kFc1Fwd | kFc1Bwd)) != 0;
    if (use_tc) {
        if (tid == 0) {
            tc::mbar_init(su32(&mma_bar), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(su32(&tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const uint32_t tmem = use_tc ? tmem_slot : 0u;
    Mbar bar{su32(&mma_bar), 0u};
    bool first = true;
    auto between = [&]() { if (!first) grid_sync(a.barrier, (unsigned)a.grid); first = false; };

    if (mask & kConvFwd) { between(); phase_conv_fwd(a, sm, cta); }
    if (mask & kFc1Fwd) {
        between();
        phase_fc1_fwd(a, sm, cta, tmem, bar);
        if (mask & kFc1Bwd) prefetch_fc1_bwd_weights(a, sm, cta);   // lands while the head runs
    }
    if (mask & kHead) {
        between();
        if (a.n_out <= 10) phase_head<10>(a, sm, cta); else phase_head<16>(a, sm, cta);
    }
    unsigned long long pending = 0;
    bool have_pending = false;
    if (mask & kFc1Bwd) {
        between();
        if (!(mask & kFc1Fwd)) prefetch_fc1_bwd_weights(a, sm, cta);
        const bool split = (mask & kConvBwd) != 0;   // the convolution backward only waits for the input gradient
        pending = phase_fc1_bwd(a, sm, cta, tmem, bar, split);
        have_pending = split;
    }
    auto between_after_fc1_bwd = [&]() {
        if (have_pending) {   // this CTA arrived inside phase 4; only the wait is left
            if (tid == 0) grid_wait(a.barrier, pending);
            __syncthreads();
            have_pending = false; first = false;
        } else between();
    };
    if (mask & kConvBwd) { between_after_fc1_bwd(); phase_conv_bwd(a, sm, cta); }
    if (mask & kUpdate) { between_after_fc1_bwd(); phase_update(a, cta); }

    if (use_tc) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------------
Layout make_layout(const Shape& s, int cap, int n_sms) {
    Layout l{};
    const int NK = s.F * NPOS;
    l.groups_c = std::max(1, n_sms / (s.F / 2));
    l.groups_f = std::max(1, n_sms / s.F);
    l.grid = std::max((s.F / 2) * l.groups_c, s.F * l.groups_f);
    l.pooled = (size_t)cap * NK; l.sel = (size_t)cap * NK; l.dpool = (size_t)cap * NK;
    l.z1part = (size_t)s.F * cap * s.H1;
    l.g1 = (size_t)cap * 128 + 128 * 64;   // slack: chunk copies read whole 16-byte columns up to column 127 of the last row only
    l.dw1part = (size_t)l.groups_f * s.H1 * NK;
    l.db1part = (size_t)l.groups_f * 128;
    l.convpart = (size_t)l.groups_c * s.F * (NPIX + KK);
    l.headpart = (size_t)l.grid * s.n_out * (s.H1 + 1);
    l.costpart = (size_t)l.grid;
    return l;
}

size_t smem_bytes(int) { return (size_t)kDynSmem + 1024; }

int launch(const Args& args, cudaStream_t stream) {
    static bool attr_done = false;
    if (!attr_done) {
        DLB_CUDA(cudaFuncSetAttribute(lenet_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(args.H1)));
        attr_done = true;
    }
    const int mask = args.phase_mask;
    const bool several = (mask & (mask - 1)) != 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)args.grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem_bytes(args.H1); cfg.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeCooperative;
    attr.val.cooperative = several ? 1 : 0;   // co-residency of all CTAs is what makes the grid barriers safe
    cfg.attrs = &attr; cfg.numAttrs = 1;
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (g_timing_on) timing_begin(stream, "lenet_step_kernel");
    DLB_CUDA(cudaLaunchKernelEx(&cfg, lenet_step_kernel, args));
    if (g_timing_on) timing_end(stream);
    return 1;
}

}  // namespace mega
}  // namespace dlb

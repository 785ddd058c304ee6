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
#include "mega_args.h"

#include <algorithm>
#include <stdexcept>

#include "fused.cuh"

namespace dlb {
namespace mega {

constexpr int kThreads = 384;
constexpr int kWarps = kThreads / 32;
constexpr int IW = 28, OHW = 24, PHW = 12, NPOS = 144, NPIX = 576, IMG = 784, KK = 25;
constexpr int kBiasBlk = 33;                  // float2 slots per 4 x 8 output block in the staged untied bias (32 + 1: bank spread)
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
}
__device__ __forceinline__ void grid_sync(unsigned long long* ctr, unsigned grid) {
    __syncthreads();
    if (threadIdx.x == 0) grid_wait(ctr, grid_arrive(ctr, grid));
    __syncthreads();
}

// development: clock64 stamp of CTA thread 0 into slot `slot` (16..31: inside the phases)
__device__ __forceinline__ void pstamp(const Args& a, int cta, int slot) { if (a.prof && threadIdx.x == 0) a.prof[(size_t)cta * 40 + slot] = clock64(); }

__device__ __forceinline__ int group_begin(int g, int groups, int B) { return (int)((unsigned)g * (unsigned)B / (unsigned)groups); }   // g * B < 2^32 (B < 2^24)
// The batch position: the cursor is read ONCE per launch (before anybody advances it); step s of the launch uses cursor + s * advance.
__device__ __forceinline__ const int* gather_at(const Args& a, int cursor0, int step) { return a.gather_base ? a.gather_base + cursor0 + step * a.advance : nullptr; }

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
// Phase 1: CLayer::act + PLayer::act  (CLayer.inl:108-137, PLayer.inl:115-138, Tensor.cpp:673-716)
// ---------------------------------------------------------------------------------------------------------------
// RELU: compile-time ReLU (the benchmark's activation) or the run-time switch; min pooling compares -v (same strict first-wins rule).
// Results go to shared staging [sample][2 filters][144] (values) / (selectors); the CTA then writes them out as whole 1152- / 288-byte rows.
template <bool RELU>
__device__ __forceinline__ void conv_fwd_epilogue(const f32x2 (&acc)[4][8], const float2* __restrict__ bias2, float* __restrict__ pst,
                                                  unsigned char* __restrict__ sst, int s, int by, int bx, int act, float sign) {
#pragma unroll
    for (int fi = 0; fi < 2; ++fi) {
#pragma unroll
        for (int wy = 0; wy < 2; ++wy) {
            float vals[4]; unsigned sb[4];
#pragma unroll
            for (int wx = 0; wx < 4; ++wx) {
                float best = -real_limits<float>::max(), bestv = 0.f;
                unsigned bid = 255u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rr = 2 * wy + (q >> 1), cc = 2 * wx + (q & 1);
                    const float2 bz = bias2[(by * 3 + bx) * kBiasBlk + rr * 8 + cc];
                    const float z = (fi ? hi_of(acc[rr][cc]) : lo_of(acc[rr][cc])) + (fi ? bz.y : bz.x);
                    float v, d;
                    if (RELU) v = z > 0.f ? z : 0.f; else v = apply_activation<float>(act, z, d);
                    const float key = v * sign;
                    if (best < key) { best = key; bestv = v; bid = (unsigned)q; }   // strict: the first wins ties
                }
                vals[wx] = bid == 255u ? (sign > 0.f ? -real_limits<float>::max() : real_limits<float>::max()) : bestv; sb[wx] = bid;
            }
            const int o = (s * 2 + fi) * NPOS + (2 * by + wy) * PHW + 4 * bx;
            *reinterpret_cast<float4*>(pst + o) = make_float4(vals[0], vals[1], vals[2], vals[3]);
            *reinterpret_cast<uchar4*>(sst + o) = make_uchar4((unsigned char)sb[0], (unsigned char)sb[1], (unsigned char)sb[2], (unsigned char)sb[3]);
        }
    }
}

__device__ __forceinline__ void load_images(const Args& a, const int* gather, float* img, int s0, int ns) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int s = warp; s < ns; s += kWarps) {   // a warp copies whole samples, lane = 16-byte chunk (no divisions in the loop)
        const int row = gather ? __ldg(gather + s0 + s) : s0 + s;
        const float* src = a.x + (size_t)row * IMG;
        float* dst = img + s * IMG;
#pragma unroll
        for (int c = lane; c < IMG / 4; c += 32) cp16(dst + c * 4, src + c * 4, 16u);
    }
}

// W1 slice of this CTA's NLayer role as K-major tiles [round8(H1)][32] x kKT (rows >= H1 and columns >= 144 read as zero)
__device__ __forceinline__ void copy_w1_kmajor(const Args& a, const Smem& sm, int cta) {
    if (cta >= a.F * a.groups_f) return;
    const int f = cta % a.F, NK = a.F * NPOS;
    copy_tiles<false>(sm.w_hi, sm.tw, sm.tw / 128, kKT, a.params + a.off_w1 + f * NPOS, (size_t)NK, 0, NPOS, a.H1);
}

// First-pass images of this CTA's convolution role for the NEXT step of the launch; issued before the update phase (which uses no shared memory).
__device__ __noinline__ void prefetch_next_images(const Args& a, const Smem& sm, int cta, const int* gather) {
    const int NFP = a.F / 2;
    if (cta >= NFP * a.groups_c) return;
    const int g = cta / NFP;
    const int gs = group_begin(g, a.groups_c, a.B), ge = group_begin(g + 1, a.groups_c, a.B);
    const int n = ge - gs;
    if (n <= 0) return;
    const int n_sub = (n + kSubMax - 1) / kSubMax;
    load_images(a, gather, reinterpret_cast<float*>(sm.r), gs, group_begin(1, n_sub, n));
    cp_commit();
}

// w1_prefetch: the W1 slice of the NLayer forward was requested just before this phase (one cp.async group younger than the first
// images); it is split into hi / lo at the end of the phase, off the next phase's critical path.
__device__ __noinline__ void phase_conv_fwd(const Args& a, const Smem& sm, int cta, bool w1_prefetch, const int* gather, bool images_prefetched) {
    const int NFP = a.F / 2;
    if (cta >= NFP * a.groups_c) {
        if (w1_prefetch) { cp_wait_all(); __syncthreads(); make_lo(sm.w_hi, sm.w_lo, kKT * sm.tw / 16); }
        return;
    }
    const int fp = cta % NFP, g = cta / NFP, f0 = 2 * fp;
    const int gs = group_begin(g, a.groups_c, a.B), ge = group_begin(g + 1, a.groups_c, a.B);
    const int n = ge - gs;
    if (n <= 0) {
        if (w1_prefetch) { cp_wait_all(); __syncthreads(); make_lo(sm.w_hi, sm.w_lo, kKT * sm.tw / 16); }
        return;
    }
    float* img = reinterpret_cast<float*>(sm.r);                                   // [kSubMax][784]
    float2* bias2 = reinterpret_cast<float2*>(sm.r + kSubMax * IMG * 4);           // [576] (filter f0, filter f0 + 1)
    float* pst = img;                                                              // [kSubMax][2][144] pooled values (staging; the images are dead by then)
    unsigned char* sst = reinterpret_cast<unsigned char*>(pst + kSubMax * 2 * NPOS);     // [kSubMax][2][144] selectors (staging)
    const int tid = threadIdx.x;
    const int NK = a.F * NPOS;
    const int n_sub = (n + kSubMax - 1) / kSubMax;
    if (!images_prefetched) {   // (in a multi-step launch the previous step requested them while its update ran)
        load_images(a, gather, img, gs, group_begin(1, n_sub, n));   // first pass (passes are balanced: pass i covers [i n / n_sub, (i+1) n / n_sub))
        cp_commit();
    }
    if (w1_prefetch) { copy_w1_kmajor(a, sm, cta); cp_commit(); }
    f32x2 w[KK];
    {
        // parameters are read through L2 (ld.cg): they change between the steps of one launch.  The 50 taps of the filter pair are fetched
        // once per CTA (every thread asking L2 for the same 50 words cost ~6 us) and broadcast from shared memory.
        float* wsm = reinterpret_cast<float*>(bias2 + kTasks * kBiasBlk);   // [2][25]
        const float* filt = a.params + a.off_k0 + (size_t)f0 * KK;
        if (tid < 2 * KK) wsm[tid] = ldcg(filt + tid);
        const float* bias = a.params + a.off_b0 + (size_t)f0 * NPIX;
        for (int i = tid; i < NPIX; i += kThreads) {   // staged per 4 x 8 block of outputs, blocks 33 pairs apart: the lanes of a warp (= blocks) hit distinct banks
            const int row = i / OHW, col = i - row * OHW;
            bias2[((row >> 2) * 3 + (col >> 3)) * kBiasBlk + (row & 3) * 8 + (col & 7)] = make_float2(ldcg(bias + i), ldcg(bias + NPIX + i));
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < KK; ++t) w[t] = pack2(wsm[t], wsm[KK + t]);
    }
    const int s = tid / kTasks, blk = tid - s * kTasks;
    const int by = blk / 3, bx = blk - by * 3;
    for (int sc = 0; sc < n_sub; ++sc) {
        const int s0 = gs + group_begin(sc, n_sub, n), s1 = gs + group_begin(sc + 1, n_sub, n);
        const int ns = s1 - s0;
        if (sc > 0) { load_images(a, gather, img, s0, ns); cp_commit(); cp_wait_all(); }
        else if (w1_prefetch) asm volatile("cp.async.wait_group 1;" ::: "memory");   // the images have landed; W1 may still be in flight
        else cp_wait_all();
        pstamp(a, cta, 13);
        __syncthreads();
        pstamp(a, cta, 14);
        f32x2 acc[4][8];
        if (s < ns) {
            const float* xs = img + s * IMG + (4 * by) * IW + 8 * bx;
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
        }
        pstamp(a, cta, 15);
        __syncthreads();   // everybody is done with the images: the staging below takes their place
        if (s < ns) {
            const float sign = a.is_max ? 1.f : -1.f;
            if (a.act0 == DLB_ACT_RELU) conv_fwd_epilogue<true>(acc, bias2, pst, sst, s, by, bx, a.act0, sign);
            else conv_fwd_epilogue<false>(acc, bias2, pst, sst, s, by, bx, a.act0, sign);
        }
        __syncthreads();   // staging complete (and everybody is done with the images)
        {   // whole rows out: per sample 288 floats (both filters are adjacent in [b][f][144]) and 288 selector bytes
            const int warp = tid >> 5, lane = tid & 31;
            for (int ss = warp; ss < ns; ss += kWarps) {
                const size_t o = (size_t)(s0 + ss) * NK + (size_t)f0 * NPOS;
                const float4* ps4 = reinterpret_cast<const float4*>(pst + ss * 2 * NPOS);
                float4* pd4 = reinterpret_cast<float4*>(a.pooled + o);
#pragma unroll
                for (int c = lane; c < 2 * NPOS / 4; c += 32) pd4[c] = ps4[c];
                if (lane < 2 * NPOS / 16) reinterpret_cast<uint4*>(a.sel + o)[lane] = reinterpret_cast<const uint4*>(sst + ss * 2 * NPOS)[lane];
            }
        }
        __syncthreads();   // the next pass overwrites the images and the staging
    }
    if (w1_prefetch) { cp_wait_all(); __syncthreads(); make_lo(sm.w_hi, sm.w_lo, kKT * sm.tw / 16); }
}

// ---------------------------------------------------------------------------------------------------------------
// Phase 2: hidden NLayer forward (NLayer.inl:131-142), per-filter partial pre-activations.
//   D1[n][b] = sum_{k in slice} W1[n][k0 + k] * pooled[b][k0 + k]          A = W1 slice (K-major), B = pooled chunk (K-major)
// ---------------------------------------------------------------------------------------------------------------
struct Mbar { uint32_t addr; uint32_t phase; };

__device__ __noinline__ void phase_fc1_fwd(const Args& a, const Smem& sm, int cta, uint32_t tmem, Mbar& bar, bool w1_prefetched) {
    if (cta >= a.F * a.groups_f) return;
    const int f = cta % a.F, g = cta / a.F;
    const int gs = group_begin(g, a.groups_f, a.B), ge = group_begin(g + 1, a.groups_f, a.B);
    const int NK = a.F * NPOS, k0 = f * NPOS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H1 = a.H1, B = a.B;
    float* __restrict__ z1part = a.z1part;
    const float* __restrict__ pooled = a.pooled;
    if (!w1_prefetched) { copy_w1_kmajor(a, sm, cta); cp_commit(); }
    unsigned char* p_hi = sm.r; unsigned char* p_lo = sm.r + kKT * kTXB;
    bool w_ready = w1_prefetched;
    for (int b0 = gs; b0 < ge; b0 += kRX) {
        const int nb = min(kRX, ge - b0);
        copy_tiles<false>(p_hi, kTXB, kRX, kKT, pooled + (size_t)b0 * NK + k0, (size_t)NK, 0, NPOS, nb);
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
        {   // epilogue, all warps: TMEM lane quadrant = warp & 3 (hidden unit n = lane of the quadrant), 16 of the kNB columns (samples) per
            // warp; z1part[f][b][n]: consecutive lanes -> consecutive n
            tc::mbar_wait(bar.addr, bar.phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int quad = warp & 3, c = (warp >> 2) * 16;
            const int nrow = quad * 32 + lane;
            if (c < nb) {
                uint32_t v[16];
                read16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(kColD1 + c), v);
                if (nrow < H1) {
                    float* out = z1part + ((size_t)f * B + b0 + c) * H1 + nrow;
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (c + q < nb) out[(size_t)q * H1] = __uint_as_float(v[q]);
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
// samples per CTA of the head phase (all CTAs take part: the phase is bound by the latency of one sample's chain of loads)
__device__ __forceinline__ int head_samples_per_cta(const Args& a) { return (a.B + a.grid - 1) / a.grid; }
constexpr int kHeadBlock = 16;   // samples staged at a time

__device__ __noinline__ void phase_head(const Args& a, const Smem& sm, int cta, const int* gather) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int H1 = a.H1, n = a.n_out, F = a.F;
    const int HP = n * (H1 + 1);
    const int spc = head_samples_per_cta(a);
    const int bs = min(a.B, cta * spc), be = min(a.B, bs + spc);
    if (bs >= be) return;   // phase 6 only reads the partials of the CTAs that had samples
    const float* __restrict__ z1part = a.z1part;
    float* __restrict__ g1 = a.g1;
    const size_t fstride = (size_t)a.B * H1;
    const int act1 = a.act1, act2 = a.act2, cost_id = a.cost_id;
    const bool want_cost = a.cost_out != nullptr;
    // shared staging: W2 [n][H1] | b2 [16] | a1 [blk][128] | act' [blk][128] | references [blk][16] | g2 [blk][16] | cost [blk]
    float* W2s = reinterpret_cast<float*>(sm.r);
    float* b2s = W2s + n * H1;
    float* a1s = b2s + 16;
    float* d1s = a1s + kHeadBlock * 128;
    float* refs_s = d1s + kHeadBlock * 128;
    float* g2s = refs_s + kHeadBlock * 16;
    float* cost_s = g2s + kHeadBlock * 16;
    {   // head parameters: every load first (clamped addresses), then the stores
        const float* W2 = a.params + a.off_w2;
        constexpr int WPT = (16 * 104 + kThreads - 1) / kThreads;   // 5
        float wv[WPT];
#pragma unroll
        for (int q = 0; q < WPT; ++q) wv[q] = ldcg(W2 + min(tid + q * kThreads, n * H1 - 1));
        const float bv = ldcg(a.params + a.off_b2 + min(tid, n - 1));
#pragma unroll
        for (int q = 0; q < WPT; ++q)
            if (tid + q * kThreads < n * H1) W2s[tid + q * kThreads] = wv[q];
        if (tid < 16) b2s[tid] = tid < n ? bv : 0.f;
    }
    const float* bias1 = a.params + a.off_b1;
    // this thread's elements of the head's gradient partial [bias n | W n x H1]: e = tid + kThreads * j
    constexpr int EPT = 8;   // n (H1 + 1) <= 16 * 105 = 1680 <= 8 * 384 ... checked on the host
    float acc[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) acc[j] = 0.f;
    double cost_acc = 0.0;
    for (int blk = bs; blk < be; blk += kHeadBlock) {
        const int nsb = min(kHeadBlock, be - blk);
        // ---- the block's operands in ONE round of loads.  References (two dependent loads) first.
        size_t ref_row = 0;
        if (tid < nsb * 16) ref_row = gather ? (size_t)__ldcg(gather + blk + (tid >> 4)) : (size_t)(blk + (tid >> 4));   // its dependent load is issued after the partials below
        // hidden pre-activations: ordered sum of the filter partials, four outputs per thread (16-byte loads: the phase is bound by the
        // number of outstanding requests per SM, not by bytes), all F loads of an item in flight at once
        {
            const int H4 = H1 >> 2;   // H1 % 4 == 0 (checked on the host)
            for (int it0 = 0; it0 < nsb * H4; it0 += kThreads) {
                if (it0 + (tid & ~31) >= nsb * H4) {   // this warp has no item (but may own a reference below)
                    if (it0 == 0 && tid < nsb * 16) { const int j = tid & 15; refs_s[tid] = j < n ? __ldg(a.refs + ref_row * n + j) : 0.f; }
                    continue;
                }
                const int it = min(it0 + tid, nsb * H4 - 1);   // clamped: the loads below are unconditional
                const int s = it / H4, n4 = it - s * H4;
                const float4* src = reinterpret_cast<const float4*>(z1part + ((size_t)blk + s) * H1) + n4;
                const size_t fs4 = fstride >> 2;
                float4 zs = make_float4(0.f, 0.f, 0.f, 0.f);
                constexpr int FB = 20;
                for (int fb = 0; fb < F; fb += FB) {
                    float4 v[FB];
#pragma unroll
                    for (int ff = 0; ff < FB; ++ff) v[ff] = __ldcg(src + (size_t)min(fb + ff, F - 1) * fs4);
#pragma unroll
                    for (int ff = 0; ff < FB; ++ff)
                        if (fb + ff < F) { zs.x += v[ff].x; zs.y += v[ff].y; zs.z += v[ff].z; zs.w += v[ff].w; }   // filter order: fixed
                }
                const float4 bz = __ldcg(reinterpret_cast<const float4*>(bias1) + n4);
                if (it0 == 0 && tid < nsb * 16) { const int j = tid & 15; refs_s[tid] = j < n ? __ldg(a.refs + ref_row * n + j) : 0.f; }
                if (it0 + tid < nsb * H4) {
                    float4 av, dv;
                    av.x = apply_activation<float>(act1, zs.x + bz.x, dv.x); av.y = apply_activation<float>(act1, zs.y + bz.y, dv.y);
                    av.z = apply_activation<float>(act1, zs.z + bz.z, dv.z); av.w = apply_activation<float>(act1, zs.w + bz.w, dv.w);
                    *reinterpret_cast<float4*>(a1s + s * 128 + 4 * n4) = av;
                    *reinterpret_cast<float4*>(d1s + s * 128 + 4 * n4) = dv;
                }
            }
        }
        pstamp(a, cta, 16);
        __syncthreads();
        pstamp(a, cta, 17);
        // ---- head forward, cost derivative and activation gradient: thread (sample s, output j), 16 lanes per sample
        if (tid < nsb * 16) {   // whole 16-lane groups; nsb * 16 is a multiple of 16, so the shuffles below see complete groups
            const int s = tid >> 4, j = tid & 15;
            const bool live = j < n;
            const int jc = live ? j : 0;
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;   // four interleaved partial sums, combined in a fixed order
            const float* wrow = W2s + jc * H1;
            const float* arow = a1s + s * 128;
            int nn = 0;
            for (; nn + 3 < H1; nn += 4) { p0 += wrow[nn] * arow[nn]; p1 += wrow[nn + 1] * arow[nn + 1]; p2 += wrow[nn + 2] * arow[nn + 2]; p3 += wrow[nn + 3] * arow[nn + 3]; }
            for (; nn < H1; ++nn) p0 += wrow[nn] * arow[nn];
            const float z = ((p0 + p1) + (p2 + p3)) + b2s[jc];
            const float refl = refs_s[tid];
            const unsigned gmask = 0xFFFFu << (lane & 16);   // this sample's 16 lanes
            float yl, auxl, ggl;
            if (act2 == DLB_ACT_SOFTMAX) {   // whole-vector softmax with the un-normalised exponent as auxiliary data (ActivationFunction.inl:47-63)
                float m = live ? z : -real_limits<float>::max();
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) { const float t = __shfl_xor_sync(gmask, m, o, 16); m = t > m ? t : m; }
                auxl = live ? dl_exp<float>(z - m) : 0.f;
                float soft_sum = auxl;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) soft_sum += __shfl_xor_sync(gmask, soft_sum, o, 16);
                const float ood = 1.f / soft_sum;                        // ActivationFunctionHelper.cpp:50-58
                yl = auxl * ood;
                ggl = live ? auxl * cost_derivative<float>(cost_id, yl, refl) * ood : 0.f;
                float t = ggl;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(gmask, t, o, 16);
                ggl = ggl - auxl * (t * ood);
            } else {
                yl = apply_activation<float>(act2, z, auxl);
                ggl = live ? auxl * cost_derivative<float>(cost_id, yl, refl) : 0.f;
            }
            g2s[tid] = ggl;
            if (want_cost) {
                float c = live ? cost_value<float>(cost_id, yl, refl) : 0.f;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(gmask, c, o, 16);
                if (j == 0) cost_s[s] = c;
            }
        }
        pstamp(a, cta, 18);
        __syncthreads();
        // ---- input gradient of the head times the hidden layer's derivative: g1[b][nn] (NLayer.inl:166-169, then calc_in_grad of the hidden layer)
        for (int it = tid; it < nsb * H1; it += kThreads) {
            const int s = it / H1, nn = it - s * H1;
            float sx = 0.f;
            for (int j = 0; j < n; ++j) sx = W2s[j * H1 + nn] * g2s[s * 16 + j] + sx;
            g1[(size_t)(blk + s) * 128 + nn] = d1s[s * 128 + nn] * sx;
        }
        // ---- this CTA's partial of the head's gradient: samples in order
#pragma unroll
        for (int jj = 0; jj < EPT; ++jj) {
            const int e = tid + kThreads * jj;
            if (e < HP) {
                float t = acc[jj];
                if (e < n) { for (int s = 0; s < nsb; ++s) t += g2s[s * 16 + e]; }
                else { const int j = (e - n) / H1, nn = (e - n) - j * H1; for (int s = 0; s < nsb; ++s) t += g2s[s * 16 + j] * a1s[s * 128 + nn]; }
                acc[jj] = t;
            }
        }
        if (want_cost && tid == 0) { for (int s = 0; s < nsb; ++s) cost_acc += (double)cost_s[s]; }
        pstamp(a, cta, 19);
        __syncthreads();   // the next block overwrites the staging
    }
#pragma unroll
    for (int jj = 0; jj < EPT; ++jj) {
        const int e = tid + kThreads * jj;
        if (e < HP) a.headpart[(size_t)e * a.grid + cta] = acc[jj];   // [element][CTA]: phase 6 reads an element's partials as one coalesced row
    }
    if (tid == 0) a.costpart[cta] = cost_acc;
}

// ---------------------------------------------------------------------------------------------------------------
// Phase 4: hidden NLayer backward (NLayer.inl:144-171) for this CTA's filter slice and sample group.
//   dX:  D2[k][b] = sum_n W1[n][k0 + k] * g1[b][n]          A = W1 slice read MN-major (M = k, K = n), B = g1 chunk (K-major)
//   dW:  D3[n][k] += sum_b g1[b][n] * pooled[b][k0 + k]     A = g1 chunk MN-major (M = n, K = b), B = pooled chunk MN-major
// The input gradient is what the next phase waits for: the CTA arrives at the grid barrier as soon as it is stored and
// computes its weight-gradient partial between arrive and wait.
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void prefetch_fc1_bwd_weights(const Args& a, const Smem& sm, int cta) {   // W1 slice, MN-major tiles (rows = n); issued right after phase 2
    if (cta >= a.F * a.groups_f) return;
    const int f = cta % a.F;
    const int NK = a.F * NPOS;
    copy_tiles<true>(sm.w_hi, sm.tw, sm.tw / 128, kKT, a.params + a.off_w1 + f * NPOS, (size_t)NK, 0, NPOS, a.H1);
    cp_commit();
}

__device__ __noinline__ unsigned long long phase_fc1_bwd(const Args& a, const Smem& sm, int cta, uint32_t tmem, Mbar& bar, bool arrive_after_dx) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned long long target = 0;
    const bool active = cta < a.F * a.groups_f;
    const int f = active ? cta % a.F : 0, g = active ? cta / a.F : 0;
    const int gs = active ? group_begin(g, a.groups_f, a.B) : 0, ge = active ? group_begin(g + 1, a.groups_f, a.B) : 0;
    const int NK = a.F * NPOS, k0 = f * NPOS;
    const int H1 = a.H1;
    const float* __restrict__ g1 = a.g1;
    const float* __restrict__ pooled = a.pooled;
    float* __restrict__ dpool = a.dpool;
    unsigned char* g_hi = sm.r; unsigned char* g_lo = sm.r + kNT * kTXB;                    // g1 chunk: K-major for dX, then MN-major for dW
    unsigned char* p_hi = sm.r + 2 * kNT * kTXB; unsigned char* p_lo = p_hi + kKT * kTXB;   // pooled chunk, MN-major
    const int n_chunks = (ge - gs + kRX - 1) / kRX;
    bool w_ready = false;
    float dbias = 0.f;   // threads < 128 of the f == 0 CTAs: column sums of g1 = the hidden layer's bias gradient (NLayer.inl:156-160)
    for (int ci = 0; ci < n_chunks; ++ci) {
        const int b0 = gs + ci * kRX, nb = min(kRX, ge - b0);
        const bool last = ci == n_chunks - 1;
        pstamp(a, cta, 20);
        copy_tiles<false>(g_hi, kTXB, kRX, kNT, g1 + (size_t)b0 * 128, 128, 0, 128, nb);
        copy_tiles<true>(p_hi, kTXB, kRX, kKT, pooled + (size_t)b0 * NK + k0, (size_t)NK, 0, NPOS, nb);
        cp_commit();
        cp_wait_all();
        __syncthreads();
        pstamp(a, cta, 21);
        if (!w_ready) { make_lo(sm.w_hi, sm.w_lo, kKT * sm.tw / 16); w_ready = true; }
        make_lo(g_hi, g_lo, kNT * kTXB / 16);
        make_lo(p_hi, p_lo, kKT * kTXB / 16);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            pstamp(a, cta, 22);
            const uint32_t id = idesc(128, kNB, 1, 0);
            const uint32_t wh = su32(sm.w_hi), wl = su32(sm.w_lo), gh = su32(g_hi), gl = su32(g_lo);
            const int n_ks = (H1 + 7) / 8;
#pragma unroll 1
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll 1
                for (int ks = 0; ks < n_ks; ++ks) {
                    const uint32_t ao = (uint32_t)(4 * mt * sm.tw + ks * 1024), bo = (uint32_t)((ks >> 2) * kTXB + (ks & 3) * 32);
                    mma3(tmem + kColD2 + 64 * mt, desc_mn(wh + ao, sm.tw), desc_mn(wl + ao, sm.tw), desc_k(gh + bo), desc_k(gl + bo), id, ks > 0);
                }
            tc::umma_commit(bar.addr);
        }
        {   // all warps: dpool[b][k0 + k]: TMEM lane = k (quadrant warp & 3), 16 of the kNB columns (samples) per warp; consecutive lanes -> consecutive k
            tc::mbar_wait(bar.addr, bar.phase);
            pstamp(a, cta, 23);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int quad = warp & 3, c = (warp >> 2) * 16;
#pragma unroll 1
            for (int mt = 0; mt < 2; ++mt) {
                const int k = mt * 128 + quad * 32 + lane;
                if (mt == 1 && quad > 0) break;      // only k < 144 exists
                if (c < nb) {
                    uint32_t v[16];
                    read16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(kColD2 + 64 * mt + c), v);
                    if (k < NPOS) {
                        float* out = dpool + (size_t)(b0 + c) * NK + k0 + k;
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            if (c + q < nb) out[(size_t)q * NK] = __uint_as_float(v[q]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        bar.phase ^= 1;
        __syncthreads();   // the input gradient of this chunk is stored; the K-major g1 tiles are free
        pstamp(a, cta, 24);
        if (last && arrive_after_dx && tid == 0) target = grid_arrive(a.barrier, (unsigned)a.grid);
        pstamp(a, cta, 25);
        // ---- weight gradient of the chunk.  The MN-major read wants the other swizzle: every 128-byte row of the g1 tiles (hi and lo) is
        // re-ordered in place by one thread (same chunks, SWIZZLE_128B -> SWIZZLE_128B_BASE32B); no second trip to L2.
        if (f == 0 && tid >= 256 && tid < 384) {   // hidden-layer bias gradient = column sums of g1 (NLayer.inl:156-160), rows in order; off the critical path here
            const int nn = tid - 256;
            const int t = nn >> 5, c16 = (nn & 31) >> 2, e = nn & 3;
            for (int r = 0; r < nb; ++r) dbias += *reinterpret_cast<const float*>(g_hi + t * kTXB + off_k(r, c16) + 4 * e);
        }
        __syncthreads();
        for (int i = tid; i < 2 * kNT * kRX; i += kThreads) {
            unsigned char* rowp = g_hi + (size_t)(i / kRX) * kTXB + (i % kRX) * 128;   // tiles of hi then lo are contiguous
            const int r = i % kRX;
            float4 v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(rowp + ((c ^ (r & 7)) << 4));
#pragma unroll
            for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(rowp + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4))) = v[c];
        }
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
        pstamp(a, cta, 26);
        tc::mbar_wait(bar.addr, bar.phase);   // everybody: the tiles are refilled by the next chunk
        pstamp(a, cta, 27);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        bar.phase ^= 1;
        __syncthreads();
    }
    if (!w_ready) cp_wait_all();
    if (active) {
        float* out = a.dw1part + (size_t)g * a.H1 * NK;
        {   // dW1 partial, all warps: TMEM lane = n (quadrant warp & 3), 48 of the 144 columns (k) per warp -> shared staging [128][148] ->
            // whole 576-byte rows of dw1part[g][n][k0 ..] (the direct stores were 32 lines per instruction)
            float* stg = reinterpret_cast<float*>(sm.r);
            constexpr int LDS_ = NPOS + 4;
            const int quad = warp & 3, c0 = (warp >> 2) * 48;
            const int nrow = quad * 32 + lane;
#pragma unroll 1
            for (int c = c0; c < c0 + 48; c += 16) {
                uint32_t v[16];
                if (n_chunks > 0) read16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(kColD3 + c), v);
                else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = 0u;
                }
                float4* o4 = reinterpret_cast<float4*>(stg + nrow * LDS_ + c);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    o4[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            for (int nr = warp; nr < H1; nr += kWarps) {   // a warp writes one row: 36 float4
                const float4* s4 = reinterpret_cast<const float4*>(stg + nr * LDS_);
                float4* d4 = reinterpret_cast<float4*>(out + (size_t)nr * NK + k0);
                d4[lane] = s4[lane];
                if (lane < NPOS / 4 - 32) d4[32 + lane] = s4[32 + lane];
            }
            if (f == 0 && tid >= 256) a.db1part[(size_t)g * 128 + tid - 256] = dbias;
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
__device__ __noinline__ void phase_conv_bwd(const Args& a, const Smem& sm, int cta, const int* gather) {
    const int NFP = a.F / 2;
    if (cta >= NFP * a.groups_c) return;
    const int fp = cta % NFP, g = cta / NFP, f0 = 2 * fp;
    const int gs = group_begin(g, a.groups_c, a.B), ge = group_begin(g + 1, a.groups_c, a.B);
    const int n = ge - gs;
    float* img = reinterpret_cast<float*>(sm.r);                                            // [kSubMax][784]
    float* gv = reinterpret_cast<float*>(sm.r + kSubMax * IMG * 4);                         // [kSubMax][288] pooled gradients, then dC/dz at the arg-max
    unsigned char* sb = sm.r + kSubMax * IMG * 4 + kSubMax * 2 * NPOS * 4;                  // [kSubMax][288] selectors
    float* pv = reinterpret_cast<float*>(sm.w_hi);                                          // [kSubMax][288] pooled values (the W1 region is free in this phase)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NK = a.F * NPOS;
    const int n_sub = n > 0 ? (n + kSubMax - 1) / kSubMax : 0;
    const int s = tid / kTasks, blk = tid - s * kTasks;
    const int by = blk / 3, bx = blk - by * 3;
    f32x2 dk[KK];
#pragma unroll
    for (int t = 0; t < KK; ++t) dk[t] = pack2(0.f, 0.f);
    float dbv[3] = {0.f, 0.f, 0.f};
    for (int sc = 0; sc < n_sub; ++sc) {
        const int s0 = gs + group_begin(sc, n_sub, n), s1 = gs + group_begin(sc + 1, n_sub, n);
        const int ns = s1 - s0;
        load_images(a, gather, img, s0, ns);
        // the windows' pooled values, pooled gradients and selectors (raw, 16-byte copies: every row of the three arrays starts 16-byte aligned);
        // a warp copies whole samples
        for (int ss = warp; ss < ns; ss += kWarps) {
            const size_t o = (size_t)(s0 + ss) * NK + (size_t)f0 * NPOS;
#pragma unroll
            for (int c = lane; c < 2 * NPOS / 4; c += 32) {
                cp16(pv + ss * 2 * NPOS + 4 * c, a.pooled + o + 4 * c, 16u);
                cp16(gv + ss * 2 * NPOS + 4 * c, a.dpool + o + 4 * c, 16u);
            }
            if (lane < 2 * NPOS / 16) cp16(sb + ss * 2 * NPOS + 16 * lane, a.sel + o + 16 * lane, 16u);
        }
        cp_commit();
        pstamp(a, cta, 28);
        cp_wait_all();
        __syncthreads();
        // g = dC/dz at the arg-max position: derivative from the pooled VALUE (bit-identical to the stored trace) times the pooled gradient
        for (int i = tid; i < ns * 2 * NPOS; i += kThreads) gv[i] = deriv_from_value<float>(a.act0, pv[i]) * gv[i];
        __syncthreads();
        pstamp(a, cta, 29);
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
        pstamp(a, cta, 30);
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
        pstamp(a, cta, 31);
        __syncthreads();
    }
    // filter gradient: every thread's 50 sums go to shared memory (the images are dead), then 350 threads add 7 x 49 rows of a column in
    // order and 50 threads the 7 sub-sums (fixed order; was a 250-shuffle tree per warp)
    float* out = a.convpart + (size_t)g * ((size_t)a.F * (NPIX + KK));
    {
        float* dks = reinterpret_cast<float*>(sm.r);   // [kThreads][52]
        constexpr int LD = 2 * KK + 2;
        __syncthreads();
#pragma unroll
        for (int t = 0; t < KK; ++t) { dks[tid * LD + t] = lo_of(dk[t]); dks[tid * LD + KK + t] = hi_of(dk[t]); }
        __syncthreads();
        float* sub = dks + kThreads * LD;              // [7][50]
        if (tid < 7 * 2 * KK) {
            const int part = tid / (2 * KK), col = tid - part * (2 * KK);
            const int r0 = part * 55, r1 = min(kThreads, r0 + 55);
            float sacc = 0.f;
            for (int r = r0; r < r1; ++r) sacc += dks[r * LD + col];
            sub[part * 2 * KK + col] = sacc;
        }
        __syncthreads();
        if (tid < 2 * KK) {
            float sacc = 0.f;
#pragma unroll
            for (int part = 0; part < 7; ++part) sacc += sub[part * 2 * KK + tid];
            out[(size_t)a.F * NPIX + (size_t)(f0 + tid / KK) * KK + tid % KK] = sacc;
        }
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
__device__ __forceinline__ float partial_sum(const float* base, size_t stride, int count) {   // 8 loads in flight, adds in order
    float s = 0.f;
    for (int i0 = 0; i0 < count; i0 += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ldcg(base + (size_t)min(i0 + i, count - 1) * stride);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += i0 + i < count ? v[i] : 0.f;
    }
    return s;
}

__device__ __noinline__ void phase_update(const Args& a, int cta) {
    const int NK = a.F * NPOS;
    const long long n0 = (long long)a.F * (NPIX + KK);          // layer 0: [bias F x 576 | filters F x 25]
    const long long HP = (long long)a.n_out * (a.H1 + 1);       // head:    [bias | weights]
    const long long gtid = (long long)cta * kThreads + threadIdx.x, gthreads = (long long)a.grid * kThreads;
    const int hspc = head_samples_per_cta(a);
    const int n_head = (a.B + hspc - 1) / hspc;                 // CTAs of the head phase that had samples
    if (gtid == 0) {
        if (a.cursor_rw) *a.cursor_rw += a.advance;   // every reader of the cursor passed a grid barrier before this phase
        if (a.cost_out) {
            double c = 0.0;
            for (int i = 0; i < n_head; ++i) c += a.costpart[i];   // (only when a batch cost was asked for: host-buffer steps)
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
    const float* __restrict__ dw1part = a.dw1part;
    const int groups_f = a.groups_f;
    const size_t wn = (size_t)a.H1 * NK;
    for (long long i = gtid; i < w1n4; i += gthreads) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int g0 = 0; g0 < groups_f; g0 += 8) {
            float4 v[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) v[g] = __ldcg(reinterpret_cast<const float4*>(dw1part + (size_t)min(g0 + g, groups_f - 1) * wn) + i);
#pragma unroll
            for (int g = 0; g < 8; ++g)
                if (g0 + g < groups_f) { s.x += v[g].x; s.y += v[g].y; s.z += v[g].z; s.w += v[g].w; }
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
    // conv layer and hidden bias, one element per thread (spread from the far end of the grid so that it overlaps the loop above)
    const long long rest = n0 + a.H1;
    for (long long j = gthreads - 1 - gtid; j < rest; j += gthreads) {
        if (j < n0) apply(a.off_b0 + j, partial_sum(a.convpart + j, (size_t)n0, a.groups_c), j >= (long long)a.F * NPIX);
        else apply(a.off_b1 + (j - n0), partial_sum(a.db1part + (j - n0), 128, a.groups_f), false);
    }
    // head: up to `grid` partials per element - one WARP per element: lane l adds partials l, l + 32, ... in order, then a fixed shuffle tree
    {
        const int lane = threadIdx.x & 31;
        const long long gwarp = gtid >> 5, gwarps = gthreads >> 5;
        for (long long e = gwarps - 1 - gwarp; e < HP; e += gwarps) {
            float v[5];   // grid <= 160 partials
#pragma unroll
            for (int q = 0; q < 5; ++q) v[q] = ldcg(a.headpart + (size_t)e * a.grid + min(lane + 32 * q, n_head - 1));
            float sacc = 0.f;
#pragma unroll
            for (int q = 0; q < 5; ++q) sacc += lane + 32 * q < n_head ? v[q] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (lane == 0) apply(a.off_b2 + e, sacc, e >= a.n_out);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) lenet_step_kernel(const __grid_constant__ Args a) {
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
    int stamp_i = 0;
    auto stamp = [&]() { if (a.prof && tid == 0 && stamp_i < 16) a.prof[(size_t)cta * 40 + stamp_i] = clock64(); ++stamp_i; };
    const bool w1_prefetch = (mask & kConvFwd) && (mask & kFc1Fwd) && !(a.flags & 1);
    const int cursor0 = a.cursor ? __ldcg(a.cursor) : 0;
    bool images_prefetched = false;
    for (int step = 0; step < a.n_steps; ++step) {
        bool first = true;
        stamp_i = 0;
        auto between = [&]() { stamp(); if (!first) grid_sync(a.barrier, (unsigned)a.grid); first = false; stamp(); };
        if (step > 0) grid_sync(a.barrier, (unsigned)a.grid);   // the update of the previous step is complete everywhere

        const int* gather = gather_at(a, cursor0, step);
        if (mask & kConvFwd) { between(); phase_conv_fwd(a, sm, cta, w1_prefetch, gather, images_prefetched); images_prefetched = false; }
        if (mask & kFc1Fwd) {
            between();
            phase_fc1_fwd(a, sm, cta, tmem, bar, w1_prefetch);
            if (mask & kFc1Bwd) prefetch_fc1_bwd_weights(a, sm, cta);   // lands while the head runs
        }
        if (mask & kHead) {
            between();
            phase_head(a, sm, cta, gather);
            if (a.flags & 2) {   // development: the same phase again, instruction cache warm (slots 34 / 35)
                __syncthreads();
                pstamp(a, cta, 34);
                phase_head(a, sm, cta, gather);
                __syncthreads();
                pstamp(a, cta, 35);
            }
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
                stamp();
                if (tid == 0) grid_wait(a.barrier, pending);
                __syncthreads();
                have_pending = false; first = false;
                stamp();
            } else between();
        };
        if (mask & kConvBwd) { between_after_fc1_bwd(); phase_conv_bwd(a, sm, cta, gather); }
        if (mask & kUpdate) {
            between_after_fc1_bwd();
            if ((mask & kConvFwd) && step + 1 < a.n_steps && !(a.flags & 4)) { prefetch_next_images(a, sm, cta, gather_at(a, cursor0, step + 1)); images_prefetched = true; }
            phase_update(a, cta);
        }
        stamp();
    }

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
    static const bool coop = [] { const char* e = std::getenv("DLB_MEGA_COOP"); return !(e && e[0] == '0'); }();   // 0: plain launch (development: launch-latency A/B)
    const bool several = coop && ((mask & (mask - 1)) != 0 || args.n_steps > 1);
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

// ---------------------------------------------------------------------------------------------------------------
// Runner: owns the kernel's scratch and turns a StepIn into launches.
// ---------------------------------------------------------------------------------------------------------------
class Runner {
public:
    Shape shape{};
    int device = 0, n_sms = 148, cap = 0;
    Layout lay{};
    unsigned char* sel = nullptr; float* z1part = nullptr; float* g1 = nullptr; float* dw1part = nullptr; float* db1part = nullptr;
    float* convpart = nullptr; float* headpart = nullptr; double* costpart = nullptr; unsigned long long* barrier = nullptr;
    long long* prof = nullptr;
    bool profile = [] { const char* e = std::getenv("DLB_MEGA_PROF"); return e && e[0] == '1'; }();

    template <class T> static T* alloc(size_t n) { T* p = nullptr; DLB_CUDA(cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T))); return p; }
    void release() {
        for (void* p : {(void*)sel, (void*)z1part, (void*)g1, (void*)dw1part, (void*)db1part, (void*)convpart, (void*)headpart, (void*)costpart, (void*)prof})
            if (p) cudaFree(p);
        sel = nullptr; z1part = g1 = dw1part = db1part = convpart = headpart = nullptr; costpart = nullptr; prof = nullptr;
        cap = 0;
    }
    void ensure(int want_cap, cudaStream_t stream) {
        if (want_cap <= cap) return;
        DLB_CUDA(cudaStreamSynchronize(stream));
        release();
        lay = make_layout(shape, want_cap, n_sms);
        sel = alloc<unsigned char>(lay.sel); z1part = alloc<float>(lay.z1part); g1 = alloc<float>(lay.g1); dw1part = alloc<float>(lay.dw1part);
        db1part = alloc<float>(lay.db1part); convpart = alloc<float>(lay.convpart); headpart = alloc<float>(lay.headpart); costpart = alloc<double>(lay.costpart);
        prof = alloc<long long>((size_t)lay.grid * 40);
        if (!barrier) { barrier = alloc<unsigned long long>(1); DLB_CUDA(cudaMemsetAsync(barrier, 0, sizeof(unsigned long long), stream)); }
        DLB_CUDA(cudaMemsetAsync(g1, 0, lay.g1 * sizeof(float), stream));   // columns >= H1 stay zero for good
        DLB_CUDA(cudaMemsetAsync(prof, 0, (size_t)lay.grid * 40 * sizeof(long long), stream));
        cap = want_cap;
    }
    ~Runner() { release(); if (barrier) cudaFree(barrier); }
};

Runner* runner_create(const Shape& shape, int device) {
    auto* r = new Runner();
    r->shape = shape; r->device = device;
    DLB_CUDA(cudaDeviceGetAttribute(&r->n_sms, cudaDevAttrMultiProcessorCount, device));
    return r;
}
void runner_destroy(Runner* r) { delete r; }

int runner_step(Runner* r, const StepIn& in) {
    r->ensure(in.cap, in.stream);
    Args a{};
    a.x = in.x; a.refs = in.refs; a.gather_base = in.gather_base; a.cursor = in.cursor; a.cursor_rw = in.cursor_rw; a.advance = in.advance; a.B = in.B;
    a.F = r->shape.F; a.H1 = r->shape.H1; a.n_out = r->shape.n_out;
    a.act0 = in.act0; a.is_max = in.is_max; a.act1 = in.act1; a.act2 = in.act2; a.cost_id = in.cost_id;
    a.params = in.params; a.grads = in.grads; a.n_params = in.n_params;
    a.off_b0 = in.off_b0; a.off_k0 = in.off_k0; a.off_b1 = in.off_b1; a.off_w1 = in.off_w1; a.off_b2 = in.off_b2; a.off_w2 = in.off_w2;
    a.pooled = in.pooled; a.sel = r->sel; a.z1part = r->z1part; a.g1 = r->g1; a.dpool = in.dpool;
    a.dw1part = r->dw1part; a.db1part = r->db1part; a.convpart = r->convpart; a.headpart = r->headpart; a.costpart = r->costpart;
    a.cost_out = in.cost_out; a.barrier = r->barrier;
    a.grid = r->lay.grid; a.groups_c = r->lay.groups_c; a.groups_f = r->lay.groups_f;
    const float l_rate = -static_cast<float>(in.lr), reg = static_cast<float>(in.reg_factor);   // update scalars in Real, like the reference (ALayer.inl:108-121)
    a.param_scale = 1.f + reg * l_rate; a.grad_scale = l_rate * (1.f / static_cast<float>(in.batch_items));
    a.nontrivial_reg = reg > 0.f ? 1 : 0; a.do_update = 1;
    a.world = 1; a.rank = 0;
    a.prof = r->profile ? r->prof : nullptr;
    { const char* e = std::getenv("DLB_MEGA_FLAGS"); a.flags = e ? std::atoi(e) : 0; }
    g_tag_layer = -1;
    int launches = 0;
    const int n_steps = std::max(in.n_steps, 1);
    if (in.split_phases || g_timing_on) {
        static const char* names[6] = {"mega_conv_fwd", "mega_fc1_fwd", "mega_head", "mega_fc1_bwd", "mega_conv_bwd", "mega_update"};
        a.n_steps = 1;
        for (int step = 0; step < n_steps; ++step)
            for (int ph = 0; ph < 6; ++ph) { TagScope tag_(names[ph]); a.phase_mask = 1 << ph; launches += launch(a, in.stream); }
    } else {   // all steps in ONE launch: the kernel loops, a grid barrier separates consecutive steps
        TagScope tag_("mega_step");
        a.phase_mask = kAll;
        a.n_steps = n_steps;
        launches += launch(a, in.stream);
    }
    return launches;
}

size_t runner_debug_read(Runner* r, int what, void* dst, size_t bytes, cudaStream_t stream) {
    const void* src = nullptr; size_t have = 0;
    switch (what) {
    case 101: src = r->z1part; have = r->lay.z1part * 4; break;
    case 102: src = r->g1; have = (size_t)r->cap * 128 * 4; break;
    case 103: src = r->dw1part; have = r->lay.dw1part * 4; break;
    case 104: src = r->convpart; have = r->lay.convpart * 4; break;
    case 105: src = r->headpart; have = r->lay.headpart * 4; break;
    case 106: src = r->db1part; have = r->lay.db1part * 4; break;
    case 107: src = r->prof; have = (size_t)r->lay.grid * 40 * 8; break;
    default: return 0;
    }
    if (!src) return 0;
    const size_t n = std::min(have, bytes);
    DLB_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, stream));
    DLB_CUDA(cudaStreamSynchronize(stream));
    return n;
}

}  // namespace mega
}  // namespace dlb

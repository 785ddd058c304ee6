// conv28.cu — the first LeNet stage at large mini-batches (sm_100a, Real = float): CLayer 5x5 over one 28x28 channel
// + untied bias + activation + PLayer 2x2 max|min, forward and backward, on the packed-fp32 pipe (FFMA2).
//
// Why not tensor cores: as an implicit GEMM the stage has K = 25 and N = F = 20.  With the 3xTF32 split the im2col tile
// (128 B per output pixel, hi and lo) has to be written to shared memory and read back three times by tcgen05.mma; that is
// ~144 KB written + ~270 KB read per sample, ~3000 cycles per sample and SM at the measured ~128 B/clk of shared-memory
// operand bandwidth - MORE than the 2250 cycles the same sample costs on the FFMA2 pipe at full rate (288 k fused
// multiply-adds / (64 lanes x 2 per clock)).  The round-1 kernels ran that pipe at 49 % (forward) and 14 % (backward: a
// sparse walk over the arg-max positions, bound by scattered LDS); these two run the DENSE correlation with everything
// but the image in registers:
//
//   one thread = one 4 x 8 block of convolution outputs of one sample for a PAIR of filters.  The pair's 25 taps (forward)
//   or 25 tap-gradient accumulators (backward) are 64-bit register pairs, the image value is the broadcast scalar operand
//   of fma.rn.f32x2: 800 FFMA2 per 24 16-byte shared-memory loads, no index arithmetic inside.
//
// forward   CTA = (subset of the filter pairs, every cols-th pass of 21 samples); images double-buffered by cp.async one
//           pass ahead; bias + activation + 2x2 max|min (strict compare, first wins: Tensor.cpp:700-709) in registers;
//           pooled value and a ONE-BYTE window selector leave as 16- and 4-byte stores (the int32 index map of trace mode
//           is not needed to train: -30 % bytes).
// backward  CTA = (filter pair, sample group); per pass of 19 samples the images and the pair's pooled values, pooled
//           gradients and selectors are staged one pass ahead.  dC/dz at the arg-max = derivative recomputed from the pooled
//           VALUE x pooled gradient (CLayer.inl:157-165), routed to its window position in registers; filter gradient =
//           correlation with the image (Tensor.cpp:466-546), untied-bias gradient = the routed gradient summed over the
//           samples in order.  Sums across threads and CTAs are ordered (no atomics): bit-identical from run to run.
// The dense form multiplies the zeros of the routed gradient with the image too (the reference skips them,
// Tensor.cpp:484-486): identical for finite images.
#include "conv28.h"

#include <algorithm>
#include <cstdlib>
#include <stdexcept>

#include "fused.cuh"

namespace dlb {
namespace conv28 {

constexpr int IW = 28, OHW = 24, PHW = 12, NPOS = 144, NPIX = 576, IMG = 784, KK = 25;
// (A bank-conflict-free swizzled image layout - 32-float rows, chunk c of row y rotated by 3 (y >> 2) + ... - was measured: the 3-way
// conflicts of the 16-byte image loads in this natural layout are NOT what limits these kernels, 137.3 vs 138.5 us forward at 8192.)
constexpr int kTasks = 18;                    // 6 x 3 blocks of 4 x 8 convolution outputs per sample
constexpr int kBiasBlk = 34;                  // float2 slots per block in the staged untied bias (32 + 2: 16-byte aligned blocks, 4 banks apart)
constexpr int kFTBig = 512, kFTSmall = 384;   // forward: threads per CTA (512: 128 registers, 4 warps per scheduler - from mini-batch 2048 up; 384: finer passes)
constexpr int kBT = 352, kBS = 19;            // backward: threads (342 compute), samples per pass
constexpr int kMaxFpItem = 10;                // filter pairs a forward CTA keeps biases and taps for
constexpr int fwd_smem(int threads) { return 2 * (threads / kTasks) * IMG * 4 + kMaxFpItem * kTasks * kBiasBlk * 8 + kMaxFpItem * 2 * KK * 4; }
constexpr int kBwdBuf = kBS * (IMG * 4 + 2 * NPOS * 4 * 2 + 2 * NPOS);   // images | pooled values | pooled gradients | selectors
constexpr int kBwdSmem = 2 * kBwdBuf;
static_assert(kBwdBuf % 16 == 0, "staging buffers must stay 16-byte aligned");
static_assert(kBT * (2 * KK + 2) * 4 + 7 * 2 * KK * 4 <= kBwdSmem, "the final reduction reuses the staging buffers");

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(su32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ int group_begin(int g, int groups, int B) { return (int)((unsigned)g * (unsigned)B / (unsigned)groups); }

// a warp copies whole samples, lane = 16-byte chunk
template <int WARPS>
__device__ __forceinline__ void load_images(const float* __restrict__ x, const int* __restrict__ gather, float* img, int s0, int ns) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int s = warp; s < ns; s += WARPS) {
        const int row = gather ? __ldg(gather + s0 + s) : s0 + s;
        const float* src = x + (size_t)row * IMG;
        float* dst = img + s * IMG;
#pragma unroll
        for (int c = lane; c < IMG / 4; c += 32) cp16(dst + c * 4, src + c * 4);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
struct FwdArgs {
    const float* x; const int* gather_base; const int* cursor; const float* bias; const float* filt; float* pooled; unsigned char* sel;
    int B, F, act, is_max, n_fsplit, fp_per_item, n_pass;
};

// One 2 x 2 window of one filter: activation + strict first-wins extremum (Tensor.cpp:700-709).  RELU_MAX is the benchmark's pair
// (ReLU, max): the comparison chain runs on the rectified values directly; the general form compares sign * value and knows "no candidate".
template <bool RELU_MAX>
__device__ __forceinline__ void pool_window(const float (&z)[4], int act, float sign, float& value, unsigned& sel) {
    if (RELU_MAX) {
        const float v0 = fmaxf(z[0], 0.f), v1 = fmaxf(z[1], 0.f), v2 = fmaxf(z[2], 0.f), v3 = fmaxf(z[3], 0.f);
        float m = v0; unsigned q = 0u;
        if (m < v1) { m = v1; q = 1u; }
        if (m < v2) { m = v2; q = 2u; }
        if (m < v3) { m = v3; q = 3u; }
        value = m; sel = q;
    } else {
        float best = -real_limits<float>::max(), bestv = 0.f;
        unsigned bid = 255u;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            float d; const float v = apply_activation<float>(act, z[qd], d);
            const float key = v * sign;
            if (best < key) { best = key; bestv = v; bid = (unsigned)qd; }
        }
        value = bid == 255u ? (sign > 0.f ? -real_limits<float>::max() : real_limits<float>::max()) : bestv;
        sel = bid;
    }
}

template <bool RELU_MAX, int kFT>
__global__ void __launch_bounds__(kFT, 1) conv28_forward_kernel(const FwdArgs a) {
    constexpr int kFS = kFT / kTasks;   // samples per pass (28 | 21)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* img = reinterpret_cast<float*>(smem_raw);                                         // [2][kFS][784]
    float2* bias2 = reinterpret_cast<float2*>(smem_raw + 2 * kFS * IMG * 4);                 // [pairs][18 blocks][34] (filter 2p, filter 2p + 1)
    float* wsm = reinterpret_cast<float*>(bias2 + kMaxFpItem * kTasks * kBiasBlk);           // [pairs][2][25]
    const int* gather = resolve_gather(a.gather_base, a.cursor);
    const int tid = threadIdx.x;
    const int NFP = a.F / 2;
    const int fsub = blockIdx.x % a.n_fsplit, col = blockIdx.x / a.n_fsplit, cols = gridDim.x / a.n_fsplit;
    const int fp0 = fsub * a.fp_per_item, nfp = min(NFP, fp0 + a.fp_per_item) - fp0;
    if (nfp <= 0 || col >= a.n_pass) return;
    load_images<kFT / 32>(a.x, gather, img, col * kFS, min(kFS, a.B - col * kFS));
    cp_commit();
    {   // the subset's untied biases, staged per 4 x 8 block of outputs (blocks 34 pairs apart: 16-byte aligned, the lanes of a warp hit distinct banks), and taps
        const float* bias = a.bias + (size_t)(2 * fp0) * NPIX;
        for (int i = tid; i < nfp * NPIX; i += kFT) {
            const int p = i / NPIX, pos = i - p * NPIX;
            const int row = pos / OHW, cc = pos - row * OHW;
            bias2[(p * kTasks + (row >> 2) * 3 + (cc >> 3)) * kBiasBlk + (row & 3) * 8 + (cc & 7)] =
                make_float2(__ldg(bias + (size_t)(2 * p) * NPIX + pos), __ldg(bias + (size_t)(2 * p + 1) * NPIX + pos));
        }
        const float* filt = a.filt + (size_t)(2 * fp0) * KK;
        for (int i = tid; i < nfp * 2 * KK; i += kFT) wsm[i] = __ldg(filt + i);
    }
    const int s = tid / kTasks, blk = tid - s * kTasks;
    const int by = blk / 3, bx = blk - by * 3;
    const float sign = a.is_max ? 1.f : -1.f;
    int it = 0;
    for (int pass = col; pass < a.n_pass; pass += cols, ++it) {
        cp_wait_all();
        __syncthreads();   // this pass's images (and, the first time, biases and taps) are visible; everybody left the other buffer
        if (pass + cols < a.n_pass) load_images<kFT / 32>(a.x, gather, img + ((it + 1) & 1) * kFS * IMG, (pass + cols) * kFS, min(kFS, a.B - (pass + cols) * kFS));
        cp_commit();
        const int s0 = pass * kFS, ns = min(kFS, a.B - s0);
        if (s >= ns) continue;
        const float* xs = img + (it & 1) * kFS * IMG + s * IMG + (4 * by) * IW + 8 * bx;
        for (int p = 0; p < nfp; ++p) {
            f32x2 w[KK];
#pragma unroll
            for (int t = 0; t < KK; ++t) w[t] = pack2(wsm[p * 2 * KK + t], wsm[p * 2 * KK + KK + t]);
            f32x2 acc[4][8];
#pragma unroll
            for (int dy = 0; dy < 4; ++dy)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[dy][c] = pack2(0.f, 0.f);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float xr[12];
                const float4* q = reinterpret_cast<const float4*>(xs + r * IW);
                const float4 v0 = q[0], v1 = q[1], v2 = q[2];
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
            // bias last (CLayer.inl:130; same bits as the small-batch kernel), activation, 2 x 2 window
            const float4* bz4 = reinterpret_cast<const float4*>(bias2 + (p * kTasks + blk) * kBiasBlk);   // [4 rows][4 column pairs] of (f0, f1, f0, f1)
            const int f0 = 2 * (fp0 + p);
            const size_t o0 = ((size_t)(s0 + s) * a.F + f0) * NPOS + (2 * by) * PHW + 4 * bx;
#pragma unroll
            for (int wy = 0; wy < 2; ++wy) {
                float vals[2][4]; unsigned sb[2][4];
#pragma unroll
                for (int wx = 0; wx < 4; ++wx) {
                    const float4 bt = bz4[(2 * wy) * 4 + wx], bb = bz4[(2 * wy + 1) * 4 + wx];
                    const f32x2 a00 = acc[2 * wy][2 * wx], a01 = acc[2 * wy][2 * wx + 1], a10 = acc[2 * wy + 1][2 * wx], a11 = acc[2 * wy + 1][2 * wx + 1];
                    const float za[4] = {lo_of(a00) + bt.x, lo_of(a01) + bt.z, lo_of(a10) + bb.x, lo_of(a11) + bb.z};
                    const float zb[4] = {hi_of(a00) + bt.y, hi_of(a01) + bt.w, hi_of(a10) + bb.y, hi_of(a11) + bb.w};
                    pool_window<RELU_MAX>(za, a.act, sign, vals[0][wx], sb[0][wx]);
                    pool_window<RELU_MAX>(zb, a.act, sign, vals[1][wx], sb[1][wx]);
                }
#pragma unroll
                for (int fi = 0; fi < 2; ++fi) {
                    const size_t o = o0 + (size_t)fi * NPOS + wy * PHW;
                    *reinterpret_cast<float4*>(a.pooled + o) = make_float4(vals[fi][0], vals[fi][1], vals[fi][2], vals[fi][3]);
                    *reinterpret_cast<unsigned*>(a.sel + o) = sb[fi][0] | (sb[fi][1] << 8) | (sb[fi][2] << 16) | (sb[fi][3] << 24);
                }
            }
        }
    }
    cp_wait_all();
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
struct BwdArgs {
    const float* x; const int* gather_base; const int* cursor; const float* dpool; const unsigned char* sel; const float* pooled; float* partial;
    int B, F, act, G, n_par, stagger;
};

__global__ void __launch_bounds__(kBT, 1) conv28_backward_kernel(const BwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int* gather = resolve_gather(a.gather_base, a.cursor);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NFP = a.F / 2;
    const int fp = blockIdx.x % NFP, g = blockIdx.x / NFP, f0 = 2 * fp;
    const int gs = group_begin(g, a.G, a.B), ge = group_begin(g + 1, a.G, a.B);
    const int n = ge - gs;
    const int n_pass = (n + kBS - 1) / kBS;
    const int NK = a.F * NPOS;
    constexpr int kOffPv = kBS * IMG * 4, kOffGv = kOffPv + kBS * 2 * NPOS * 4, kOffSb = kOffGv + kBS * 2 * NPOS * 4;
    auto stage = [&](int p, unsigned char* buf) {
        const int s0 = gs + p * kBS, ns = min(kBS, ge - s0);
        load_images<kBT / 32>(a.x, gather, reinterpret_cast<float*>(buf), s0, ns);
        float* pv = reinterpret_cast<float*>(buf + kOffPv);
        float* gv = reinterpret_cast<float*>(buf + kOffGv);
        unsigned char* sb = buf + kOffSb;
        for (int ss = warp; ss < ns; ss += kBT / 32) {   // the pair's windows of a sample: 288 contiguous elements of each array
            const size_t o = (size_t)(s0 + ss) * NK + (size_t)f0 * NPOS;
#pragma unroll
            for (int c = lane; c < 2 * NPOS / 4; c += 32) {
                cp16(pv + ss * 2 * NPOS + 4 * c, a.pooled + o + 4 * c);
                cp16(gv + ss * 2 * NPOS + 4 * c, a.dpool + o + 4 * c);
            }
            if (lane < 2 * NPOS / 16) cp16(sb + ss * 2 * NPOS + 16 * lane, a.sel + o + 16 * lane);
        }
    };
    if (n_pass > 0) stage(0, smem_raw);
    cp_commit();
    const int s = tid / kTasks, blk = tid - s * kTasks;
    const int by = blk / 3, bx = blk - by * 3;
    f32x2 dk[KK];
#pragma unroll
    for (int t = 0; t < KK; ++t) dk[t] = pack2(0.f, 0.f);
    float db[4] = {0.f, 0.f, 0.f, 0.f};   // threads < 288: the untied-bias gradient of the four positions of window tid of the pair
    const bool bias_first = warp >= 4 && a.stagger;
    for (int p = 0; p < n_pass; ++p) {
        cp_wait_all();
        __syncthreads();
        unsigned char* buf = smem_raw + (p & 1) * kBwdBuf;
        if (p + 1 < n_pass) stage(p + 1, smem_raw + ((p + 1) & 1) * kBwdBuf);
        cp_commit();
        const int ns = min(kBS, n - p * kBS);
        const float* img = reinterpret_cast<const float*>(buf);
        const float* pv = reinterpret_cast<const float*>(buf + kOffPv);
        const float* gv = reinterpret_cast<const float*>(buf + kOffGv);
        const unsigned char* sb = buf + kOffSb;
        // untied bias: window tid of the pair, the samples in order.  Warps 4 and up run it BEFORE their multiply-adds, warps 0-3 after:
        // out of the barrier the FMA pipe of a scheduler then serves one warp while the other two are in their latency-bound parts.
        auto bias_sums = [&] {
            if (tid < 2 * NPOS) {
#pragma unroll
                for (int ss = 0; ss < kBS; ++ss) {   // fully unrolled (no loop bookkeeping); samples beyond ns read as "no position"; one predicated add per position
                    const unsigned q = ss < ns ? (unsigned)sb[ss * 2 * NPOS + tid] : 255u;
                    const float gz = deriv_from_value<float>(a.act, pv[ss * 2 * NPOS + tid]) * gv[ss * 2 * NPOS + tid];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        asm("{ .reg .pred p; setp.eq.u32 p, %2, %3; @p add.f32 %0, %0, %1; }" : "+f"(db[k]) : "f"(gz), "r"(q), "r"((unsigned)k));
                }
            }
        };
        if (bias_first) bias_sums();
        if (s < ns) {
            // the thread's 4 x 8 positions of dC/dz as (filter f0, filter f0 + 1) pairs: the pooled gradient times the derivative
            // recomputed from the pooled value, at the window's arg-max position, zero elsewhere
            f32x2 Gd[4][8];
#pragma unroll
            for (int wy = 0; wy < 2; ++wy) {
                const int w0 = s * 2 * NPOS + (2 * by + wy) * PHW + 4 * bx;
                const float4 pa = *reinterpret_cast<const float4*>(pv + w0), pb = *reinterpret_cast<const float4*>(pv + w0 + NPOS);
                const float4 ga = *reinterpret_cast<const float4*>(gv + w0), gb = *reinterpret_cast<const float4*>(gv + w0 + NPOS);
                const uchar4 qa = *reinterpret_cast<const uchar4*>(sb + w0), qb = *reinterpret_cast<const uchar4*>(sb + w0 + NPOS);
                const float pav[4] = {pa.x, pa.y, pa.z, pa.w}, pbv[4] = {pb.x, pb.y, pb.z, pb.w};
                const float gav[4] = {ga.x, ga.y, ga.z, ga.w}, gbv[4] = {gb.x, gb.y, gb.z, gb.w};
                const unsigned qav[4] = {qa.x, qa.y, qa.z, qa.w}, qbv[4] = {qb.x, qb.y, qb.z, qb.w};
#pragma unroll
                for (int wx = 0; wx < 4; ++wx) {
                    const float g0 = deriv_from_value<float>(a.act, pav[wx]) * gav[wx], g1 = deriv_from_value<float>(a.act, pbv[wx]) * gbv[wx];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        Gd[2 * wy + (q >> 1)][2 * wx + (q & 1)] = pack2(qav[wx] == (unsigned)q ? g0 : 0.f, qbv[wx] == (unsigned)q ? g1 : 0.f);
                }
            }
            const float* xs = img + s * IMG + (4 * by) * IW + 8 * bx;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float xr[12];
                const float4* q = reinterpret_cast<const float4*>(xs + r * IW);
                const float4 v0 = q[0], v1 = q[1], v2 = q[2];
                xr[0] = v0.x; xr[1] = v0.y; xr[2] = v0.z; xr[3] = v0.w; xr[4] = v1.x; xr[5] = v1.y; xr[6] = v1.z; xr[7] = v1.w;
                xr[8] = v2.x; xr[9] = v2.y; xr[10] = v2.z; xr[11] = v2.w;
                // column outermost: the up to 20 multiply-adds of one column (4 gradient rows x 5 taps) go to 20 DIFFERENT accumulators, so
                // consecutive FFMA2 are independent (tap-outermost order chained 8 of them through one accumulator)
#pragma unroll
                for (int c = 0; c < 8; ++c)
#pragma unroll
                    for (int dy = 0; dy < 4; ++dy) {
                        const int ky = r - dy;
                        if (ky < 0 || ky > 4) continue;
#pragma unroll
                        for (int kx = 0; kx < 5; ++kx) dk[ky * 5 + kx] = ffma2_bcast(Gd[dy][c], xr[c + kx], dk[ky * 5 + kx]);
                    }
            }
        }
        if (!bias_first) bias_sums();
    }
    cp_wait_all();
    __syncthreads();
    float* out = a.partial + (size_t)g * a.n_par;
    {   // filter gradient: every thread's 50 sums go to shared memory, 350 threads add 7 x ~51 rows of a column in order, 50 threads the 7 sub-sums
        float* dks = reinterpret_cast<float*>(smem_raw);   // [kBT][52]
        constexpr int LD = 2 * KK + 2;
#pragma unroll
        for (int t = 0; t < KK; ++t) { dks[tid * LD + t] = lo_of(dk[t]); dks[tid * LD + KK + t] = hi_of(dk[t]); }
        __syncthreads();
        float* sub = dks + kBT * LD;                        // [7][50]
        constexpr int kRows = (kBT + 6) / 7;
        if (tid < 7 * 2 * KK) {
            const int part = tid / (2 * KK), colx = tid - part * (2 * KK);
            const int r0 = part * kRows, r1 = min(kBT, r0 + kRows);
            float sacc = 0.f;
            for (int r = r0; r < r1; ++r) sacc += dks[r * LD + colx];
            sub[part * 2 * KK + colx] = sacc;
        }
        __syncthreads();
        if (tid < 2 * KK) {
            float sacc = 0.f;
#pragma unroll
            for (int part = 0; part < 7; ++part) sacc += sub[part * 2 * KK + tid];
            out[(size_t)a.F * NPIX + (size_t)(f0 + tid / KK) * KK + tid % KK] = sacc;
        }
    }
    if (tid < 2 * NPOS) {
        const int fi = tid / NPOS, wdx = tid - fi * NPOS;
        const int py = wdx / PHW, px = wdx - py * PHW;
        float* o = out + (size_t)(f0 + fi) * NPIX + (2 * py) * OHW + 2 * px;
        *reinterpret_cast<float2*>(o) = make_float2(db[0], db[1]);
        *reinterpret_cast<float2*>(o + OHW) = make_float2(db[2], db[3]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
bool eligible(int in_c, int in_h, int in_w, int win_h, int win_w, int pad_h, int pad_w, int stride_h, int stride_w, int filters, int activation,
              int pool_h, int pool_w, int pool_type) {
    return in_c == 1 && in_h == 28 && in_w == 28 && win_h == 5 && win_w == 5 && pad_h == 0 && pad_w == 0 && stride_h == 1 && stride_w == 1 &&
           filters >= 2 && filters <= 64 && filters % 2 == 0 && activation != DLB_ACT_SOFTMAX && pool_h == 2 && pool_w == 2 &&
           (pool_type == DLB_POOL_MAX || pool_type == DLB_POOL_MIN);
}

void forward(cudaStream_t st, const Stage& s, const float* x, const int* gather_base, const int* cursor, const float* bias, const float* filt,
             float* pooled, unsigned char* sel, int B) {
    if (B <= 0) return;
    FwdArgs a{};
    a.x = x; a.gather_base = gather_base; a.cursor = cursor; a.bias = bias; a.filt = filt; a.pooled = pooled; a.sel = sel;
    a.B = B; a.F = s.F; a.act = s.act; a.is_max = s.is_max;
    const bool big = B > 1024;
    const int kFS = (big ? kFTBig : kFTSmall) / kTasks;
    a.n_pass = ceil_div(B, kFS);
    const int NFP = s.F / 2;
    // split of the filter pairs over CTAs: the fewest rounds of (passes per CTA column) x (pairs per CTA); ties go to the fewer splits
    long best_cost = -1;
    for (int n = 1; n <= NFP && n <= kNumSMs; ++n) {
        const int fpi = ceil_div(NFP, n);
        if (fpi > kMaxFpItem || (n > 1 && fpi * (n - 1) >= NFP)) continue;   // (skip splits that leave a subset empty)
        const int cols = std::min(a.n_pass, kNumSMs / n);
        const long cost = (long)ceil_div(a.n_pass, cols) * (fpi * 16 + 1);     // (+1: the image staging every subset repeats)
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; a.n_fsplit = n; a.fp_per_item = fpi; }
    }
    if (best_cost < 0) throw std::runtime_error("conv28::forward: no decomposition");
    const int cols = std::min(a.n_pass, kNumSMs / a.n_fsplit);
    TagScope tag_("conv_pool_fwd");
    const dim3 grid((unsigned)(cols * a.n_fsplit));
    const bool relu_max = s.act == DLB_ACT_RELU && s.is_max;
    if (big) {
        if (relu_max) DLB_LAUNCH_SMEM((conv28_forward_kernel<true, kFTBig>), grid, dim3(kFTBig), fwd_smem(kFTBig), st, a);
        else DLB_LAUNCH_SMEM((conv28_forward_kernel<false, kFTBig>), grid, dim3(kFTBig), fwd_smem(kFTBig), st, a);
    } else {
        if (relu_max) DLB_LAUNCH_SMEM((conv28_forward_kernel<true, kFTSmall>), grid, dim3(kFTSmall), fwd_smem(kFTSmall), st, a);
        else DLB_LAUNCH_SMEM((conv28_forward_kernel<false, kFTSmall>), grid, dim3(kFTSmall), fwd_smem(kFTSmall), st, a);
    }
}

int backward(cudaStream_t st, const Stage& s, const float* x, const int* gather_base, const int* cursor, const float* dpool, const unsigned char* sel,
             const float* pooled, float* partial, int max_rows, int B) {
    if (B <= 0) return 0;
    BwdArgs a{};
    a.x = x; a.gather_base = gather_base; a.cursor = cursor; a.dpool = dpool; a.sel = sel; a.pooled = pooled; a.partial = partial;
    a.B = B; a.F = s.F; a.act = s.act; a.n_par = s.F * (NPIX + KK);
    const int NFP = s.F / 2;
    a.G = std::max(1, std::min({max_rows, kNumSMs / NFP, (int)ceil_div(B, kBS)}));
    static const int stagger_env = [] { const char* e = std::getenv("DLB_CONV28_STAGGER"); return e ? std::atoi(e) : 1; }();
    a.stagger = stagger_env;
    TagScope tag_("conv_pool_bwd");
    DLB_LAUNCH_SMEM((conv28_backward_kernel), dim3((unsigned)(a.G * NFP)), dim3(kBT), kBwdSmem, st, a);
    return a.G;
}

}  // namespace conv28
}  // namespace dlb

// tc_gemm.cuh — fp32 contraction  C[i][j] = sum_r A(i, r) * B(j, r)  on the 5th-generation tensor cores (tcgen05 / UMMA,
// sm_100a) at fp32 accuracy: every fp32 operand is split into a TF32 "hi" part and an fp32 remainder "lo"
// (x = hi + lo exactly), and three kind::tf32 MMAs per k-step (lo*hi, hi*lo, hi*hi) accumulate into the same fp32
// accumulator in tensor memory (the classic 3xTF32 scheme; the dropped lo*lo term is O(2^-22) relative).  This is how
// "tensor cores in the reference's Real precision" is met for Real = float: plain TF32 (10-bit mantissa) cannot hold
// the 1e-5 parity bar over K = 2880, and tcgen05 has no IEEE fp32 operand kind.  Real = double stays on the FMA path
// (gemm.cuh): tcgen05 has no fp64 kind.
//
// Structure of one CTA (one 128 x BN output tile, one split of the reduction):
//   warps 0-3  producers: cp.async global -> shared raw fp32 tiles (3-deep ring, ~96 KB in flight per SM; coalesced along
//              the operand's contiguous index, optional row gather, zero fill, optional implicit column of ones), already
//              in the canonical K-major SWIZZLE_128B UMMA layout (row = 32 fp32 = 128 B; 16-byte chunk c of row r stored
//              at chunk c ^ (r & 7)); then raw -> hi/lo split into the 2-stage operand ring; the same warps run the
//              epilogue at the end (tcgen05.ld of their 32 TMEM lanes -> epilogue functor -> global)
//   warp 4     lane 0 issues the MMAs (tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = BN, K = 8 per instruction),
//              tcgen05.commit releases the shared-memory stage / signals the epilogue; the warp owns the TMEM allocation
//   mbarrier ring of kStages operand stages: full[s] (128 producer arrivals) / empty[s] (1 tcgen05.commit arrival).
// Operands that are not K-contiguous in memory (W^T, G^T, X^T of the backward contractions) are transposed on the fly by
// the producers, so all three NLayer contractions (NLayer.inl:131-171) use the same K-major MMA.
#pragma once

#include "gemm.cuh"

namespace dlb {

// How an operand tile is fetched: value(outer, red) =
//   kRedContig   : base[row(outer) * ld + red]      (gather applies to outer)
//   kOuterContig : base[row(red)   * ld + outer]    (gather applies to red); outer == ones_col yields 1
struct TcOperand {
    const float* base;
    long long ld;
    int outer_contig;          // 0: kRedContig, 1: kOuterContig
    const int* gather_base;    // optional row gather (device int32), addressed through *cursor when cursor != nullptr
    const int* cursor;
    int ones_col;              // outer index that reads as 1 (bias folded into the contraction); -1: none
};

#ifndef DLB_CUDA_EMU

namespace tc {

constexpr int kBM = 128, kBN = 128, kBK = 32;
constexpr int kStages = 2;                                 // hi/lo operand stages consumed by the tensor core
constexpr int kRaw = 2;                                    // raw fp32 stages filled by cp.async (the prefetch depth)
constexpr int kTileBytes = kBM * kBK * 4;                  // 16 KiB: one raw, hi or lo tile of either operand
constexpr int kStageBytes = 4 * kTileBytes;                // A_hi, A_lo, B_hi, B_lo
constexpr int kRawBytes = 2 * kTileBytes;                  // A_raw, B_raw
constexpr int kEpiBytes = 4 * 32 * 33 * 4;                 // one padded 32 x 32 fp32 staging block per epilogue warp
constexpr int kSmemBytes = kStages * kStageBytes + kRaw * kRawBytes + kEpiBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int kThreads = 288;                              // 4 producer warps, 1 MMA warp, 4 epilogue warps
constexpr int kAccCols = 2 * kBN;                          // two fp32 accumulators in tensor memory

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
// Bounded spin: a protocol error becomes a trap (launch failure) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (long long spin = 0; !ok; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1LL << 26)) __trap();
    }
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): start address
// >> 4 in bits [0,14), LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B >> 4 in [32,46) (distance between
// 8-row groups), descriptor version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 @ bit 4), A = B = TF32 (2 @ bits 7 and 10), both
// operands K-major (bits 15, 16 = 0), N >> 3 @ bit 17, M >> 4 @ bit 24.
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Byte offset of 16-byte chunk c of tile row `row` in the K-major SWIZZLE_128B layout.
__device__ __forceinline__ int chunk_off(int row, int c) { return row * 128 + ((c ^ (row & 7)) << 4); }

// Producer thread t owns a fixed set of 8 chunks of every operand tile - it both copies and transforms exactly those, so
// the producers need no barrier among themselves (only the 128-arrival `full` mbarrier towards the MMA warp):
//   reduction-contiguous operand: chunks q = t + 128 i  (row q >> 3, chunk q & 7): one 16-byte cp.async each
//   outer-contiguous operand    : all 8 chunks of tile row t, filled by 32 4-byte cp.async (transposed on the fly)
// Out-of-range elements are zero-filled by cp.async's src-size operand (addresses are clamped into the operand).
struct TileOwner {
    long long grow[8];   // reduction-contiguous: stored row of each owned chunk (gather resolved once per CTA)
};

__device__ __forceinline__ void owner_init(const TcOperand& op, int o0, int o_end, int t, TileOwner& w) {
    if (op.outer_contig) return;
    const int* gather = op.gather_base ? op.gather_base + (op.cursor ? *op.cursor : 0) : nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int o = min(o0 + ((t + 128 * i) >> 3), o_end - 1);
        w.grow[i] = gather ? (long long)__ldg(gather + o) : (long long)o;
    }
}

__device__ __forceinline__ void issue_tile(const TcOperand& op, const TileOwner& w, int o0, int o_end, int r0, int r_end, int t,
                                           unsigned char* raw_tile) {
    const uint32_t raw = smem_u32(raw_tile);
    if (!op.outer_contig) {
        const bool vec = ((op.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(op.base) & 15) == 0) && (((r_end - r0) & 3) == 0 || r0 + kBK <= r_end);
        if (vec) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int q = t + 128 * i;
                const int row = q >> 3, c = q & 7;
                const int r = r0 + 4 * c;
                const bool ok = (o0 + row < o_end) && (r < r_end);
                cp_async_16(raw + chunk_off(row, c), op.base + w.grow[i] * op.ld + min(r, r_end - 4), ok ? 16u : 0u);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int q = t + 128 * i;
                const int row = q >> 3, c = q & 7;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = r0 + 4 * c + j;
                    const bool ok = (o0 + row < o_end) && (r < r_end);
                    cp_async_4(raw + chunk_off(row, c) + 4 * j, op.base + w.grow[i] * op.ld + min(r, r_end - 1), ok ? 4u : 0u);
                }
            }
        }
    } else {
        // raw layout of an outer-contiguous operand: plain [32 reduction rows][128 outer] fp32 (512 B per row), copied as 16-byte
        // chunks along the contiguous (outer) index; transform_tile transposes it into the K-major tiles
        const int* gather = op.gather_base ? op.gather_base + (op.cursor ? *op.cursor : 0) : nullptr;
        const int o_lim = min(o_end, (int)op.ld);                   // the column of ones (o == ld) is synthesised in transform_tile
        const bool vec = ((op.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(op.base) & 15) == 0);
        if (vec) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int q = t + 128 * i;
                const int k = q >> 5, oc = (q & 31) * 4;
                const int r = min(r0 + k, r_end - 1);
                const long long grow = gather ? (long long)__ldg(gather + r) : (long long)r;
                const int o = o0 + oc;
                int n_ok = (r0 + k < r_end) ? min(4, o_lim - o) : 0;
                n_ok = n_ok < 0 ? 0 : n_ok;
                cp_async_16(raw + (k * 128 + oc) * 4, op.base + grow * op.ld + min(o, (int)op.ld - 4), (uint32_t)(4 * n_ok));
            }
        } else {
            const int o = o0 + t;
            const int oc = min(min(o, o_end - 1), (int)op.ld - 1);
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int r = min(r0 + k, r_end - 1);
                const long long grow = gather ? (long long)__ldg(gather + r) : (long long)r;
                cp_async_4(raw + (k * 128 + t) * 4, op.base + grow * op.ld + oc, (o < o_lim && r0 + k < r_end) ? 4u : 0u);
            }
        }
    }
}

// raw fp32 chunk -> hi (TF32-representable) and lo (remainder) chunks at the same offsets of the hi / lo tiles
__device__ __forceinline__ void split_store(float4 v, unsigned char* hi_tile, unsigned char* lo_tile, int off) {
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
}

__device__ __forceinline__ void transform_tile(const TcOperand& op, int o0, int o_end, int r0, int r_end, int t,
                                               const unsigned char* raw_tile, unsigned char* hi_tile, unsigned char* lo_tile) {
    if (!op.outer_contig) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int q = t + 128 * i;
            const int off = chunk_off(q >> 3, q & 7);
            split_store(*reinterpret_cast<const float4*>(raw_tile + off), hi_tile, lo_tile, off);
        }
    } else {
        // thread t produces tile row t (outer index o0 + t): column t of the raw [k][o] block (conflict-free: lanes read
        // consecutive words), or the implicit column of ones
        const float* raw = reinterpret_cast<const float*>(raw_tile);
        const bool ones = (o0 + t < o_end) && (o0 + t == op.ones_col);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float4 v;
            if (ones) {
                v = make_float4(r0 + 4 * c < r_end ? 1.f : 0.f, r0 + 4 * c + 1 < r_end ? 1.f : 0.f,
                                r0 + 4 * c + 2 < r_end ? 1.f : 0.f, r0 + 4 * c + 3 < r_end ? 1.f : 0.f);
            } else {
                v = make_float4(raw[(4 * c) * 128 + t], raw[(4 * c + 1) * 128 + t], raw[(4 * c + 2) * 128 + t], raw[(4 * c + 3) * 128 + t]);
            }
            split_store(v, hi_tile, lo_tile, chunk_off(t, c));
        }
    }
}

// Decodes a linear tile index into output-tile coordinates and its slice of the reduction.
struct TileCoord { int i0, j0, kb, ke, nkb, z; };
__device__ __forceinline__ TileCoord decode_tile(int t, int tiles_m, int tiles_n, int K, int k_per_split) {
    TileCoord c;
    const int per_z = tiles_m * tiles_n;
    c.z = t / per_z;
    const int rem = t - c.z * per_z;
    const int mi = rem / tiles_n, ni = rem - mi * tiles_n;
    c.i0 = mi * kBM; c.j0 = ni * kBN;
    c.kb = c.z * k_per_split;
    c.ke = min(K, c.kb + k_per_split);
    c.nkb = (c.ke - c.kb + kBK - 1) / kBK;
    return c;
}

// Persistent kernel: CTA c processes tiles c, c + gridDim.x, ...  Three roles run concurrently on different tiles:
// producers fill the operand ring for tile j+1 while the MMA lane accumulates tile j+1 into one tensor-memory
// accumulator and the epilogue warps drain tile j from the other one.
template <class EP>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(int M, int N, int K, int k_per_split, int tiles_m, int tiles_n, int n_tiles, int n_splits, TcOperand a, TcOperand b, EP ep,
               float* __restrict__ partials) {
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* raw_ring = smem + kStages * kStageBytes;
    float* epi_stage = reinterpret_cast<float*>(raw_ring + kRaw * kRawBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(epi_stage) + kEpiBytes);
    // bars: full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2]
    uint64_t* bar_full = bars; uint64_t* bar_empty = bars + kStages;
    uint64_t* bar_tfull = bars + 2 * kStages; uint64_t* bar_tempty = bars + 2 * kStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 128); mbar_init(smem_u32(&bar_empty[s]), 1); }
        for (int q = 0; q < 2; ++q) { mbar_init(smem_u32(&bar_tfull[q]), 1); mbar_init(smem_u32(&bar_tempty[q]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {   // the MMA warp owns the tensor-memory allocation: two accumulators of kBN fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kAccCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ---------------- producers: one continuous stream of k-blocks over all of this CTA's tiles ----------------
        int nt_tile = blockIdx.x;            // tile of the next k-block to issue
        int nt_k = 0;
        TileCoord nc{};
        TileOwner nwa, nwb;
        auto open_next = [&]() {
            if (nt_tile < n_tiles) {
                nc = decode_tile(nt_tile, tiles_m, tiles_n, K, k_per_split);
                owner_init(a, nc.i0, M, tid, nwa);
                owner_init(b, nc.j0, N, tid, nwb);
            }
        };
        open_next();
        long long issued = 0;
        auto issue_next = [&]() {            // issues one k-block (or nothing at the end of the stream) and commits a group
            if (nt_tile < n_tiles) {
                unsigned char* raw = raw_ring + (int)(issued % kRaw) * kRawBytes;
                const int r0 = nc.kb + nt_k * kBK;
                issue_tile(a, nwa, nc.i0, M, r0, nc.ke, tid, raw);
                issue_tile(b, nwb, nc.j0, N, r0, nc.ke, tid, raw + kTileBytes);
                ++issued;
                if (++nt_k == nc.nkb) { nt_k = 0; nt_tile += gridDim.x; open_next(); }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int p = 0; p < kRaw - 1; ++p) issue_next();
        long long git = 0;                   // k-blocks consumed so far (stage / phase bookkeeping shared with the MMA lane)
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const TileCoord c = decode_tile(t, tiles_m, tiles_n, K, k_per_split);
            for (int it = 0; it < c.nkb; ++it, ++git) {
                cp_async_wait<kRaw - 2>();                          // this thread's copies of the current k-block have landed
                // every producer's copies have landed, and every producer is done reading the raw slot refilled next
                asm volatile("bar.sync 1, 128;" ::: "memory");
                issue_next();
                const int s = (int)(git % kStages);
                const uint32_t phase = (uint32_t)((git / kStages) & 1);
                mbar_wait(smem_u32(&bar_empty[s]), phase ^ 1);      // hi/lo stage free (first pass: returns at once)
                const unsigned char* raw = raw_ring + (int)(git % kRaw) * kRawBytes;
                unsigned char* st = smem + s * kStageBytes;
                const int r0 = c.kb + it * kBK;
                transform_tile(a, c.i0, M, r0, c.ke, tid, raw, st, st + kTileBytes);
                transform_tile(b, c.j0, N, r0, c.ke, tid, raw + kTileBytes, st + 2 * kTileBytes, st + 3 * kTileBytes);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async-proxy (UMMA) reads
                mbar_arrive(smem_u32(&bar_full[s]));
            }
        }
        cp_async_wait<0>();
    } else if (warp == 4) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t idesc = make_idesc(kBM, kBN);
            long long git = 0;
            int j = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++j) {
                const TileCoord c = decode_tile(t, tiles_m, tiles_n, K, k_per_split);
                const int acc = j & 1;
                mbar_wait(smem_u32(&bar_tempty[acc]), (uint32_t)(((j >> 1) & 1) ^ 1));   // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kBN);
                uint32_t accumulate = 0;
                for (int it = 0; it < c.nkb; ++it, ++git) {
                    const int s = (int)(git % kStages);
                    const uint32_t phase = (uint32_t)((git / kStages) & 1);
                    mbar_wait(smem_u32(&bar_full[s]), phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem_u32(smem + s * kStageBytes);
                    const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + kTileBytes);
                    const uint64_t b_hi = make_desc(st + 2 * kTileBytes), b_lo = make_desc(st + 3 * kTileBytes);
#pragma unroll
                    for (int k = 0; k < kBK / 8; ++k) {
                        const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);   // +32 bytes per k-step inside the 128-byte swizzle row
                        umma_tf32(tmem_d, a_lo + adv, b_hi + adv, idesc, accumulate);
                        umma_tf32(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
                        umma_tf32(tmem_d, a_hi + adv, b_hi + adv, idesc, 1u);
                        accumulate = 1u;
                    }
                    umma_commit(smem_u32(&bar_empty[s]));           // frees the stage when these MMAs have read it
                }
                umma_commit(smem_u32(&bar_tfull[acc]));             // accumulator complete -> epilogue
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue warps 5..8: TMEM lane quadrant = warp % 4 ----------------
        const int quad = warp & 3;
        float* tile = epi_stage + (warp - 5) * (32 * 33);
        const bool split = n_splits > 1;
        int j = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++j) {
            const TileCoord c = decode_tile(t, tiles_m, tiles_n, K, k_per_split);
            const int acc = j & 1;
            mbar_wait(smem_u32(&bar_tfull[acc]), (uint32_t)((j >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int n_chunks = min(kBN, N - c.j0 + 31 & ~31) / 32;   // 32-column blocks that hold valid columns
#pragma unroll 1
            for (int cc = 0; cc < kBN / 32; ++cc) {
                const int c0 = cc * 32;
                if (cc < n_chunks) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kBN + c0);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (cc == n_chunks - 1) {   // last read of this accumulator: hand it back to the MMA lane
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
                    }
                    // 32 x 32 block: lane = row in registers -> shared (padded) -> lane = column, so that every warp-wide access
                    // of the epilogue (stores, bias / previous-activation loads) covers 128 contiguous bytes
#pragma unroll
                    for (int q = 0; q < 32; ++q) tile[lane * 33 + q] = __uint_as_float(v[q]);
                    __syncwarp();
                    const int gj = c.j0 + c0 + lane;
                    if (gj < N) {
#pragma unroll 4
                        for (int r = 0; r < 32; ++r) {
                            const int gr = c.i0 + quad * 32 + r;
                            if (gr >= M) break;
                            const float accv = tile[r * 33 + lane];
                            if (split) partials[((size_t)c.z * M + gr) * N + gj] = accv;
                            else ep(gr, gj, accv);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kAccCols) : "memory");
    }
}

}  // namespace tc

// Host launcher.  Falls back to nothing: callers decide (by Real type and shape) whether to use this or gemm<R>.
template <class EP>
void tc_gemm(cudaStream_t st, int M, int N, int K, const TcOperand& a, const TcOperand& b, EP ep, const GemmWorkspace& ws) {
    if (M <= 0 || N <= 0) return;
    const long tiles = (long)ceil_div(M, tc::kBM) * ceil_div(N, tc::kBN);
    const int kblocks = ceil_div(K > 0 ? K : 1, tc::kBK);
    // Persistent CTAs (one per SM): split the reduction only as far as the split tiles still fit in ONE wave, so that no
    // SM gets a second tile while others idle; keep at least three k-blocks per split.
    int splits = 1;
    if (tiles < kNumSMs && kblocks > 3) {
        splits = (int)(kNumSMs / tiles);
        if (splits > kblocks / 3) splits = kblocks / 3;
        const size_t per_split = (size_t)M * N * sizeof(float);
        const size_t max_by_ws = per_split ? ws.bytes / per_split : 0;
        if ((size_t)splits > max_by_ws) splits = (int)max_by_ws;
        if (splits < 1) splits = 1;
    }
    int k_per_split = ceil_div(kblocks, splits) * tc::kBK;
    splits = ceil_div(K > 0 ? K : 1, k_per_split);
    const int tiles_m = ceil_div(M, tc::kBM), tiles_n = ceil_div(N, tc::kBN);
    const int n_tiles = tiles_m * tiles_n * splits;
    const dim3 grid((unsigned)std::min(n_tiles, kNumSMs));
    float* partials = static_cast<float*>(ws.ptr);
    DLB_LAUNCH_SMEM((tc::tc_gemm_kernel<EP>), grid, dim3(tc::kThreads), (size_t)tc::kSmemBytes, st, M, N, K, k_per_split, tiles_m, tiles_n,
                    n_tiles, splits, a, b, ep, partials);
    if (splits > 1) {
        const size_t total = (size_t)M * N;
        DLB_LAUNCH((gemm_finalize_kernel<float, EP>), dim3((unsigned)ceil_div<size_t>(total, 256)), dim3(256), st,
                   (const float*)partials, splits, M, N, ep);
    }
}

#endif  // !DLB_CUDA_EMU

}  // namespace dlb

// tc_gemm.cuh — fp32 contraction  C[i][j] = sum_r A(i, r) * B(j, r)  on the 5th-generation tensor cores (tcgen05 / UMMA,
// sm_100a) at fp32 accuracy: every fp32 operand is split into a TF32 "hi" part and an fp32 remainder "lo"
// (x = hi + lo exactly), and three kind::tf32 MMAs per k-step (lo*hi, hi*lo, hi*hi) accumulate into the same fp32
// accumulator in tensor memory (the classic 3xTF32 scheme; the dropped lo*lo term is O(2^-22) relative).  This is how
// "tensor cores in the reference's Real precision" is met for Real = float: plain TF32 (10-bit mantissa) cannot hold
// the 1e-5 parity bar over K = 2880, and tcgen05 has no IEEE fp32 operand kind.  Real = double stays on the FMA path
// (gemm.cuh): tcgen05 has no fp64 kind.
//
// Structure of one CTA (one 128 x BN output tile, one split of the reduction):
//   warps 0-7  producers (tc_tiles.cuh): cp.async global -> shared raw fp32 tiles (2-deep ring; coalesced along the
//              operand's contiguous index, optional row gather, zero fill, optional implicit column of ones, or the
//              receptive-field index maps of the convolutions), already in the canonical K-major SWIZZLE_128B UMMA layout
//              where the operand is reduction-contiguous; then raw -> hi/lo split into the 2-stage operand ring.  Eight
//              warps because the producers are bound by instruction latency, not bandwidth: with one warp per scheduler
//              (the first version) a k-block took ~2 900 cycles against ~800 of tensor time (ncu, profiles/r1_v8_full_*).
//   warp 8     lane 0 issues the MMAs (tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = mma_n, K = 8 per instruction),
//              tcgen05.commit releases the shared-memory stage / signals the epilogue; the warp owns the TMEM allocation
//   warps 9-12 epilogue: tcgen05.ld of their 32 TMEM lanes -> shared-memory staging -> epilogue functor -> global
//   mbarrier ring of kStages operand stages: full[s] (256 producer arrivals) / empty[s] (1 tcgen05.commit arrival).
// Operands that are not K-contiguous in memory (W^T, G^T, X^T of the backward contractions, im2col rows) are transposed
// on the fly by the producers, so the NLayer contractions (NLayer.inl:131-171) and the implicit-GEMM convolutions
// (CLayer.inl:108-197) all use the same K-major MMA.
#pragma once

#include <cstdlib>
#include <type_traits>

#ifndef DLB_CUDA_EMU
#include <cuda.h>   // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no link against libcuda)
#endif

#include "gemm.cuh"
#include "tc_tiles.cuh"

namespace dlb {

#ifndef DLB_CUDA_EMU

namespace tc {

// Two shared-memory plans (both 192 KiB), chosen at compile time by the operand kinds:
//  * an operand that has to be transposed (kOuterContig) is copied into a 2-deep ring of raw [k][outer] blocks and
//    transformed into a 2-deep ring of stages {A_hi, A_lo, B_hi, B_lo};
//  * when BOTH operands arrive K-major (the NLayer forward and all three convolution contractions) the copy lands directly
//    where the MMA reads its "hi" operand and only "lo" is computed: kind::tf32 ignores the low 13 mantissa bits of its
//    32-bit operands (checked on B200: feeding x instead of x & 0xFFFFE000 gives bit-identical accumulators), so x itself
//    serves as hi.  One slot = {A, A_lo, B, B_lo}; three slots: one being filled, one being transformed, one under the MMAs.
//    A quarter less transform work and shared-memory traffic per k-block, and one more k-block in flight.
constexpr int kStages = 2;                                 // hi/lo operand stages of the transposing plan
constexpr int kRaw = 2;                                    // raw fp32 blocks of the transposing plan
constexpr int kSlots = 3;                                  // slots of the K-major plan
constexpr int kStageBytes = 4 * kTileBytes;                // A_hi, A_lo, B_hi, B_lo  (K-major plan: A, A_lo, B, B_lo)
constexpr int kRawBytes = 2 * kTileBytes;                  // A_raw, B_raw
static_assert(kSlots * kStageBytes == kStages * kStageBytes + kRaw * kRawBytes, "both plans use the same shared-memory size");
template <int KIND> struct KMajor { static constexpr bool value = KIND != kOuterContig; };
constexpr int kEpiBytes = 4 * 32 * 33 * 4;                 // one padded 32 x 32 fp32 staging block per epilogue warp
constexpr int kSmemBytes = kStages * kStageBytes + kRaw * kRawBytes + kEpiBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int kMmaWarp = kProducerWarps;                   // warp index of the MMA issuer
constexpr int kEpiWarp0 = kProducerWarps + 1;              // first of the 4 epilogue warps (kEpiWarp0 % 4 == 1: quadrant = warp & 3)
constexpr int kThreads = (kProducerWarps + 5) * 32;        // producer warps, 1 MMA warp, 4 epilogue warps
constexpr int kAccCols = 2 * kBN;                          // two fp32 accumulators in tensor memory
static_assert(kProducerWarps % 4 == 0, "epilogue warp -> TMEM lane quadrant mapping assumes kEpiWarp0 % 4 == 1");

// Epilogues that set `static constexpr bool kLaneIsRow = true` are called with consecutive lanes on consecutive ROWS of
// the output tile (the convolution layouts [b][f][pixel]: rows are pixels); the default is consecutive columns.
// Epilogues with the optional row-segment protocol (row4_ok / row4): the accumulator registers of a lane ARE 32 consecutive columns of
// its row, so they are stored straight from registers as eight 16-byte segments - no shared-memory transposition, no warp barriers.
template <class EP, class = void> struct HasRow4 { static constexpr bool value = false; };
template <class EP> struct HasRow4<EP, std::void_t<decltype(&EP::row4_ok)>> { static constexpr bool value = true; };
template <class EP, class = void> struct LaneIsRow { static constexpr bool value = false; };
template <class EP> struct LaneIsRow<EP, std::void_t<decltype(EP::kLaneIsRow)>> { static constexpr bool value = EP::kLaneIsRow; };

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

// TMA: one elected thread fetches a whole operand tile (box {32 fp32, rows}) straight into the SWIZZLE_128B layout; the bytes
// are counted on the slot's mbarrier (expect_tx / complete_tx).
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifdef DLB_TC_PROFILE
// Development builds only (tools/dev_build.sh -DDLB_TC_PROFILE, read by tools/tc_prof.py through dlb_debug_tc_profile):
// clock64 sums of the producer phases (thread 0) and of the MMA lane's wait, over all CTAs and launches.
__device__ unsigned long long g_tc_prof[16];
#define TC_PROF_T(var) const long long var = clock64()
#define TC_PROF_ADD(slot, t0, t1) do { if (tid == prof_tid) prof_acc[slot] += (t1) - (t0); } while (0)
#else
#define TC_PROF_T(var)
#define TC_PROF_ADD(slot, t0, t1)
#endif

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

// Host side of the TMA plan.  cuTensorMapEncodeTiled comes from the driver through the runtime (no libcuda link);
// DLB_TC_TMA=0 in the environment keeps every contraction on the cp.async producers (A/B measurements).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q{};
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
inline bool tma_enabled() {
    static const bool on = [] { const char* e = std::getenv("DLB_TC_TMA"); return !(e && e[0] == '0') && encode_tiled() != nullptr; }();
    return on;
}
// value(outer, red) = base[outer * ld + red] as a 2-D tensor {red, outer}; false when the operand does not qualify.
inline bool make_row_map(CUtensorMap& map, const TcOperand& op, int rows, int K, int box_rows) {
    if (op.gather_base || (op.ld & 3) || (reinterpret_cast<uintptr_t>(op.base) & 15) || K <= 0 || rows <= 0) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)op.ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return encode_tiled()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(op.base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Persistent kernel: CTA c processes tiles c, c + gridDim.x, ...  Three roles run concurrently on different tiles:
// producers fill the operand ring for tile j+1 while the MMA lane accumulates tile j+1 into one tensor-memory
// accumulator and the epilogue warps drain tile j from the other one.
// mma_n: the N of the MMA instruction (a multiple of 16; < 128 only when the whole output is one column tile narrower
// than that - the convolutions with 32 / 64 filters or channels - so that neither the tensor pipe nor the producers
// spend time on columns that do not exist).
template <class EP, int KA, int KB>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(int M, int N, int K, int k_per_split, int tiles_m, int tiles_n, int n_tiles, int n_splits, int mma_n, TcOperand a, TcOperand b, EP ep,
               float* __restrict__ partials, int use_tma, const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b) {
    extern __shared__ unsigned char tc_smem_raw[];
    // 1024-byte alignment (SWIZZLE_128B atoms) by an OFFSET into the shared array, not by an integer round trip of the pointer:
    // the compiler then still knows these are shared-memory addresses and emits LDS / STS instead of generic LD / ST
    // (the first version's transform ran on generic accesses: ncu source page, LD.E.128 / ST.E.128 with long-scoreboard stalls).
    unsigned char* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    unsigned char* raw_ring = smem + kStages * kStageBytes;
    float* epi_stage = reinterpret_cast<float*>(raw_ring + kRaw * kRawBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(epi_stage) + kEpiBytes);
    // bars: full[kSlots] | empty[kSlots] | tmem_full[2] | tmem_empty[2]
    uint64_t* bar_full = bars; uint64_t* bar_empty = bars + kSlots;
    uint64_t* bar_tfull = bars + 2 * kSlots; uint64_t* bar_tempty = bars + 2 * kSlots + 2;
    uint64_t* bar_landed = bars + 2 * kSlots + 4;           // [kSlots] TMA plan: the slot's tiles have landed (1 arrival + transaction bytes)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_landed + kSlots);
    static_assert((3 * kSlots + 4) * 8 + 8 <= 256, "barrier block");
    constexpr bool kDirect = KMajor<KA>::value && KMajor<KB>::value;     // the K-major plan
    constexpr int kRing = kDirect ? kSlots : kStages;         // operand stages the MMA lane cycles through

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kSlots; ++s) { mbar_init(smem_u32(&bar_full[s]), kProducers); mbar_init(smem_u32(&bar_empty[s]), 1); }
        for (int q = 0; q < 2; ++q) { mbar_init(smem_u32(&bar_tfull[q]), 1); mbar_init(smem_u32(&bar_tempty[q]), 4); }
        for (int s = 0; s < kSlots; ++s) mbar_init(smem_u32(&bar_landed[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {   // the MMA warp owns the tensor-memory allocation: two accumulators of kBN fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kAccCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above ran beside the previous kernel of the stream (if launched programmatically); its results are visible from here

#ifdef DLB_TC_PROFILE
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int prof_tid = warp < kProducerWarps ? 0 : kMmaWarp * 32;
#endif
    if (warp < kProducerWarps) {
        // ---------------- producers: one continuous stream of k-blocks over all of this CTA's tiles ----------------
        bool tma_done = false;
        if constexpr (KA == kRedContig && KB == kRedContig) {
            if (use_tma) {
                // ---- TMA plan (dense, ungathered, 16-byte aligned operands): thread 0 asks the TMA unit for whole tiles, two
                // k-blocks ahead; nobody issues per-chunk copies (the phase clocks showed 1 200 cycles per k-block inside that
                // issue - per-warp LDGSTS back-pressure), all 256 threads only compute lo.
                const uint32_t tile_bytes = (uint32_t)kTileBytes + (uint32_t)mma_n * 128u;
                int it_tile = blockIdx.x, it_k = 0;
                TileCoord ic{};
                if (it_tile < n_tiles) ic = decode_tile(it_tile, tiles_m, tiles_n, K, k_per_split);
                long long sent = 0;
                auto send_next = [&]() {
                    if (it_tile >= n_tiles) return;
                    const int s = (int)(sent % kSlots);
                    mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)(((sent / kSlots) & 1) ^ 1));   // the MMAs of this slot's previous use are done
                    const uint32_t bar = smem_u32(&bar_landed[s]);
                    const uint32_t dst = smem_u32(smem + s * kStageBytes);
                    const int r0 = ic.kb + it_k * kBK;
                    mbar_expect_tx(bar, tile_bytes);
                    tma_load_2d(dst, &tma_a, r0, ic.i0, bar);
                    tma_load_2d(dst + 2 * kTileBytes, &tma_b, r0, ic.j0, bar);
                    ++sent;
                    if (++it_k == ic.nkb) { it_k = 0; it_tile += gridDim.x; if (it_tile < n_tiles) ic = decode_tile(it_tile, tiles_m, tiles_n, K, k_per_split); }
                };
                if (tid == 0) { send_next(); send_next(); }
                long long git = 0;
                for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                    const TileCoord c = decode_tile(t, tiles_m, tiles_n, K, k_per_split);
                    for (int it = 0; it < c.nkb; ++it, ++git) {
                        const int s = (int)(git % kSlots);
                        unsigned char* st = smem + s * kStageBytes;
                        TC_PROF_T(p0);
                        mbar_wait(smem_u32(&bar_landed[s]), (uint32_t)((git / kSlots) & 1));
                        TC_PROF_T(p1);
                        lo_tile<KA>(tid, kBM, st, st + kTileBytes);
                        lo_tile<KB>(tid, mma_n, st + 2 * kTileBytes, st + 3 * kTileBytes);
                        TC_PROF_T(p2);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // lo stores -> async-proxy (UMMA) reads
                        mbar_arrive(smem_u32(&bar_full[s]));
                        TC_PROF_T(p3);
                        if (tid == 0) send_next();                      // k-block git + 2
                        TC_PROF_T(p4);
                        TC_PROF_ADD(0, p0, p1); TC_PROF_ADD(4, p1, p2); TC_PROF_ADD(5, p2, p3); TC_PROF_ADD(2, p3, p4); TC_PROF_ADD(6, 0, 1);
                    }
                }
                tma_done = true;
            }
        }
        if (!tma_done) {
        int nt_tile = blockIdx.x;            // tile of the next k-block to issue
        int nt_k = 0;
        TileCoord nc{};
        TileOwner<KA> nwa; TileOwner<KB> nwb;
        auto open_next = [&]() {
            if (nt_tile < n_tiles) {
                nc = decode_tile(nt_tile, tiles_m, tiles_n, K, k_per_split);
                owner_init<KA>(a, nc.i0, M, tid, nwa);
                owner_init<KB>(b, nc.j0, N, tid, nwb);
            }
        };
        open_next();
        long long issued = 0;
        auto issue_next = [&]() {            // issues one k-block (or nothing at the end of the stream) and commits a group
            if (nt_tile < n_tiles) {
                unsigned char* raw_a; unsigned char* raw_b;
                if constexpr (kDirect) {
                    const int s = (int)(issued % kSlots);
                    mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)(((issued / kSlots) & 1) ^ 1));   // the MMAs of this slot's previous use are done
                    raw_a = smem + s * kStageBytes; raw_b = raw_a + 2 * kTileBytes;
                } else {
                    raw_a = raw_ring + (int)(issued % kRaw) * kRawBytes; raw_b = raw_a + kTileBytes;
                }
                const int r0 = nc.kb + nt_k * kBK;
                issue_tile<KA>(a, nwa, nc.i0, M, r0, nc.ke, tid, kBM, raw_a);
                issue_tile<KB>(b, nwb, nc.j0, N, r0, nc.ke, tid, mma_n, raw_b);
                ++issued;
                if (++nt_k == nc.nkb) { nt_k = 0; nt_tile += gridDim.x; open_next(); }
            }
            cp_async_commit();
        };
        issue_next();
        if constexpr (kDirect) issue_next();   // K-major plan: two k-blocks in flight
        long long git = 0;                   // k-blocks consumed so far (stage / phase bookkeeping shared with the MMA lane)
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const TileCoord c = decode_tile(t, tiles_m, tiles_n, K, k_per_split);
            for (int it = 0; it < c.nkb; ++it, ++git) {
                const int s = (int)(git % kRing);
                unsigned char* st = smem + s * kStageBytes;
                if constexpr (kDirect) {
                    TC_PROF_T(p0);
                    cp_async_wait<1>();                             // this thread's copies of the current k-block have landed (the next one may be in flight)
                    TC_PROF_T(p1);
                    asm volatile("bar.sync 1, %0;" ::"n"(kProducers) : "memory");   // ... and every other producer's
                    TC_PROF_T(p2);
                    lo_tile<KA>(tid, kBM, st, st + kTileBytes);
                    lo_tile<KB>(tid, mma_n, st + 2 * kTileBytes, st + 3 * kTileBytes);
                    TC_PROF_T(p3);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // landed copies and lo stores -> async-proxy (UMMA) reads
                    mbar_arrive(smem_u32(&bar_full[s]));
                    TC_PROF_T(p4);
                    issue_next();                                   // k-block git + 2 into the slot the MMAs of k-block git - 1 are leaving
                    TC_PROF_T(p5);
                    TC_PROF_ADD(0, p0, p1); TC_PROF_ADD(1, p1, p2); TC_PROF_ADD(4, p2, p3); TC_PROF_ADD(5, p3, p4); TC_PROF_ADD(2, p4, p5);
                    TC_PROF_ADD(6, 0, 1);
                } else {
                    TC_PROF_T(p0);
                    cp_async_wait<0>();                             // this thread's copies of the current k-block have landed
                    TC_PROF_T(p1);
                    // every producer's copies have landed, and every producer is done reading the raw block refilled next
                    asm volatile("bar.sync 1, %0;" ::"n"(kProducers) : "memory");
                    TC_PROF_T(p2);
                    issue_next();
                    TC_PROF_T(p3);
                    mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)(((git / kStages) & 1) ^ 1));   // hi/lo stage free (first pass: returns at once)
                    TC_PROF_T(p4);
                    const unsigned char* raw = raw_ring + (int)(git % kRaw) * kRawBytes;
                    const int r0 = c.kb + it * kBK;
                    transform_tile<KA>(a, c.i0, M, r0, c.ke, tid, kBM, raw, st, st + kTileBytes);
                    transform_tile<KB>(b, c.j0, N, r0, c.ke, tid, mma_n, raw + kTileBytes, st + 2 * kTileBytes, st + 3 * kTileBytes);
                    TC_PROF_T(p5);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async-proxy (UMMA) reads
                    mbar_arrive(smem_u32(&bar_full[s]));
                    TC_PROF_T(p6);
                    TC_PROF_ADD(0, p0, p1); TC_PROF_ADD(1, p1, p2); TC_PROF_ADD(2, p2, p3); TC_PROF_ADD(3, p3, p4); TC_PROF_ADD(4, p4, p5);
                    TC_PROF_ADD(5, p5, p6); TC_PROF_ADD(6, 0, 1);
                }
            }
        }
        cp_async_wait<0>();
        }   // !tma_done
    } else if (warp == kMmaWarp) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t idesc = make_idesc(kBM, mma_n);
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
                    const int s = (int)(git % kRing);
                    const uint32_t phase = (uint32_t)((git / kRing) & 1);
                    TC_PROF_T(q0);
                    mbar_wait(smem_u32(&bar_full[s]), phase);
                    TC_PROF_T(q1);
                    TC_PROF_ADD(0, q0, q1); TC_PROF_ADD(1, 0, 1);
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
        // ---------------- epilogue warps: TMEM lane quadrant = warp % 4 ----------------
        const int quad = warp & 3;
        float* tile = epi_stage + (warp - kEpiWarp0) * (32 * 33);
        const bool split = n_splits > 1;
        // Row segments straight from the accumulator registers (see HasRow4) - only for outputs one column tile wide (the NLayer forward and
        // its split partials: rows of <= 512 bytes, so the 16-byte pieces of a warp's store fill whole lines between them).  Measured on the
        // 2 880-wide input gradient the lane-per-row stores (16 bytes at an 11.5 KB stride) are SLOWER than the staged, line-wide ones:
        // 55.9 -> 71.3 us at mini-batch 8 192; on the forward 49.7 -> 49.0 us.
        bool direct_rows = false;
        if constexpr (!LaneIsRow<EP>::value) {
            if ((N & 3) == 0 && tiles_n == 1) {
                if (split) direct_rows = (reinterpret_cast<uintptr_t>(partials) & 15) == 0;
                else if constexpr (HasRow4<EP>::value) direct_rows = ep.row4_ok();
            }
        }
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
                    if constexpr (LaneIsRow<EP>::value) {
                        // lane = tile row (as the registers come out of tensor memory); staged column-major so that the
                        // loop over columns below indexes shared memory, not registers (conflict-free: bank = q + lane)
#pragma unroll
                        for (int q = 0; q < 32; ++q) tile[q * 33 + lane] = __uint_as_float(v[q]);
                        __syncwarp();
                        const int gr = c.i0 + quad * 32 + lane;
                        if (gr < M) {
                            const int n_cols = min(32, N - (c.j0 + c0));
                            // (a full block runs without a trip-count test: the shared-memory loads of eight columns are in flight together -
                            // with ONE warp per scheduler draining the accumulator, the epilogue is bound by its own dependent-issue latency)
                            if (n_cols == 32 && !split) {
#pragma unroll 8
                                for (int q = 0; q < 32; ++q) ep(gr, c.j0 + c0 + q, tile[q * 33 + lane]);
                            } else {
#pragma unroll 4
                                for (int q = 0; q < n_cols; ++q) {
                                    const float accv = tile[q * 33 + lane];
                                    const int gj = c.j0 + c0 + q;
                                    if (split) partials[((size_t)c.z * M + gr) * N + gj] = accv;
                                    else ep(gr, gj, accv);
                                }
                            }
                        }
                        __syncwarp();
                    } else if (direct_rows) {
                        // lane = row: its 32 registers are 128 contiguous bytes of the output row (or of the split's partial row)
                        const int gr = c.i0 + quad * 32 + lane;
                        if (gr < M) {
                            float* prow = split ? partials + ((size_t)c.z * M + gr) * N : nullptr;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int gj = c.j0 + c0 + 4 * i;
                                if (gj < N) {   // (N is a multiple of 4 here: a group of four columns is valid as a whole)
                                    const float4 a4 = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                                                  __uint_as_float(v[4 * i + 3]));
                                    if (split) *reinterpret_cast<float4*>(prow + gj) = a4;
                                    else if constexpr (HasRow4<EP>::value) ep.row4(gr, gj, a4);
                                }
                            }
                        }
                    } else {
                        // 32 x 32 block: lane = row in registers -> shared (padded) -> lane = column, so that every warp-wide access
                        // of the epilogue (stores, bias / previous-activation loads) covers 128 contiguous bytes
#pragma unroll
                        for (int q = 0; q < 32; ++q) tile[lane * 33 + q] = __uint_as_float(v[q]);
                        __syncwarp();
                        const int gj = c.j0 + c0 + lane;
                        if (gj < N) {
                            const int gr0 = c.i0 + quad * 32;
                            if (gr0 + 32 <= M) {   // a full block: no row test, eight rows in flight (see above)
                                if (split) {
                                    float* pp = partials + ((size_t)c.z * M + gr0) * N + gj;
#pragma unroll 8
                                    for (int r = 0; r < 32; ++r) pp[(size_t)r * N] = tile[r * 33 + lane];
                                } else {
#pragma unroll 8
                                    for (int r = 0; r < 32; ++r) ep(gr0 + r, gj, tile[r * 33 + lane]);
                                }
                            } else {
#pragma unroll 4
                                for (int r = 0; r < 32; ++r) {
                                    const int gr = gr0 + r;
                                    if (gr >= M) break;
                                    const float accv = tile[r * 33 + lane];
                                    if (split) partials[((size_t)c.z * M + gr) * N + gj] = accv;
                                    else ep(gr, gj, accv);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    }
#ifdef DLB_TC_PROFILE
    if (tid == 0) for (int i = 0; i < 7; ++i) atomicAdd(&g_tc_prof[i], (unsigned long long)prof_acc[i]);
    if (tid == kMmaWarp * 32) for (int i = 0; i < 2; ++i) atomicAdd(&g_tc_prof[8 + i], (unsigned long long)prof_acc[i]);
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kAccCols) : "memory");
    }
}

}  // namespace tc

// Host launcher, instantiated per pair of operand kinds (the kernel carries only the producers it needs).  Falls back to
// nothing: callers decide (by Real type and shape) whether to use this or gemm<R>.
template <int KA, int KB, class EP>
void tc_gemm(cudaStream_t st, int M, int N, int K, const TcOperand& a, const TcOperand& b, EP ep, const GemmWorkspace& ws, bool pdl = false,
             int* defer_finalize = nullptr) {   // non-null: a split reduction leaves its partials [splits][M][N] in ws for the consumer; receives `splits`
    if (defer_finalize) *defer_finalize = 1;
    if (M <= 0 || N <= 0) return;
    if (a.kind != KA || b.kind != KB) throw std::logic_error("tc_gemm: operand kind does not match the instantiation");
    if (a.ld >= (1LL << 31) || b.ld >= (1LL << 31)) throw std::invalid_argument("tc_gemm: leading dimension exceeds the 32-bit range");
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
    const int mma_n = tiles_n == 1 ? std::min(tc::kBN, ceil_div(N, 16) * 16) : tc::kBN;
    // TMA for dense, ungathered, 16-byte aligned row operands (the NLayer forward): box {32 fp32, rows}, SWIZZLE_128B, zero fill
    CUtensorMap map_a{}, map_b{};
    int use_tma = 0;
    if constexpr (KA == kRedContig && KB == kRedContig) {
        use_tma = tc::tma_enabled() && tc::make_row_map(map_a, a, M, K, tc::kBM) && tc::make_row_map(map_b, b, N, K, mma_n);
    }
    const dim3 grid((unsigned)std::min(n_tiles, kNumSMs));
    float* partials = static_cast<float*>(ws.ptr);
    if (pdl && !g_timing_on) {   // programmatic dependent launch: the prologue overlaps the previous kernel of the stream (common.cuh)
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        auto kernel = tc::tc_gemm_kernel<EP, KA, KB>;
        allow_big_smem((const void*)kernel, (size_t)tc::kSmemBytes);
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = dim3(tc::kThreads); cfg.dynamicSmemBytes = (size_t)tc::kSmemBytes; cfg.stream = st;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        DLB_CUDA(cudaLaunchKernelEx(&cfg, kernel, M, N, K, k_per_split, tiles_m, tiles_n, n_tiles, splits, mma_n, a, b, ep, partials, use_tma, map_a, map_b));
    } else {
        DLB_LAUNCH_SMEM((tc::tc_gemm_kernel<EP, KA, KB>), grid, dim3(tc::kThreads), (size_t)tc::kSmemBytes, st, M, N, K, k_per_split, tiles_m, tiles_n,
                        n_tiles, splits, mma_n, a, b, ep, partials, use_tma, map_a, map_b);
    }
    if (defer_finalize) { *defer_finalize = splits; return; }
    if (splits > 1) {
        const size_t total = (size_t)M * N;
        DLB_LAUNCH((gemm_finalize_kernel<float, EP>), dim3((unsigned)ceil_div<size_t>(total, 256)), dim3(256), st,
                   (const float*)partials, splits, M, N, ep);
    }
}

#endif  // !DLB_CUDA_EMU

}  // namespace dlb

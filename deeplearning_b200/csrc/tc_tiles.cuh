// tc_tiles.cuh — operand tiles of the tensor-core contraction (tc_gemm.cuh): which global element lands where in the
// shared-memory tile that tcgen05.mma reads.  Pure index arithmetic around two copy primitives, written as
// __host__ __device__ code so that tests/cpp/test_tc_tiles.cu can run every producer thread's part on the CPU and compare
// the finished tiles with the definition value(outer, red) element by element (the only part of the tensor-core path that
// can be checked without a GPU).
//
// An operand is a function value(outer, red): `outer` is the output row (A) or column (B) of the contraction, `red` the
// reduction index.  Kinds:
//   kRedContig   base[row(outer) * ld + red]                 NLayer forward X and W (NLayer.inl:131-142)
//   kOuterContig base[row(red) * ld + outer]                 W^T, G^T, X^T of NLayer::backpropagate (NLayer.inl:144-171)
//   kPatchOuter  receptive-field matrix, outer = pixel (b, py, px), red = (c, ky, kx)
//                                                            CLayer::act (Tensor.cpp:371-408: rows of the im2col matrix) and,
//                                                            with the window mirrored (sgn = -1), the input gradient
//                                                            (Tensor.cpp:500-546)
//   kPatchRed    the same matrix with outer = (c, ky, kx), red = pixel: the filter gradient (Tensor.cpp:466-498)
//   kPixelRed    G[b][f][p] with outer = f, red = (b, p): the other operand of the filter gradient
// A mini-batch is [B][C][H][W]; `gather` (layer 0) maps the batch slot to the row of the resident data set.
#pragma once

#include <cstdint>
#include <cstring>

#ifndef __CUDACC__
#define DLB_HD inline
#else
#define DLB_HD __host__ __device__ __forceinline__
#endif

namespace dlb {

enum TcKind : int { kRedContig = 0, kOuterContig = 1, kPatchOuter = 2, kPatchRed = 3, kPixelRed = 4 };

// Geometry of a receptive-field operand: source tensor [SC][SH][SW] per sample, pixel grid PH x PW, window KH x KW.
// Window tap (ky, kx) of pixel (py, px) reads source row py * sy + dy + sgn * ky, column px * sx + dx + sgn * kx
// (forward: dy = -pad, sgn = +1; input gradient of a stride-1 convolution: dy = +pad, sgn = -1, source = output gradient).
// mKK / mKW / mPW: ceil(2^32 / d) reciprocals (0 for d = 1), exact for the operand ranges tc_conv_ok admits.
struct TcPatch {
    int SC, SH, SW, PH, PW, KH, KW, sy, sx, dy, dx, sgn;
    int SHW, KK, PP;     // SH * SW, KH * KW, PH * PW
    unsigned mKK, mKW, mPW;
};

struct TcOperand {
    const float* base;
    long long ld;              // kRedContig / kOuterContig: leading dimension; kPixelRed: unused
    int kind;                  // TcKind (the kernel is instantiated per pair of kinds; kept here for the launchers)
    const int* gather_base;    // optional row / sample gather (device int32), addressed through *cursor when cursor != nullptr
    const int* cursor;
    int ones_col;              // kOuterContig: outer index that reads as 1 (bias folded into the contraction); -1: none
    TcPatch patch;             // kPatchOuter / kPatchRed: the geometry; kPixelRed: SC = F, PP = P
};

inline unsigned tc_reciprocal(int d) { return d <= 1 ? 0u : (unsigned)((0x100000000ull + (unsigned)d - 1) / (unsigned)d); }

inline TcPatch make_patch(int SC, int SH, int SW, int PH, int PW, int KH, int KW, int sy, int sx, int dy, int dx, int sgn) {
    TcPatch p{SC, SH, SW, PH, PW, KH, KW, sy, sx, dy, dx, sgn, SH * SW, KH * KW, PH * PW,
              tc_reciprocal(KH * KW), tc_reciprocal(KW), tc_reciprocal(PW)};
    return p;
}

namespace tc {

constexpr int kBM = 128, kBN = 128, kBK = 32;
constexpr int kProducerWarps = 8;                         // the receptive-field lane mappings assume 8 (one 16-byte chunk column per warp)
constexpr int kProducers = kProducerWarps * 32;            // producer threads
constexpr int kOwned = 1024 / kProducers;                  // 16-byte chunks of a 128 x 32 fp32 tile per producer thread
constexpr int kTileBytes = kBM * kBK * 4;                  // 16 KiB: one raw, hi or lo tile of either operand

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
#endif

// The two copy primitives: asynchronous global -> shared copies on the device (zero fill through cp.async's src-size
// operand), plain copies in the host check.
DLB_HD void tile_copy16(unsigned char* dst, const float* src, int ok_bytes) {
#ifdef __CUDA_ARCH__
    cp_async_16(smem_u32(dst), src, (uint32_t)ok_bytes);
#else
    std::memset(dst, 0, 16);
    std::memcpy(dst, src, (size_t)ok_bytes);
#endif
}
DLB_HD void tile_copy4(unsigned char* dst, const float* src, bool ok) {
#ifdef __CUDA_ARCH__
    cp_async_4(smem_u32(dst), src, ok ? 4u : 0u);
#else
    const float v = ok ? *src : 0.f;
    std::memcpy(dst, &v, 4);
#endif
}
struct alignas(16) Quad { float x, y, z, w; };

DLB_HD Quad tile_load16(const unsigned char* p) { return *reinterpret_cast<const Quad*>(p); }
DLB_HD float tile_load4(const unsigned char* p) { return *reinterpret_cast<const float*>(p); }
DLB_HD void tile_store16(unsigned char* p, Quad v) { *reinterpret_cast<Quad*>(p) = v; }
DLB_HD int load_index(const int* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
DLB_HD int imin(int a, int b) { return a < b ? a : b; }

// Byte offset of 16-byte chunk c of tile row `row` in the K-major SWIZZLE_128B layout (row = 32 fp32 = 128 B; chunk c of
// row r is stored at chunk c ^ (r & 7)).
DLB_HD int chunk_off(int row, int c) { return row * 128 + ((c ^ (row & 7)) << 4); }

// n / d for n * d < 2^32 through the precomputed reciprocal m = ceil(2^32 / d) (m = 0 encodes d = 1)
DLB_HD int fast_div(int n, unsigned m) { return m ? (int)(((unsigned long long)(unsigned)n * m) >> 32) : n; }

// Bits [lo, hi] of a 16-bit field (empty when lo > hi); lo, hi are clamped to [0, 14].
DLB_HD unsigned bit_range(int lo, int hi) {
    lo = lo < 0 ? 0 : lo; hi = hi > 14 ? 14 : hi;
    return lo > hi ? 0u : (((2u << hi) - 1u) & ~((1u << lo) - 1u));
}
// Which window taps of a pixel whose tap (0, 0) reads source (iy0, ix0) fall inside the source: bit ky (rows) and bit
// 16 + kx (columns).  A tap is valid iff both of its bits are set; windows are at most 15 x 15 here (tc_conv_ok).
DLB_HD unsigned tap_mask(const TcPatch& p, int iy0, int ix0) {
    const unsigned my = p.sgn > 0 ? bit_range(-iy0, p.SH - 1 - iy0) : bit_range(iy0 - p.SH + 1, iy0);
    const unsigned mx = p.sgn > 0 ? bit_range(-ix0, p.SW - 1 - ix0) : bit_range(ix0 - p.SW + 1, ix0);
    return my | (mx << 16);
}
constexpr unsigned kNoTap = 0x80008000u;   // a selector no tap_mask contains

// Window tap k = (c, ky, kx) -> source offset relative to the pixel's tap (0, 0) and the selector matched against tap_mask
DLB_HD void tap_decode(const TcPatch& p, int k, int& koff, unsigned& sel) {
    const int c = fast_div(k, p.mKK), q = k - c * p.KK;
    const int ky = fast_div(q, p.mKW), kx = q - ky * p.KW;
    koff = c * p.SHW + p.sgn * (ky * p.SW + kx);
    sel = (1u << ky) | (1u << (16 + kx));
}

// What a producer thread keeps per operand while it works on one output tile.
constexpr int kPatchElems = 128 * kBK / kProducers;   // receptive-field elements per thread and k-block (16)
template <int KIND> struct TileOwner {};
template <> struct TileOwner<kRedContig> { int grow[kOwned]; };           // stored row of each owned chunk
template <> struct TileOwner<kPixelRed> { int grow[kOwned]; };            // f * P of each owned tile row
template <> struct TileOwner<kOuterContig> {};
template <> struct TileOwner<kPatchOuter> {   // this thread's 16 pixels (tile rows 8 g + (lane >> 2))
    int pbase[kPatchElems];        // element offset of tap (0, 0) of the pixel (may point outside: masked)
    unsigned mask[kPatchElems];    // tap_mask of the pixel, 0 beyond the operand
};
template <> struct TileOwner<kPatchRed> {     // this thread's 16 window taps (tile rows warp + 8 i)
    int koff[kPatchElems];
    unsigned sel[kPatchElems];     // kNoTap beyond the operand
};

DLB_HD const int* resolve_rows(const TcOperand& op) {
    return op.gather_base ? op.gather_base + (op.cursor ? load_index(op.cursor) : 0) : nullptr;
}

// Thread t of the producers; the tile covers outer indices [o0, o0 + 128) of [0, o_end).
template <int KIND>
DLB_HD void owner_init(const TcOperand& op, int o0, int o_end, int t, TileOwner<KIND>& w) {
    if constexpr (KIND == kRedContig || KIND == kPixelRed) {
        const int* gather = KIND == kRedContig ? resolve_rows(op) : nullptr;
#pragma unroll
        for (int i = 0; i < kOwned; ++i) {
            const int o = imin(o0 + ((t + kProducers * i) >> 3), o_end - 1);
            if constexpr (KIND == kPixelRed) w.grow[i] = o * op.patch.PP;
            else w.grow[i] = gather ? load_index(gather + o) : o;
        }
    } else if constexpr (KIND == kPatchOuter) {
        const TcPatch& p = op.patch;
        const int* gather = resolve_rows(op);
        const int lane = t & 31;
#pragma unroll
        for (int g = 0; g < kPatchElems; ++g) {
            const int pix = o0 + 8 * g + (lane >> 2);
            if (pix < o_end) {
                const int b = pix / p.PP, rem = pix - b * p.PP;
                const int py = fast_div(rem, p.mPW), px = rem - py * p.PW;
                const int s = gather ? load_index(gather + b) : b;
                const int iy0 = py * p.sy + p.dy, ix0 = px * p.sx + p.dx;
                w.pbase[g] = s * p.SC * p.SHW + iy0 * p.SW + ix0;
                w.mask[g] = tap_mask(p, iy0, ix0);
            } else {
                w.pbase[g] = 0; w.mask[g] = 0;
            }
        }
    } else if constexpr (KIND == kPatchRed) {
        const TcPatch& p = op.patch;
#pragma unroll
        for (int i = 0; i < kPatchElems; ++i) {
            const int k = o0 + (t >> 5) + 8 * i;
            if (k < o_end) tap_decode(p, k, w.koff[i], w.sel[i]);
            else { w.koff[i] = 0; w.sel[i] = kNoTap; }
        }
    }
}

// Issues this thread's copies of one k-block (reduction indices [r0, r0 + 32) of [.., r_end)) into the raw tile.
// `rows`: tile rows the MMA reads (128 for A, the MMA's N for B); rows beyond are neither copied nor transformed.
// Raw layouts: every kind but kOuterContig is copied straight into the K-major swizzled layout; kOuterContig as plain
// [32 reduction rows][128 outer] and transposed by transform_tile.
// Receptive-field kinds move single fp32 elements (patch rows are not 16-byte aligned).  Their lane mappings make both
// sides of the copy efficient: kPatchOuter gives a warp 8 consecutive pixels x 4 consecutive taps of one 16-byte chunk
// column (global: a few 128-byte lines; shared: the swizzle spreads the 8 rows over all 32 banks), kPatchRed gives a
// warp 32 consecutive pixels of one tap (global: consecutive addresses; shared: one whole 128-byte tile row).
template <int KIND>
DLB_HD void issue_tile(const TcOperand& op, const TileOwner<KIND>& w, int o0, int o_end, int r0, int r_end, int t, int rows,
                       unsigned char* raw) {
    if constexpr (KIND == kRedContig) {
        const int ld = (int)op.ld;   // leading dimensions are below 2^31 (checked by the launcher): one 32 x 32 -> 64 multiply per address
        const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(op.base) & 15) == 0) && (((r_end - r0) & 3) == 0 || r0 + kBK <= r_end);
        const int c = t & 7;
#pragma unroll
        for (int i = 0; i < kOwned; ++i) {
            const int row = (t + kProducers * i) >> 3;
            if (row >= rows) continue;
            if (vec) {
                const int r = r0 + 4 * c;
                const bool ok = (o0 + row < o_end) && (r < r_end);
                tile_copy16(raw + chunk_off(row, c), op.base + (long long)w.grow[i] * ld + imin(r, r_end - 4), ok ? 16 : 0);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = r0 + 4 * c + j;
                    const bool ok = (o0 + row < o_end) && (r < r_end);
                    tile_copy4(raw + chunk_off(row, c) + 4 * j, op.base + (long long)w.grow[i] * ld + imin(r, r_end - 1), ok);
                }
            }
        }
    } else if constexpr (KIND == kOuterContig) {
        const int* gather = resolve_rows(op);
        const int ld = (int)op.ld;
        const int o_lim = imin(o_end, ld);                         // the column of ones (o == ld) is synthesised in transform_tile
        const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(op.base) & 15) == 0);
        if (vec) {
            const int oc = (t & 31) * 4;
            if (oc < rows) {
#pragma unroll
                for (int i = 0; i < kOwned; ++i) {
                    const int k = (t + kProducers * i) >> 5;
                    const int r = imin(r0 + k, r_end - 1);
                    const int grow = gather ? load_index(gather + r) : r;
                    const int o = o0 + oc;
                    int n_ok = (r0 + k < r_end) ? imin(4, o_lim - o) : 0;
                    n_ok = n_ok < 0 ? 0 : n_ok;
                    tile_copy16(raw + (k * 128 + oc) * 4, op.base + (long long)grow * ld + imin(o, ld - 4), 4 * n_ok);
                }
            }
        } else {
            const int tt = t & 127, k0 = (t >> 7) * (kBK * 128 / kProducers);
            if (tt < rows) {
                const int o = o0 + tt;
                const int oc = imin(imin(o, o_end - 1), ld - 1);
#pragma unroll
                for (int kk = 0; kk < kBK * 128 / kProducers; ++kk) {
                    const int k = k0 + kk;
                    const int r = imin(r0 + k, r_end - 1);
                    const int grow = gather ? load_index(gather + r) : r;
                    tile_copy4(raw + (k * 128 + tt) * 4, op.base + (long long)grow * ld + oc, o < o_lim && r0 + k < r_end);
                }
            }
        }
    } else if constexpr (KIND == kPatchOuter) {
        // this thread: tap k = r0 + 4 * warp + (lane & 3) of its 16 pixels
        const int lane = t & 31, warp = t >> 5;
        const int k = r0 + 4 * warp + (lane & 3);
        int koff = 0; unsigned sel = kNoTap;
        if (k < r_end) tap_decode(op.patch, k, koff, sel);
        unsigned char* dst = raw + chunk_off(lane >> 2, warp) + 4 * (lane & 3);
#pragma unroll
        for (int g = 0; g < kPatchElems; ++g) {
            if (8 * g >= rows) continue;
            const bool ok = (w.mask[g] & sel) == sel;
            tile_copy4(dst + g * 1024, op.base + (ok ? w.pbase[g] + koff : 0), ok);
        }
    } else if constexpr (KIND == kPatchRed) {
        // this thread: pixel r0 + lane under its 16 taps
        const TcPatch& p = op.patch;
        const int lane = t & 31, warp = t >> 5;
        const int pix = r0 + lane;
        int sbase = 0; unsigned mask = 0;
        if (pix < r_end) {
            const int* gather = resolve_rows(op);
            const int b = pix / p.PP, rem = pix - b * p.PP;
            const int py = fast_div(rem, p.mPW), px = rem - py * p.PW;
            const int s = gather ? load_index(gather + b) : b;
            const int iy0 = py * p.sy + p.dy, ix0 = px * p.sx + p.dx;
            sbase = s * p.SC * p.SHW + iy0 * p.SW + ix0;
            mask = tap_mask(p, iy0, ix0);
        }
        unsigned char* dst = raw + chunk_off(warp, lane >> 2) + 4 * (lane & 3);
#pragma unroll
        for (int i = 0; i < kPatchElems; ++i) {
            if (warp + 8 * i >= rows) continue;
            const bool ok = (mask & w.sel[i]) == w.sel[i];
            tile_copy4(dst + i * 1024, op.base + (ok ? sbase + w.koff[i] : 0), ok);
        }
    } else {   // kPixelRed: G[(b * F + f) * P + p], tile row = f, reduction = (b, p)
        const TcPatch& p = op.patch;
        const int c = t & 7;
        const int r = r0 + 4 * c;
        int b = r / p.PP, pp = r - b * p.PP;
        const bool vec = ((p.PP & 3) == 0) && ((reinterpret_cast<uintptr_t>(op.base) & 15) == 0);
        if (vec) {
            const int off = b * p.SC * p.PP + pp;                  // B * F * P < 2^31 (tc_conv_ok)
#pragma unroll
            for (int i = 0; i < kOwned; ++i) {
                const int row = (t + kProducers * i) >> 3;
                if (row >= rows) continue;
                const bool ok = (o0 + row < o_end) && (r < r_end);
                tile_copy16(raw + chunk_off(row, c), op.base + (ok ? off + w.grow[i] : 0), ok ? 16 : 0);
            }
        } else {
            int off[4]; bool okr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                okr[j] = r + j < r_end;
                off[j] = b * p.SC * p.PP + pp;
                if (++pp == p.PP) { pp = 0; ++b; }
            }
#pragma unroll
            for (int i = 0; i < kOwned; ++i) {
                const int row = (t + kProducers * i) >> 3;
                if (row >= rows) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool ok = (o0 + row < o_end) && okr[j];
                    tile_copy4(raw + chunk_off(row, c) + 4 * j, op.base + (ok ? off[j] + w.grow[i] : 0), ok);
                }
            }
        }
    }
}

// raw fp32 chunk -> hi (TF32-representable: low 13 mantissa bits cleared) and lo (exact remainder) chunks at the same
// offsets of the hi / lo tiles
DLB_HD float tf32_hi(float v) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
#else
    uint32_t u; std::memcpy(&u, &v, 4);
    u &= 0xFFFFE000u;
    float h; std::memcpy(&h, &u, 4);
    return h;
#endif
}
DLB_HD void split_store(Quad v, unsigned char* hi_tile, unsigned char* lo_tile, int off) {
    Quad h, l;
    h.x = tf32_hi(v.x); l.x = v.x - h.x;
    h.y = tf32_hi(v.y); l.y = v.y - h.y;
    h.z = tf32_hi(v.z); l.z = v.z - h.z;
    h.w = tf32_hi(v.w); l.w = v.w - h.w;
    tile_store16(hi_tile + off, h);
    tile_store16(lo_tile + off, l);
}

// K-major plan: the copied tile itself is the MMA's hi operand (the tensor core ignores the low 13 mantissa bits); only
// the remainder lo = x - (x & 0xFFFFE000) is computed.
template <int KIND>
DLB_HD void lo_tile(int t, int rows, const unsigned char* raw_tile, unsigned char* lo_tile_) {
    static_assert(KIND != kOuterContig, "transposed operands go through transform_tile");
#pragma unroll
    for (int i = 0; i < kOwned; ++i) {
        const int q = t + kProducers * i;
        if ((q >> 3) >= rows) continue;
        const int off = chunk_off(q >> 3, q & 7);
        const Quad v = tile_load16(raw_tile + off);
        Quad l;
        l.x = v.x - tf32_hi(v.x); l.y = v.y - tf32_hi(v.y); l.z = v.z - tf32_hi(v.z); l.w = v.w - tf32_hi(v.w);
        tile_store16(lo_tile_ + off, l);
    }
}

template <int KIND>
DLB_HD void transform_tile(const TcOperand& op, int o0, int o_end, int r0, int r_end, int t, int rows,
                           const unsigned char* raw_tile, unsigned char* hi_tile, unsigned char* lo_tile) {
    if constexpr (KIND != kOuterContig) {
#pragma unroll
        for (int i = 0; i < kOwned; ++i) {
            const int q = t + kProducers * i;
            if ((q >> 3) >= rows) continue;
            const int off = chunk_off(q >> 3, q & 7);
            split_store(tile_load16(raw_tile + off), hi_tile, lo_tile, off);
        }
    } else {
        // thread t produces half of tile row t & 127 (outer index o0 + (t & 127)): a column of the raw [k][o] block
        // (conflict-free: lanes read consecutive words), or the implicit column of ones
        const int tt = t & 127, c0 = (t >> 7) * (8 * 128 / kProducers);
        if (tt >= rows) return;
        const bool ones = (o0 + tt < o_end) && (o0 + tt == op.ones_col);
#pragma unroll
        for (int cc = 0; cc < 8 * 128 / kProducers; ++cc) {
            const int c = c0 + cc;
            Quad v;
            if (ones) {
                v.x = r0 + 4 * c < r_end ? 1.f : 0.f; v.y = r0 + 4 * c + 1 < r_end ? 1.f : 0.f;
                v.z = r0 + 4 * c + 2 < r_end ? 1.f : 0.f; v.w = r0 + 4 * c + 3 < r_end ? 1.f : 0.f;
            } else {
                v.x = tile_load4(raw_tile + ((4 * c) * 128 + tt) * 4); v.y = tile_load4(raw_tile + ((4 * c + 1) * 128 + tt) * 4);
                v.z = tile_load4(raw_tile + ((4 * c + 2) * 128 + tt) * 4); v.w = tile_load4(raw_tile + ((4 * c + 3) * 128 + tt) * 4);
            }
            split_store(v, hi_tile, lo_tile, chunk_off(tt, c));
        }
    }
}

}  // namespace tc
}  // namespace dlb

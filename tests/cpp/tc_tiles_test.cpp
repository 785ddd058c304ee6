// CPU check of the operand-tile producers of the tensor-core contraction (deeplearning_b200/csrc/tc_tiles.cuh): every
// producer thread's copies and hi/lo transform are run on the host and the finished K-major SWIZZLE_128B tiles are
// compared element by element with the definition of the operand — dense rows / columns (NLayer), the receptive-field
// matrix in both orientations (Tensor.cpp:40-72, 466-546) and the [b][f][pixel] gradient.  hi + lo must reproduce the
// source value exactly and hi must be TF32-representable.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../deeplearning_b200/csrc/tc_tiles.cuh"

using namespace dlb;
using namespace dlb::tc;

static std::mt19937 rng(12345);
static int fails = 0;

static float* aligned_floats(size_t n, int misalign_floats) {
    void* p = nullptr;
    if (posix_memalign(&p, 64, (n + 16) * sizeof(float)) != 0) std::abort();
    float* f = static_cast<float*>(p) + misalign_floats;
    std::uniform_real_distribution<float> d(-1.f, 1.f);
    for (size_t i = 0; i < n; ++i) f[i] = d(rng);
    return f;
}

// value(outer, red) by the definition
static float expect(const TcOperand& op, const int* gather, int o, int r) {
    const TcPatch& p = op.patch;
    switch (op.kind) {
    case kRedContig: return op.base[(long long)(gather ? gather[o] : o) * op.ld + r];
    case kOuterContig:
        if (o == op.ones_col) return 1.f;
        return op.base[(long long)(gather ? gather[r] : r) * op.ld + o];
    case kPatchOuter:
    case kPatchRed: {
        const int pix = op.kind == kPatchOuter ? o : r, k = op.kind == kPatchOuter ? r : o;
        const int b = pix / p.PP, rem = pix % p.PP, py = rem / p.PW, px = rem % p.PW;
        const int c = k / p.KK, q = k % p.KK, ky = q / p.KW, kx = q % p.KW;
        const int iy = py * p.sy + p.dy + p.sgn * ky, ix = px * p.sx + p.dx + p.sgn * kx;
        if (iy < 0 || iy >= p.SH || ix < 0 || ix >= p.SW) return 0.f;
        const long long s = gather ? gather[b] : b;
        return op.base[(s * p.SC + c) * p.SHW + iy * p.SW + ix];
    }
    default: {
        const int b = r / p.PP, pp = r % p.PP;
        return op.base[((long long)b * p.SC + o) * p.PP + pp];
    }
    }
}

template <int KIND>
static void run_kind(const TcOperand& op, int o0, int o_end, int r0, int r_end, int rows, unsigned char* raw, unsigned char* hi, unsigned char* lo) {
    for (int t = 0; t < kProducers; ++t) {
        TileOwner<KIND> w{};
        owner_init<KIND>(op, o0, o_end, t, w);
        issue_tile<KIND>(op, w, o0, o_end, r0, r_end, t, rows, raw);
    }
    for (int t = 0; t < kProducers; ++t) transform_tile<KIND>(op, o0, o_end, r0, r_end, t, rows, raw, hi, lo);
    if constexpr (KIND != kOuterContig) {
        // K-major plan: the copied tile is the hi operand as it is (the tensor core drops the low 13 mantissa bits), lo_tile
        // adds the remainder: truncated raw must equal the transform's hi and lo_tile's output its lo, bit for bit
        std::vector<unsigned char> lo2(kTileBytes, 0xCD);
        for (int t = 0; t < kProducers; ++t) lo_tile<KIND>(t, rows, raw, lo2.data());
        for (int row = 0; row < rows; ++row)
            for (int k = 0; k < kBK; ++k) {
                const int off = chunk_off(row, k >> 2) + 4 * (k & 3);
                float r, h, l, l2;
                std::memcpy(&r, raw + off, 4); std::memcpy(&h, hi + off, 4); std::memcpy(&l, lo + off, 4); std::memcpy(&l2, lo2.data() + off, 4);
                if (tf32_hi(r) != h || std::memcmp(&l, &l2, 4) != 0) {
                    if (fails < 10) std::printf("lo_tile: row %d k %d raw %g hi %g lo %g lo2 %g\n", row, k, r, h, l, l2);
                    ++fails;
                }
            }
    }
}

static void run_producers(const TcOperand& op, int o0, int o_end, int r0, int r_end, int rows, unsigned char* raw, unsigned char* hi, unsigned char* lo) {
    switch (op.kind) {
    case kRedContig: run_kind<kRedContig>(op, o0, o_end, r0, r_end, rows, raw, hi, lo); break;
    case kOuterContig: run_kind<kOuterContig>(op, o0, o_end, r0, r_end, rows, raw, hi, lo); break;
    case kPatchOuter: run_kind<kPatchOuter>(op, o0, o_end, r0, r_end, rows, raw, hi, lo); break;
    case kPatchRed: run_kind<kPatchRed>(op, o0, o_end, r0, r_end, rows, raw, hi, lo); break;
    default: run_kind<kPixelRed>(op, o0, o_end, r0, r_end, rows, raw, hi, lo); break;
    }
}

static void check_operand(const char* name, const TcOperand& op, const int* gather, int o_end, int r_total, int rows) {
    std::vector<unsigned char> raw(kTileBytes), hi(kTileBytes), lo(kTileBytes);
    // a few reduction slices: whole, and a split that ends inside a k-block
    const int r_ends[2] = {r_total, r_total > 40 ? r_total - 7 : r_total};
    for (int o0 = 0; o0 < o_end; o0 += kBM) {
        for (int re = 0; re < 2; ++re) {
            const int r_end = r_ends[re];
            if (op.kind == kPixelRed && (op.patch.PP & 3) == 0 && (r_end & 3)) continue;   // the 16-byte path needs r_end % 4 == 0 (true on the device)
            for (int r0 = 0; r0 < r_end; r0 += kBK) {
                std::fill(raw.begin(), raw.end(), 0xCD); std::fill(hi.begin(), hi.end(), 0xCD); std::fill(lo.begin(), lo.end(), 0xCD);
                run_producers(op, o0, o_end, r0, r_end, rows, raw.data(), hi.data(), lo.data());
                for (int row = 0; row < rows; ++row)
                    for (int k = 0; k < kBK; ++k) {
                        const int off = chunk_off(row, k >> 2) + 4 * (k & 3);
                        float h, l;
                        std::memcpy(&h, hi.data() + off, 4); std::memcpy(&l, lo.data() + off, 4);
                        const int o = o0 + row, r = r0 + k;
                        const float want = (o < o_end && r < r_end) ? expect(op, gather, o, r) : 0.f;
                        uint32_t hb; std::memcpy(&hb, &h, 4);
                        if (h + l != want || (hb & 0x1FFFu) != 0) {
                            if (fails < 10)
                                std::printf("%s: tile o0 %d r0 %d (r_end %d) row %d k %d: hi %g lo %g want %g\n", name, o0, r0, r_end, row, k, h, l, want);
                            ++fails;
                        }
                    }
            }
        }
    }
}

int main() {
    // dense operands (the NLayer contractions), aligned / unaligned, with and without gather and the column of ones
    for (int variant = 0; variant < 4; ++variant) {
        const int O = 150, R = 77 + 3 * variant, ld_red = variant & 1 ? 80 : R;
        std::vector<int> gather(O > R ? O : R);
        for (size_t i = 0; i < gather.size(); ++i) gather[i] = (int)((i * 7 + 3) % gather.size());
        float* base = aligned_floats((size_t)gather.size() * (size_t)(ld_red > O ? ld_red : O + 2), variant >> 1);
        TcOperand a{base, ld_red, kRedContig, variant & 1 ? gather.data() : nullptr, nullptr, -1, {}};
        check_operand("red-contig", a, a.gather_base, O, R, kBM);
        check_operand("red-contig rows 64", a, a.gather_base, 60, R, 64);
        const int ld_out = variant & 1 ? 152 : O;
        TcOperand b{base, ld_out, kOuterContig, variant & 1 ? gather.data() : nullptr, nullptr, variant & 1 ? ld_out : -1, {}};
        check_operand("outer-contig", b, b.gather_base, variant & 1 ? ld_out + 1 : O, R, kBM);
    }
    // receptive-field operands: forward geometries (pad / stride), mirrored window (input gradient)
    struct G { int C, H, W, KH, KW, ph, pw, sh, sw; };
    const G geoms[] = {{3, 9, 8, 5, 5, 2, 2, 1, 1}, {5, 13, 17, 3, 3, 3, 6, 2, 2}, {2, 9, 8, 2, 3, 1, 0, 1, 2}, {32, 16, 16, 5, 5, 2, 2, 1, 1}, {1, 12, 12, 3, 3, 0, 0, 1, 1}};
    for (const G& g : geoms) {
        const int OH = (g.H + 2 * g.ph - g.KH) / g.sh + 1, OW = (g.W + 2 * g.pw - g.KW) / g.sw + 1;
        const int B = 3;
        std::vector<int> gather = {4, 0, 2};
        float* x = aligned_floats((size_t)5 * g.C * g.H * g.W, 0);
        const int K = g.C * g.KH * g.KW;
        for (int with_gather = 0; with_gather < 2; ++with_gather) {
            const int* ga = with_gather ? gather.data() : nullptr;
            TcOperand fo{x, 0, kPatchOuter, ga, nullptr, -1, make_patch(g.C, g.H, g.W, OH, OW, g.KH, g.KW, g.sh, g.sw, -g.ph, -g.pw, 1)};
            check_operand("patch-outer", fo, ga, B * OH * OW, K, kBM);
            TcOperand fr = fo; fr.kind = kPatchRed;
            check_operand("patch-red", fr, ga, K, B * OH * OW, kBM);
        }
        if (g.sh == 1 && g.sw == 1) {   // input gradient: source = output gradient [F][OH][OW], pixels = input pixels
            const int F = 7;
            float* gr = aligned_floats((size_t)B * F * OH * OW, 0);
            TcOperand bo{gr, 0, kPatchOuter, nullptr, nullptr, -1, make_patch(F, OH, OW, g.H, g.W, g.KH, g.KW, 1, 1, g.ph, g.pw, -1)};
            check_operand("patch-outer mirrored", bo, nullptr, B * g.H * g.W, F * g.KH * g.KW, kBM);
            // and check the mirrored definition itself against the scatter form of Tensor.cpp:500-546 on one element set
            TcOperand pr{gr, 0, kPixelRed, nullptr, nullptr, -1, make_patch(F, OH, OW, OH, OW, 1, 1, 1, 1, 0, 0, 1)};
            check_operand("pixel-red", pr, nullptr, F, B * OH * OW, kBM);
            check_operand("pixel-red rows 16", pr, nullptr, F, B * OH * OW, 16);
            TcOperand pu = pr; pu.base = gr + 1;   // unaligned base: 4-byte path
            check_operand("pixel-red unaligned", pu, nullptr, F, (B - 1) * OH * OW, kBM);
        }
    }
    if (fails) { std::printf("FAILED: %d mismatches\n", fails); return 1; }
    std::printf("OK\n");
    return 0;
}

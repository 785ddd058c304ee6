// layers.cuh — batched layer operators of the training path (forward / backward of CLayer, PLayer, NLayer, the
// cost derivative and the SGD update) over a mini-batch packed as [B][C][H][W] (sample-major, each sample in the
// reference's CHW order, Tensor.cpp:216-219).  The reference runs every one of these per sample
// (Net.inl:257-277); here each runs once per step.
#pragma once

#include <algorithm>
#include <type_traits>

#include "gemm.cuh"
#include "tc_gemm.cuh"

namespace dlb {

struct ConvGeom {
    int C, H, W, KH, KW, F, OH, OW, ph, pw, sh, sw;
    int KK, K, P, HW, in_n, out_n;
};

// ConvolutionUtils.cpp:29-44
inline int conv_out_size(int in, int k, int pad, int stride) {
    const long long t = (long long)in + 2LL * pad - k;
    return t < 0 ? 0 : (int)(t / stride + 1);
}

inline ConvGeom make_conv_geom(const dlb_layer_spec& s) {
    ConvGeom g{};
    g.C = s.in_c; g.H = s.in_h; g.W = s.in_w; g.KH = s.win_h; g.KW = s.win_w; g.F = s.out_c;
    g.ph = s.pad_h; g.pw = s.pad_w; g.sh = s.stride_h; g.sw = s.stride_w;
    g.OH = conv_out_size(g.H, g.KH, g.ph, g.sh); g.OW = conv_out_size(g.W, g.KW, g.pw, g.sw);
    g.KK = g.KH * g.KW; g.K = g.C * g.KK; g.P = g.OH * g.OW; g.HW = g.H * g.W;
    g.in_n = g.C * g.HW; g.out_n = g.F * g.P;
    return g;
}

// ---------------------------------------------------------------------------------------------------------------
// Operand loaders (see gemm.cuh for the concept).  All are called as f(outer, red).
// ---------------------------------------------------------------------------------------------------------------
template <class R>
struct DenseRowLoader {  // value = base[row(outer) * ld + red]; `gather` (optional) maps outer -> stored row
    static constexpr bool kRedFast = true;
    const R* base; size_t ld; const int* gather;
    __device__ __forceinline__ R operator()(int o, int r) const {
        const size_t row = gather ? (size_t)gather[o] : (size_t)o;
        return base[row * ld + r];
    }
    // strided protocol (gemm.cuh): without a gather the value is origin()[o * outer_step() + r * red_step()]
    __host__ __device__ bool strided() const { return gather == nullptr; }
    __device__ __forceinline__ const R* origin() const { return base; }
    __device__ __forceinline__ size_t outer_step() const { return ld; }
    __device__ __forceinline__ size_t red_step() const { return 1; }
    __device__ __forceinline__ int outer_limit() const { return 0x7fffffff; }
};

template <class R>
struct DenseColLoader {  // value = base[row(red) * ld + outer]; `gather` (optional) maps red -> stored row
    static constexpr bool kRedFast = false;
    const R* base; size_t ld; const int* gather;
    __device__ __forceinline__ R operator()(int o, int r) const {
        const size_t row = gather ? (size_t)gather[r] : (size_t)r;
        return base[row * ld + o];
    }
    __host__ __device__ bool strided() const { return gather == nullptr; }
    __device__ __forceinline__ const R* origin() const { return base; }
    __device__ __forceinline__ size_t outer_step() const { return 1; }
    __device__ __forceinline__ size_t red_step() const { return ld; }
    __device__ __forceinline__ int outer_limit() const { return 0x7fffffff; }
};

// Element (pixel = (b, oy, ox), k = (c, ky, kx)) of the receptive-field matrix of a mini-batch, with zero padding
// (Tensor.cpp:40-72 builds this matrix per sample on the CPU; here it only ever exists as this index map).
template <class R, bool PIXEL_IS_RED>
struct Im2colLoader {
    static constexpr bool kRedFast = PIXEL_IS_RED;
    const R* x; const int* gather; ConvGeom g;
    __device__ __forceinline__ R operator()(int o, int r) const {
        const int pix = PIXEL_IS_RED ? r : o;
        const int k = PIXEL_IS_RED ? o : r;
        const int b = pix / g.P, p = pix - b * g.P;
        const int oy = p / g.OW, ox = p - oy * g.OW;
        const int c = k / g.KK, q = k - c * g.KK;
        const int ky = q / g.KW, kx = q - ky * g.KW;
        const int iy = oy * g.sh - g.ph + ky, ix = ox * g.sw - g.pw + kx;
        if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) return R(0);
        const size_t s = gather ? (size_t)gather[b] : (size_t)b;
        return x[(s * g.C + c) * g.HW + iy * g.W + ix];
    }
};

template <class R>
struct GradPixelLoader {  // outer = filter f, red = (b, p):  G[b][f][p]
    static constexpr bool kRedFast = true;
    const R* G; int F, P;
    __device__ __forceinline__ R operator()(int f, int r) const {
        const int b = r / P, p = r - b * P;
        return G[((size_t)b * F + f) * P + p];
    }
};

template <class R>
struct FilterTLoader {  // outer = input channel c, red = (f, q=(ky,kx)):  filters[f][c][q]
    static constexpr bool kRedFast = true;
    const R* filt; int C, KK;
    __device__ __forceinline__ R operator()(int c, int r) const {
        const int f = r / KK, q = r - f * KK;
        return filt[((size_t)f * C + c) * KK + q];
    }
};

template <class R>
struct GradGatherLoader {  // outer = (b, y, x) input pixel, red = (f, ky, kx): the output-gradient pixel this tap reaches
    static constexpr bool kRedFast = false;
    const R* G; ConvGeom g;
    __device__ __forceinline__ R operator()(int o, int r) const {
        const int b = o / g.HW, yx = o - b * g.HW;
        const int y = yx / g.W, x = yx - y * g.W;
        const int f = r / g.KK, q = r - f * g.KK;
        const int ky = q / g.KW, kx = q - ky * g.KW;
        const int ty = y + g.ph - ky, tx = x + g.pw - kx;
        if (ty < 0 || tx < 0) return R(0);
        const int oy = ty / g.sh, ox = tx / g.sw;
        if (oy * g.sh != ty || ox * g.sw != tx || oy >= g.OH || ox >= g.OW) return R(0);
        return G[((size_t)b * g.F + f) * g.P + oy * g.OW + ox];
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Epilogues.
// ---------------------------------------------------------------------------------------------------------------
template <class R>
struct EpiFullForward {  // y[b][n] = act(acc + bias[n]);  deriv optional  (NLayer.inl:131-142)
    R* y; R* deriv; const R* bias; int N; int act;
    __device__ __forceinline__ void operator()(int b, int n, R acc) const {
        R d; const R v = apply_activation<R>(act, acc + bias[n], d);
        const size_t o = (size_t)b * N + n;
        y[o] = v;
        if (deriv) deriv[o] = d;
    }
    __device__ __forceinline__ bool row4_ok() const {
        return (N & 3) == 0 && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(deriv) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0;
    }
    __device__ __forceinline__ void row4(int b, int n, float4 a) const {
        const float4 bv = *reinterpret_cast<const float4*>(bias + n);
        float4 v, d;
        v.x = apply_activation<R>(act, a.x + bv.x, d.x); v.y = apply_activation<R>(act, a.y + bv.y, d.y);
        v.z = apply_activation<R>(act, a.z + bv.z, d.z); v.w = apply_activation<R>(act, a.w + bv.w, d.w);
        const size_t o = (size_t)b * N + n;
        *reinterpret_cast<float4*>(y + o) = v;
        if (deriv) *reinterpret_cast<float4*>(deriv + o) = d;
    }
};

template <class R>
struct EpiConvForward {  // i = filter f, j = (b, p): y[b][f][p] = act(acc + bias[f][p]) with the UNTIED bias (CLayer.inl:130)
    R* y; R* deriv; const R* bias; int F, P; int act;
    __device__ __forceinline__ void operator()(int f, int j, R acc) const {
        const int b = j / P, p = j - b * P;
        R d; const R v = apply_activation<R>(act, acc + bias[(size_t)f * P + p], d);
        const size_t o = ((size_t)b * F + f) * P + p;
        y[o] = v;
        if (deriv) deriv[o] = d;
    }
};

template <class R>
struct EpiAccumulate {  // out = scale != 0 ? scale * out + acc : acc   (gradient_scale_factor, NLayer.inl:158-164)
    R* out; size_t ld; R scale;
    __device__ __forceinline__ void operator()(int i, int j, R acc) const {
        const size_t o = (size_t)i * ld + j;
        out[o] = scale != R(0) ? acc + scale * out[o] : acc;
    }
};

template <class R>
struct EpiStore {
    R* out; size_t ld;
    __device__ __forceinline__ void operator()(int i, int j, R acc) const { out[(size_t)i * ld + j] = acc; }
    // optional row-segment protocol of the tensor-core epilogue (tc_gemm.cuh): four consecutive columns j .. j + 3 of row i at once
    __device__ __forceinline__ bool row4_ok() const { return (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0; }
    __device__ __forceinline__ void row4(int i, int j, float4 a) const { *reinterpret_cast<float4*>(out + (size_t)i * ld + j) = a; }
};

template <class R>
struct EpiConvInputGrad {  // i = channel c, j = (b, yx): dx[b][c][yx]
    R* dx; int C, HW;
    __device__ __forceinline__ void operator()(int c, int j, R acc) const {
        const int b = j / HW, yx = j - b * HW;
        dx[((size_t)b * C + c) * HW + yx] = acc;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Row kernels: softmax forward (ActivationFunction.inl:47-54, ActivationFunctionHelper.cpp:44-48) and its input
// gradient (ActivationFunctionHelper.cpp:50-58).  One warp per sample; warp-shuffle reductions.
// ---------------------------------------------------------------------------------------------------------------
template <class R> __device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <class R> __device__ __forceinline__ R warp_max(R v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const R t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}

template <class R>
__global__ void softmax_forward_kernel(R* __restrict__ zy, R* __restrict__ aux, int rows, int n) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= rows) return;
    R* z = zy + (size_t)row * n;
    R m = -real_limits<R>::max();
    for (int i = lane; i < n; i += 32) m = z[i] > m ? z[i] : m;
    m = warp_max<R>(m);
    R s = R(0);
    for (int i = lane; i < n; i += 32) { const R e = dl_exp<R>(z[i] - m); z[i] = e; s += e; }
    s = warp_sum<R>(s);
    const R factor = R(1) / s;
    for (int i = lane; i < n; i += 32) {
        const R e = z[i];
        if (aux) aux[(size_t)row * n + i] = e;   // un-normalised exponents are the stored auxiliary data
        z[i] = e * factor;
    }
}

template <class R>
__global__ void softmax_in_grad_kernel(const R* __restrict__ aux, const R* __restrict__ out_grad, R* __restrict__ res, int rows, int n) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= rows) return;
    const R* a = aux + (size_t)row * n;
    const R* g = out_grad + (size_t)row * n;
    R* r = res + (size_t)row * n;
    R s = R(0);
    for (int i = lane; i < n; i += 32) s += a[i];
    s = warp_sum<R>(s);
    const R ood = R(1) / s;
    R t = R(0);
    for (int i = lane; i < n; i += 32) { const R v = a[i] * g[i] * ood; r[i] = v; t += v; }
    t = warp_sum<R>(t);
    const R temp_sum = t * ood;
    for (int i = lane; i < n; i += 32) r[i] = r[i] - a[i] * temp_sum;
}

// g[b][e] = delta[b][e] * deriv[b][e] (AFunction::calc_in_grad, ActivationFunction.inl:102-106) written once, and
// its batch sum (the bias gradient, CLayer.inl:148-156 / NLayer.inl:156-160) accumulated in the same pass:
// partial[chunk][e] = sum over the chunk's samples.  deriv == nullptr: g already holds the activation gradient.
template <class R>
__global__ void act_grad_bias_partial_kernel(const R* __restrict__ delta, const R* __restrict__ deriv, R* __restrict__ g,
                                             R* __restrict__ partial, int B, int n, int chunk) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int b0 = blockIdx.y * chunk, b1 = min(B, b0 + chunk);
    R acc = R(0);
    for (int b = b0; b < b1; ++b) {
        const size_t o = (size_t)b * n + e;
        R v;
        if (deriv) { v = deriv[o] * delta[o]; g[o] = v; } else v = g[o];
        acc += v;
    }
    partial[(size_t)blockIdx.y * n + e] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// Pooling (PLayer.inl:115-168).  Forward: Tensor::min_max_pool<true> (Tensor.cpp:673-716; strict comparison, so the
// first element in (ky, kx) scan order wins ties; the map stores the flat index into the SAMPLE's input tensor) and
// scale_pool (Tensor.cpp:755-779).  Backward: the reference zero-fills and scatters through the map
// (Tensor.cpp:721-734, thrust::scatter on the GPU, CudaTensor.cu:751-761); here every INPUT element looks up the
// one window that covers it, so the pass is a single coalesced write with no memset and no scatter.
// ---------------------------------------------------------------------------------------------------------------
template <class R, class IDX>   // IDX = unsigned whenever the element count allows it: 64-bit divisions by run-time values cost ~100 instructions each
__global__ void pool_forward_kernel(const R* __restrict__ x, R* __restrict__ y, int* __restrict__ map, IDX total,
                                    int C, int H, int W, int wh, int ww, int OH, int OW, int type) {
    const IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int ox = (int)(e % (IDX)OW); IDX t = e / (IDX)OW;
    const int oy = (int)(t % (IDX)OH); t /= (IDX)OH;
    const int c = (int)(t % (IDX)C); const size_t b = t / (IDX)C;
    const R* xs = x + b * (size_t)C * H * W;
    if (type == DLB_POOL_AVERAGE) {
        R acc = R(0);
        for (int ky = 0; ky < wh; ++ky)
            for (int kx = 0; kx < ww; ++kx) acc += xs[((size_t)c * H + oy * wh + ky) * W + ox * ww + kx];
        y[e] = acc * (R(1) / R(wh * ww));
        return;
    }
    const bool is_max = type == DLB_POOL_MAX;
    R best = is_max ? -real_limits<R>::max() : real_limits<R>::max();
    int best_id = -1;
    for (int ky = 0; ky < wh; ++ky)
        for (int kx = 0; kx < ww; ++kx) {
            const int id = (c * H + oy * wh + ky) * W + ox * ww + kx;
            const R cur = xs[id];
            if (is_max ? (best < cur) : (best > cur)) { best = cur; best_id = id; }
        }
    y[e] = best;
    if (map) map[e] = best_id;
}

template <class R, class IDX>
__global__ void pool_backward_kernel(const R* __restrict__ delta, const int* __restrict__ map, R* __restrict__ dx, IDX total,
                                     int C, int H, int W, int wh, int ww, int OH, int OW, int type) {
    const IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int x = (int)(e % (IDX)W); IDX t = e / (IDX)W;
    const int y = (int)(t % (IDX)H); t /= (IDX)H;
    const int c = (int)(t % (IDX)C); const size_t b = t / (IDX)C;
    const int oy = y / wh, ox = x / ww;
    R v = R(0);
    if (oy < OH && ox < OW) {
        const size_t o = (b * C + c) * (size_t)OH * OW + (size_t)oy * OW + ox;
        if (type == DLB_POOL_AVERAGE) v = delta[o] * (R(1) / R(wh * ww));
        else if (map[o] == (c * H + y) * W + x) v = delta[o];
    }
    dx[e] = v;
}

// The bandwidth kernels for the shape every net of BASELINE.json pools with: 2 x 2 max / min windows on maps with an even height and a
// width that is a multiple of 4.  Rows of consecutive channels and samples are contiguous, so with H even the whole batch is ONE
// [rows][W] matrix whose row pairs (2p, 2p + 1) pool into row p of a [rows / 2][W / 2] matrix: no (sample, channel, y) decomposition
// except one modulo for the map's sample-relative index.  One thread = two neighbouring windows: two 16-byte (float) or four 16-byte
// (double) loads, one 8- / 16-byte store of the outputs and one 8-byte store of the map entries; backward is the mirror image
// (every input element written exactly once, no zero fill, no scatter).  Window scan order (ky, kx) and the strict comparison of
// Tensor::min_max_pool (Tensor.cpp:673-716) are kept, so the index map is the reference's bit for bit.
template <class R> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };
__device__ __forceinline__ void load4_stream(const float* p, float (&v)[4]) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4_stream(const double* p, double (&v)[4]) {
    const double2 a = __ldcs(reinterpret_cast<const double2*>(p)), b = __ldcs(reinterpret_cast<const double2*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void store4(double* p, const double (&v)[4]) {
    reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]); reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}

template <class R, bool IS_MAX>
__global__ void __launch_bounds__(256) pool2x2_forward_kernel(const R* __restrict__ x, R* __restrict__ y, int* __restrict__ map, unsigned pairs,
                                                              unsigned W, unsigned pairs_per_row, unsigned rows_per_sample) {
    const unsigned e = blockIdx.x * 256u + threadIdx.x;
    if (e >= pairs) return;
    const unsigned prow = e / pairs_per_row, px = e - prow * pairs_per_row;
    const R* r0 = x + (size_t)prow * 2 * W + px * 4;
    R a[4], b[4];
    load4_stream(r0, a); load4_stream(r0 + W, b);
    const int id0 = (int)((prow % rows_per_sample) * 2 * W + px * 4);   // index of a[0] inside its sample
    R out[2]; int ids[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        R best = IS_MAX ? -real_limits<R>::max() : real_limits<R>::max();
        int best_id = -1;
        const R cur[4] = {a[2 * j], a[2 * j + 1], b[2 * j], b[2 * j + 1]};
        const int id[4] = {id0 + 2 * j, id0 + 2 * j + 1, id0 + (int)W + 2 * j, id0 + (int)W + 2 * j + 1};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (IS_MAX ? (best < cur[k]) : (best > cur[k])) { best = cur[k]; best_id = id[k]; }
        out[j] = best; ids[j] = best_id;
    }
    const size_t o = (size_t)prow * (2 * pairs_per_row) + px * 2;
    typename Vec2<R>::type yo; yo.x = out[0]; yo.y = out[1];
    *reinterpret_cast<typename Vec2<R>::type*>(y + o) = yo;
    if (map) *reinterpret_cast<int2*>(map + o) = make_int2(ids[0], ids[1]);
}

template <class R>
__global__ void __launch_bounds__(256) pool2x2_backward_kernel(const R* __restrict__ delta, const int* __restrict__ map, R* __restrict__ dx, unsigned pairs,
                                                               unsigned W, unsigned pairs_per_row, unsigned rows_per_sample) {
    const unsigned e = blockIdx.x * 256u + threadIdx.x;
    if (e >= pairs) return;
    const unsigned prow = e / pairs_per_row, px = e - prow * pairs_per_row;
    const size_t o = (size_t)prow * (2 * pairs_per_row) + px * 2;
    const typename Vec2<R>::type d = __ldcs(reinterpret_cast<const typename Vec2<R>::type*>(delta + o));
    const int2 m = __ldcs(reinterpret_cast<const int2*>(map + o));
    const int id0 = (int)((prow % rows_per_sample) * 2 * W + px * 4);
    const int id1 = id0 + (int)W;
    const R a[4] = {m.x == id0 ? d.x : R(0), m.x == id0 + 1 ? d.x : R(0), m.y == id0 + 2 ? d.y : R(0), m.y == id0 + 3 ? d.y : R(0)};
    const R b[4] = {m.x == id1 ? d.x : R(0), m.x == id1 + 1 ? d.x : R(0), m.y == id1 + 2 ? d.y : R(0), m.y == id1 + 3 ? d.y : R(0)};
    R* r0 = dx + (size_t)prow * 2 * W + px * 4;
    store4(r0, a); store4(r0 + W, b);
}

// ---------------------------------------------------------------------------------------------------------------
// Cost derivative (CostFunction.inl:47-51 -> CostFunctionHelper.cpp:52-57).  The reference obtains it by forward-mode
// automatic differentiation (Dual.h) of the lambdas in CostFunctionFactory.h:29-55; the dual arithmetic is expanded
// here term by term, including Func::nan_to_num (Functions.h:65-78) whose saturated branches have derivative 0.
// ---------------------------------------------------------------------------------------------------------------
template <class R> __device__ __forceinline__ bool not_finite(R v) { return isinf(v) || isnan(v); }

template <class R>
__device__ __forceinline__ R cost_derivative(int id, R x, R ref) {
    switch (id) {
    case DLB_COST_SQUARED_ERROR: { const R diff = x - ref; const R h = R(0.5) * diff; return R(1) * h + R(0.5) * diff; }
    case DLB_COST_CROSS_ENTROPY: {
        if (ref <= R(0)) {
            const R a = R(1) - x;
            const R da = R(-1) * (R(1) / a);
            const R val = -dl_log<R>(a), d = da * R(-1);
            return (not_finite(val) || not_finite(d)) ? R(0) : d;
        }
        if (ref >= R(1)) {
            const R dl = R(1) * (R(1) / x);
            const R val = -dl_log<R>(x), d = dl * R(-1);
            return (not_finite(val) || not_finite(d)) ? R(0) : d;
        }
        const R lx = dl_log<R>(x), dlx = R(1) * (R(1) / x);
        const R a = R(1) - x, la = dl_log<R>(a), dla = R(-1) * (R(1) / a);
        const R omr = R(1) - ref;
        const R v1 = ref * lx, d1 = dlx * ref + R(0) * lx;
        const R v2 = omr * la, d2 = dla * omr + R(0) * la;
        const R val = -(v1 + v2), d = (d1 + d2) * R(-1);
        return (not_finite(val) || not_finite(d)) ? R(0) : d;
    }
    case DLB_COST_LINEAR: return R(1);
    default: return R(0) / R(0);
    }
}

template <class R> __device__ __forceinline__ R nan_to_num(R v) {
    if (isinf(v)) return v < R(0) ? -real_limits<R>::max() : real_limits<R>::max();
    if (isnan(v)) return R(0);
    return v;
}

template <class R>
__device__ __forceinline__ R cost_value(int id, R x, R ref) {  // CostFunctionFactory.h:33-53
    switch (id) {
    case DLB_COST_SQUARED_ERROR: { const R diff = x - ref; return R(0.5) * diff * diff; }
    case DLB_COST_CROSS_ENTROPY:
        if (ref <= R(0)) return nan_to_num<R>(-dl_log<R>(R(1) - x));
        if (ref >= R(1)) return nan_to_num<R>(-dl_log<R>(x));
        return nan_to_num<R>(-(ref * dl_log<R>(x) + (R(1) - ref) * dl_log<R>(R(1) - x)));
    case DLB_COST_LINEAR: return x;
    default: return R(0) / R(0);
    }
}

template <class R>
__global__ void cost_gradient_kernel(const R* __restrict__ out, const R* __restrict__ ref, const int* __restrict__ gather,
                                     R* __restrict__ grad, size_t total, int n, int cost_id) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const size_t b = e / n; const int j = (int)(e - b * n);
    const size_t row = gather ? (size_t)gather[b] : b;
    grad[e] = cost_derivative<R>(cost_id, out[e], ref[row * n + j]);
}

// Sum over the mini-batch of the per-sample cost (CostFunctionHelper.cpp:27-35), one block, result in double.
template <class R>
__global__ void cost_sum_kernel(const R* __restrict__ out, const R* __restrict__ ref, const int* __restrict__ gather,
                                int B, int n, int cost_id, double* __restrict__ result) {
    __shared__ double part[256];
    double acc = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const size_t row = gather ? (size_t)gather[b] : (size_t)b;
        R s = R(0);
        for (int j = 0; j < n; ++j) s += cost_value<R>(cost_id, out[(size_t)b * n + j], ref[row * n + j]);
        acc += (double)s;
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *result = part[0];
}

// Sum of squares of a range (BasicCollection::sum_of_squares), one block, result accumulated into *result in double.
template <class R>
__global__ void sum_of_squares_kernel(const R* __restrict__ v, size_t n, double* __restrict__ result) {
    __shared__ double part[256];
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += (double)v[i] * (double)v[i];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(result, part[0]);
}

// Number of rows whose arg-max equals the arg-max of the reference row (Net.inl:430-436), accumulated into *result.
template <class R>
__global__ void count_correct_kernel(const R* __restrict__ y, const R* __restrict__ ref, int rows, int n, int* __restrict__ result) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const R* p = y + (size_t)r * n;
    const R* q = ref + (size_t)r * n;
    int bp = 0, bq = 0;
    for (int j = 1; j < n; ++j) { if (p[bp] < p[j]) bp = j; if (q[bq] < q[j]) bq = j; }
    if (bp == bq) atomicAdd(result, 1);
}

// Per-row argmax (BasicCollection::max_element_id, first maximum), used by evaluation (Net.inl:432-433).
template <class R>
__global__ void argmax_rows_kernel(const R* __restrict__ y, int* __restrict__ cls, int rows, int n) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const R* p = y + (size_t)r * n;
    int best = 0;
    for (int j = 1; j < n; ++j) if (p[best] < p[j]) best = j;
    cls[r] = best;
}

// ---------------------------------------------------------------------------------------------------------------
// SGD update of the whole flat parameter buffer in one launch (ALayer.inl:97-122 applied to every layer,
// Net.inl:309-316):  regularisable containers  w <- w * param_scale + g * grad_scale   (scale_and_add_scaled)
//                    biases (or lambda == 0)   b <- b + g * grad_scale                  (add_scaled)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxSegments = 64;
struct UpdateSegments {
    int count;
    long long end[kMaxSegments];        // exclusive end offset of each segment in the flat buffer
    unsigned char regularised[kMaxSegments];
};

template <class R>
__global__ void sgd_update_kernel(R* __restrict__ p, const R* __restrict__ g, size_t n, UpdateSegments segs,
                                  R param_scale, R grad_scale, int nontrivial_reg) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int s = 0;
    while (s < segs.count - 1 && (long long)e >= segs.end[s]) ++s;
    const bool reg = nontrivial_reg && segs.regularised[s];
    const R x = p[e], y = g[e];
    p[e] = reg ? x * param_scale + y * grad_scale : x + y * grad_scale;
}

// ---------------------------------------------------------------------------------------------------------------
// Typed host-side launchers.
// ---------------------------------------------------------------------------------------------------------------
// ApplyDropout (ALayer.inl:46-55) over a batch: out[b][i] = in[row(b)][i] * (mask ? mask[i] : scale); in place when out == in.
template <class R>
__global__ void dropout_apply_kernel(const R* in, const int* __restrict__ gather, const R* __restrict__ mask, R scale,
                                     R* out, size_t total, int n) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const size_t b = e / (size_t)n; const int i = (int)(e - b * (size_t)n);
    const size_t row = gather ? (size_t)gather[b] : b;
    out[e] = in[row * (size_t)n + i] * (mask ? mask[i] : scale);
}

struct LaunchCtx {
    cudaStream_t stream;
    GemmWorkspace ws;
    bool tc = false;     // Real = float only: contractions on the tensor cores (tc_gemm.cuh, 3xTF32) instead of the FMA path
    long long gather_rows = 0;   // rows of the array a layer-0 gather indexes into (the resident data set); 0: unknown
};

// ---------------------------------------------------------------------------------------------------------------
// Implicit-GEMM convolution on the tensor cores (Real = float).  The receptive-field matrix of the mini-batch
// (Tensor.cpp:40-72 builds it per sample on the CPU) only exists as the index map of the producers (tc_tiles.cuh);
// the 128-row MMA dimension is always the long one (pixels, or C*KH*KW for the filter gradient) and the MMA's N is the
// filter / channel count rounded up to 16, so that 32- and 64-filter layers do not pay for 128 columns.
//   forward          Y[(b,p)][f]   = sum_k  patch[(b,p)][k] * filters[f][k]            M = B*P,   N = F, K = C*KH*KW
//   filter gradient  dK[k][f]      = sum_(b,p) patch[(b,p)][k] * G[b][f][p]            M = C*KH*KW, N = F, K = B*P (split)
//   input gradient   dX[(b,yx)][c] = sum_(f,q) G-patch[(b,yx)][(f,q)] * filtersT[c][(f,q)]   M = B*H*W, N = C, K = F*KH*KW
// The input gradient gathers the output gradient through the mirrored window (stride 1 only; other strides stay on the
// FMA path) and reads the filters through a channel-major copy made by filters_channel_major_kernel.
// ---------------------------------------------------------------------------------------------------------------
#ifndef DLB_CUDA_EMU
struct EpiConvForwardPix {  // i = (b, p), j = f: y[b][f][p] = act(acc + bias[f][p]) with the UNTIED bias (CLayer.inl:130)
    static constexpr bool kLaneIsRow = true;
    float* y; float* deriv; const float* bias; int F, P; int act;
    __device__ __forceinline__ void operator()(int i, int f, float acc) const {
        const int b = i / P, p = i - b * P;
        float d; const float v = apply_activation<float>(act, acc + bias[(size_t)f * P + p], d);
        const size_t o = ((size_t)b * F + f) * P + p;
        y[o] = v;
        if (deriv) deriv[o] = d;
    }
};

struct EpiConvInputGradPix {  // i = (b, yx), j = channel c: dx[b][c][yx]
    static constexpr bool kLaneIsRow = true;
    float* dx; int C, HW;
    __device__ __forceinline__ void operator()(int i, int c, float acc) const {
        const int b = i / HW, yx = i - b * HW;
        dx[((size_t)b * C + c) * HW + yx] = acc;
    }
};

struct EpiAccumulateT {  // out[j][i] = scale != 0 ? scale * out[j][i] + acc : acc   (the transposed EpiAccumulate)
    float* out; size_t ld; float scale;
    __device__ __forceinline__ void operator()(int i, int j, float acc) const {
        const size_t o = (size_t)j * ld + i;
        out[o] = scale != 0.f ? acc + scale * out[o] : acc;
    }
};

// filters[f][c][q] -> out[c][f * KK + q]
static __global__ void filters_channel_major_kernel(const float* __restrict__ filt, float* __restrict__ out, int F, int C, int KK) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= F * C * KK) return;
    const int q = e % KK, fc = e / KK;
    const int c = fc % C, f = fc / C;
    out[((size_t)c * F + f) * KK + q] = filt[e];
}

// Ranges the producers' 32-bit offsets, 15-bit tap masks and reciprocal divisions are exact for (tc_tiles.cuh).
inline bool tc_conv_ok(const ConvGeom& g, int B, const int* gather = nullptr, long long gather_rows = 0) {
    if (gather && (gather_rows <= 0 || gather_rows * g.in_n >= (1LL << 31))) return false;   // gathered sample offsets are 32-bit
    return g.KH <= 15 && g.KW <= 15 && g.P > 0 && g.H <= 2048 && g.W <= 2048 && g.OH <= 2048 && g.OW <= 2048 &&
           (long long)g.F * g.KK < 65536 && g.K < 65536 &&
           (long long)B * std::max(g.in_n, g.out_n) < (1LL << 31);
}

inline void tc_conv_forward(const LaunchCtx& lc, int B, const ConvGeom& g, int act, const float* x, const int* gather, const int* cursor,
                            const float* bias, const float* filters, float* y, float* deriv) {
    const TcOperand A{x, 0, kPatchOuter, gather, cursor, -1, make_patch(g.C, g.H, g.W, g.OH, g.OW, g.KH, g.KW, g.sh, g.sw, -g.ph, -g.pw, 1)};
    const TcOperand Bo{filters, (long long)g.K, kRedContig, nullptr, nullptr, -1, {}};
    tc_gemm<kPatchOuter, kRedContig>(lc.stream, B * g.P, g.F, g.K, A, Bo, EpiConvForwardPix{y, deriv, bias, g.F, g.P, act}, lc.ws);
}

inline void tc_conv_filter_gradient(const LaunchCtx& lc, int B, const ConvGeom& g, const float* x, const int* gather, const int* cursor,
                                    const float* grad, float* dfilters, float scale) {
    const TcOperand A{x, 0, kPatchRed, gather, cursor, -1, make_patch(g.C, g.H, g.W, g.OH, g.OW, g.KH, g.KW, g.sh, g.sw, -g.ph, -g.pw, 1)};
    const TcOperand Bo{grad, 0, kPixelRed, nullptr, nullptr, -1, make_patch(g.F, g.OH, g.OW, g.OH, g.OW, 1, 1, 1, 1, 0, 0, 1)};
    tc_gemm<kPatchRed, kPixelRed>(lc.stream, g.K, g.F, B * g.P, A, Bo, EpiAccumulateT{dfilters, (size_t)g.K, scale}, lc.ws);
}

inline void tc_conv_input_gradient(const LaunchCtx& lc, int B, const ConvGeom& g, const float* grad, const float* filters, float* dx) {
    // channel-major filter copy at the tail of the scratch; the contraction's split partials use the rest
    const size_t wt_bytes = ((size_t)g.F * g.K * sizeof(float) + 255) & ~(size_t)255;
    float* wt = reinterpret_cast<float*>(static_cast<unsigned char*>(lc.ws.ptr) + (lc.ws.bytes - wt_bytes));
    DLB_LAUNCH((filters_channel_major_kernel), dim3((unsigned)ceil_div(g.F * g.K, 256)), dim3(256), lc.stream, filters, wt, g.F, g.C, g.KK);
    const TcOperand A{grad, 0, kPatchOuter, nullptr, nullptr, -1, make_patch(g.F, g.OH, g.OW, g.H, g.W, g.KH, g.KW, 1, 1, g.ph, g.pw, -1)};
    const TcOperand Bo{wt, (long long)g.F * g.KK, kRedContig, nullptr, nullptr, -1, {}};
    tc_gemm<kPatchOuter, kRedContig>(lc.stream, B * g.HW, g.C, g.F * g.KK, A, Bo, EpiConvInputGradPix{dx, g.C, g.HW}, GemmWorkspace{lc.ws.ptr, lc.ws.bytes - wt_bytes});
}
#endif

template <class R>
void softmax_forward(const LaunchCtx& lc, R* zy, R* aux, int rows, int n) {
    TagScope tag_("softmax_fwd");
    DLB_LAUNCH((softmax_forward_kernel<R>), dim3(ceil_div(rows, 4)), dim3(128), lc.stream, zy, aux, rows, n);
}

// y = act(W x + b) over the batch.  x rows may be gathered (layer 0 reads the resident data set through the
// step's slice of the epoch permutation).
template <class R>
void full_forward(const LaunchCtx& lc, int B, int in_n, int out_n, int act, const R* x, const int* gather,
                  const R* bias, const R* W, R* y, R* deriv) {
    {
        TagScope tag_("full_fwd");
        bool done = false;
#ifndef DLB_CUDA_EMU
        if constexpr (std::is_same_v<R, float>) {
            if (lc.tc) {
                tc_gemm<kRedContig, kRedContig>(lc.stream, B, out_n, in_n, TcOperand{x, (long long)in_n, kRedContig, gather, nullptr, -1, {}},
                        TcOperand{W, (long long)in_n, kRedContig, nullptr, nullptr, -1, {}},
                        EpiFullForward<float>{y, act == DLB_ACT_SOFTMAX ? nullptr : deriv, bias, out_n, act}, lc.ws);
                done = true;
            }
        }
#endif
        if (!done)
            gemm<R>(lc.stream, B, out_n, in_n, DenseRowLoader<R>{x, (size_t)in_n, gather}, DenseRowLoader<R>{W, (size_t)in_n, nullptr},
                    EpiFullForward<R>{y, act == DLB_ACT_SOFTMAX ? nullptr : deriv, bias, out_n, act}, lc.ws);
    }
    if (act == DLB_ACT_SOFTMAX) softmax_forward<R>(lc, y, deriv, B, out_n);
}

template <class R>
void conv_forward(const LaunchCtx& lc, int B, const ConvGeom& g, int act, const R* x, const int* gather,
                  const R* bias, const R* filters, R* y, R* deriv) {
    {
        TagScope tag_("conv_fwd");
        bool done = false;
#ifndef DLB_CUDA_EMU
        if constexpr (std::is_same_v<R, float>) {
            if (lc.tc && tc_conv_ok(g, B, gather, lc.gather_rows)) {
                tc_conv_forward(lc, B, g, act, x, gather, nullptr, bias, filters, y, act == DLB_ACT_SOFTMAX ? nullptr : deriv);
                done = true;
            }
        }
#endif
        if (!done)
            gemm<R>(lc.stream, g.F, B * g.P, g.K, DenseRowLoader<R>{filters, (size_t)g.K, nullptr}, Im2colLoader<R, false>{x, gather, g},
                    EpiConvForward<R>{y, act == DLB_ACT_SOFTMAX ? nullptr : deriv, bias, g.F, g.P, act}, lc.ws);
    }
    if (act == DLB_ACT_SOFTMAX) softmax_forward<R>(lc, y, deriv, B, g.out_n);
}

// g = dC/dz of the layer (activation gradient applied to the incoming deltas) and the bias gradient
// dbias = scale * dbias + sum_b g[b].
template <class R>
void activation_grad_and_bias(const LaunchCtx& lc, int B, int n, int act, const R* delta, const R* deriv, R* g, R* dbias, R scale) {
    TagScope tag_("act_grad_bias");
    constexpr int kChunk = 16;
    if (act == DLB_ACT_SOFTMAX)
        DLB_LAUNCH((softmax_in_grad_kernel<R>), dim3(ceil_div(B, 4)), dim3(128), lc.stream, deriv, delta, g, B, n);
    int chunk = kChunk;
    int chunks = ceil_div(B, chunk);
    const size_t cap = lc.ws.bytes / ((size_t)n * sizeof(R));
    if ((size_t)chunks > cap) { chunks = (int)(cap > 0 ? cap : 1); chunk = ceil_div(B, chunks); chunks = ceil_div(B, chunk); }
    R* partial = static_cast<R*>(lc.ws.ptr);
    DLB_LAUNCH((act_grad_bias_partial_kernel<R>), dim3(ceil_div(n, 128), chunks), dim3(128), lc.stream,
               delta, act == DLB_ACT_SOFTMAX ? (const R*)nullptr : deriv, g, partial, B, n, chunk);
    DLB_LAUNCH((gemm_finalize_kernel<R, EpiAccumulate<R>>), dim3(ceil_div(n, 256)), dim3(256), lc.stream,
               (const R*)partial, chunks, 1, n, EpiAccumulate<R>{dbias, (size_t)n, scale});
}

// NLayer<D>::backpropagate (NLayer.inl:144-171) over the batch.  dx may be nullptr (layer 0, Net.inl:272).
template <class R>
void full_backward(const LaunchCtx& lc, int B, int in_n, int out_n, int act, const R* x, const int* gather, const R* delta,
                   const R* deriv, const R* W, R* g, R* dbias, R* dW, R* dx, R scale) {
    activation_grad_and_bias<R>(lc, B, out_n, act, delta, deriv, g, dbias, scale);
#ifndef DLB_CUDA_EMU
    if constexpr (std::is_same_v<R, float>) {
        if (lc.tc) {
            {   // dW[n][k] = sum_b g[b][n] x[b][k]: both operands outer-contiguous
                TagScope tag_("full_dW");
                tc_gemm<kOuterContig, kOuterContig>(lc.stream, out_n, in_n, B, TcOperand{g, (long long)out_n, kOuterContig, nullptr, nullptr, -1, {}},
                        TcOperand{x, (long long)in_n, kOuterContig, gather, nullptr, -1, {}}, EpiAccumulate<float>{dW, (size_t)in_n, scale}, lc.ws);
            }
            if (dx) {   // dx[b][k] = sum_n g[b][n] W[n][k]
                TagScope tag_("full_dX");
                tc_gemm<kRedContig, kOuterContig>(lc.stream, B, in_n, out_n, TcOperand{g, (long long)out_n, kRedContig, nullptr, nullptr, -1, {}},
                        TcOperand{W, (long long)in_n, kOuterContig, nullptr, nullptr, -1, {}}, EpiStore<float>{dx, (size_t)in_n}, lc.ws);
            }
            return;
        }
    }
#endif
    {
        TagScope tag_("full_dW");
        gemm<R>(lc.stream, out_n, in_n, B, DenseColLoader<R>{g, (size_t)out_n, nullptr}, DenseColLoader<R>{x, (size_t)in_n, gather},
                EpiAccumulate<R>{dW, (size_t)in_n, scale}, lc.ws);
    }
    if (dx) {
        TagScope tag_("full_dX");
        gemm<R>(lc.stream, B, in_n, out_n, DenseRowLoader<R>{g, (size_t)out_n, nullptr}, DenseColLoader<R>{W, (size_t)in_n, nullptr},
                EpiStore<R>{dx, (size_t)in_n}, lc.ws);
    }
}

// CLayer<D>::backpropagate (CLayer.inl:139-197) over the batch.
template <class R>
void conv_backward(const LaunchCtx& lc, int B, const ConvGeom& cg, int act, const R* x, const int* gather, const R* delta,
                   const R* deriv, const R* filters, R* g, R* dbias, R* dfilters, R* dx, R scale) {
    activation_grad_and_bias<R>(lc, B, cg.out_n, act, delta, deriv, g, dbias, scale);
    bool tc = false;
#ifndef DLB_CUDA_EMU
    if constexpr (std::is_same_v<R, float>) tc = lc.tc && tc_conv_ok(cg, B, gather, lc.gather_rows);
#endif
    {
        TagScope tag_("conv_dK");
        bool done = false;
#ifndef DLB_CUDA_EMU
        if constexpr (std::is_same_v<R, float>) {
            if (tc) { tc_conv_filter_gradient(lc, B, cg, x, gather, nullptr, g, dfilters, scale); done = true; }
        }
#endif
        if (!done)
            gemm<R>(lc.stream, cg.F, cg.K, B * cg.P, GradPixelLoader<R>{g, cg.F, cg.P}, Im2colLoader<R, true>{x, gather, cg},
                    EpiAccumulate<R>{dfilters, (size_t)cg.K, scale}, lc.ws);
    }
    if (dx) {
        TagScope tag_("conv_dX");
        bool done = false;
#ifndef DLB_CUDA_EMU
        if constexpr (std::is_same_v<R, float>) {
            if (tc && cg.sh == 1 && cg.sw == 1) { tc_conv_input_gradient(lc, B, cg, g, filters, dx); done = true; }
        }
#endif
        if (!done)
            gemm<R>(lc.stream, cg.C, B * cg.HW, cg.F * cg.KK, FilterTLoader<R>{filters, cg.C, cg.KK}, GradGatherLoader<R>{g, cg},
                    EpiConvInputGrad<R>{dx, cg.C, cg.HW}, lc.ws);
    }
}

// out[c][r] = in[r][c] for a row-major [rows][cols] matrix (32 x 32 tiles through padded shared memory, both sides coalesced).
// Used once per step for the transposed copy of a hidden NLayer's weights that feeds the input-gradient contraction K-major (dlb200.cu).
template <class R>
__global__ void __launch_bounds__(256) transpose_kernel(const R* __restrict__ in, R* __restrict__ out, int rows, int cols) {
    __shared__ R tile[32][33];
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = ty; j < 32; j += 8)
        if (r0 + j < rows && c0 + tx < cols) tile[j][tx] = in[(size_t)(r0 + j) * cols + c0 + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
        if (c0 + j < cols && r0 + tx < rows) out[(size_t)(c0 + j) * rows + r0 + tx] = tile[tx][j];
}

// 2 x 2 max / min pooling of an even-height, multiple-of-4-width map through 16-byte aligned pointers: the vectorised kernels apply
template <class R>
inline bool pool2x2_applies(int B, const dlb_layer_spec& s, const void* big, const void* small_a, const void* small_b) {
    if (s.win_h != 2 || s.win_w != 2 || (s.in_h & 1) || (s.in_w & 3) || s.pool_type == DLB_POOL_AVERAGE) return false;
    if ((size_t)s.in_c * s.in_h * s.in_w >= (1u << 31) || (size_t)B * s.in_c * s.in_h * s.in_w / 8 >= (1ull << 32)) return false;
    const uintptr_t bits = reinterpret_cast<uintptr_t>(big) | reinterpret_cast<uintptr_t>(small_a) | reinterpret_cast<uintptr_t>(small_b);
    return (bits & 15) == 0;
}

template <class R>
void pool_forward(const LaunchCtx& lc, int B, const dlb_layer_spec& s, const R* x, R* y, int* map) {
    TagScope tag_("pool_fwd");
    const int OH = conv_out_size(s.in_h, s.win_h, 0, s.win_h), OW = conv_out_size(s.in_w, s.win_w, 0, s.win_w);
    const size_t total = (size_t)B * s.in_c * OH * OW;
    if (!total) return;
    if (pool2x2_applies<R>(B, s, x, y, map)) {
        const unsigned pairs = (unsigned)(total / 2);
        const dim3 grid((unsigned)ceil_div<size_t>(pairs, 256));
        if (s.pool_type == DLB_POOL_MAX)
            DLB_LAUNCH((pool2x2_forward_kernel<R, true>), grid, dim3(256), lc.stream, x, y, map, pairs, (unsigned)s.in_w, (unsigned)(s.in_w / 4), (unsigned)(s.in_c * OH));
        else
            DLB_LAUNCH((pool2x2_forward_kernel<R, false>), grid, dim3(256), lc.stream, x, y, map, pairs, (unsigned)s.in_w, (unsigned)(s.in_w / 4), (unsigned)(s.in_c * OH));
        return;
    }
    if (total < (1ull << 31))
        DLB_LAUNCH((pool_forward_kernel<R, unsigned>), dim3((unsigned)ceil_div<size_t>(total, 256)), dim3(256), lc.stream, x, y, map, (unsigned)total,
                   s.in_c, s.in_h, s.in_w, s.win_h, s.win_w, OH, OW, s.pool_type);
    else
        DLB_LAUNCH((pool_forward_kernel<R, size_t>), dim3((unsigned)ceil_div<size_t>(total, 256)), dim3(256), lc.stream, x, y, map, total,
                   s.in_c, s.in_h, s.in_w, s.win_h, s.win_w, OH, OW, s.pool_type);
}

template <class R>
void pool_backward(const LaunchCtx& lc, int B, const dlb_layer_spec& s, const R* delta, const int* map, R* dx) {
    TagScope tag_("pool_bwd");
    const int OH = conv_out_size(s.in_h, s.win_h, 0, s.win_h), OW = conv_out_size(s.in_w, s.win_w, 0, s.win_w);
    const size_t total = (size_t)B * s.in_c * s.in_h * s.in_w;
    if (!total) return;
    if (pool2x2_applies<R>(B, s, dx, delta, map)) {
        const unsigned pairs = (unsigned)(total / 8);
        DLB_LAUNCH((pool2x2_backward_kernel<R>), dim3((unsigned)ceil_div<size_t>(pairs, 256)), dim3(256), lc.stream, delta, map, dx, pairs,
                   (unsigned)s.in_w, (unsigned)(s.in_w / 4), (unsigned)(s.in_c * OH));
        return;
    }
    if (total < (1ull << 31))
        DLB_LAUNCH((pool_backward_kernel<R, unsigned>), dim3((unsigned)ceil_div<size_t>(total, 256)), dim3(256), lc.stream, delta, map, dx, (unsigned)total,
                   s.in_c, s.in_h, s.in_w, s.win_h, s.win_w, OH, OW, s.pool_type);
    else
        DLB_LAUNCH((pool_backward_kernel<R, size_t>), dim3((unsigned)ceil_div<size_t>(total, 256)), dim3(256), lc.stream, delta, map, dx, total,
                   s.in_c, s.in_h, s.in_w, s.win_h, s.win_w, OH, OW, s.pool_type);
}

template <class R>
void cost_gradient(const LaunchCtx& lc, int B, int n, int cost_id, const R* out, const R* ref, const int* gather, R* grad) {
    TagScope tag_("cost_grad");
    const size_t total = (size_t)B * n;
    DLB_LAUNCH((cost_gradient_kernel<R>), dim3((unsigned)ceil_div<size_t>(total, 256)), dim3(256), lc.stream, out, ref, gather, grad, total, n, cost_id);
}

template <class R>
void sgd_update(const LaunchCtx& lc, R* params, const R* grads, size_t n, const UpdateSegments& segs, R param_scale, R grad_scale, bool nontrivial_reg) {
    TagScope tag_("sgd_update");
    if (!n) return;
    DLB_LAUNCH((sgd_update_kernel<R>), dim3((unsigned)ceil_div<size_t>(n, 256)), dim3(256), lc.stream, params, grads, n, segs,
               param_scale, grad_scale, nontrivial_reg ? 1 : 0);
}

}  // namespace dlb

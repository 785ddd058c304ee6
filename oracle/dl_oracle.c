/*
 * oracle/dl_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see dl_oracle.h for the rules and the pinning).
 *
 * Plain-C, per-sample restatement of the reference's training path.  The scalar (non-USE_AVX2) branches of the
 * reference are followed literally, so the order of floating-point operations is the reference's own; the only
 * differences to the AVX2 build of oracle/_ref are at rounding level (lane-wise partial sums, FMA contraction).
 */
#include "dl_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef orc_real R;

#ifdef ORACLE_SINGLE_PRECISION
#define R_MAX FLT_MAX
#define R_EXP expf
#define R_LOG logf
#define R_TANH tanhf
#else
#define R_MAX DBL_MAX
#define R_EXP exp
#define R_LOG log
#define R_TANH tanh
#endif

static _Thread_local char g_err[256];
static int fail(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return 1; }

typedef struct {
    dlb_layer_spec s;
    int out_c, out_h, out_w;
    int64_t in_n, out_n;
    int64_t param_off, bias_n, weight_n; /* weight_n: all weights of the layer (F filters together) */
    R keep_rate;                         /* ALayer::_keep_rate, 1 = no dropout (ALayer.h:59) */
} orc_layer;

struct orc_net {
    int n_layers;
    orc_layer* layers;
    int64_t n_params;
    int64_t act_total, in_total;
    R* params;
    const R** mask;   /* per layer: the current batch's keep mask over the layer input (ALayer::_keep_mask), or NULL */
};

int orc_real_bytes(void) { return (int)sizeof(R); }
const char* orc_last_error(void) { return g_err; }

/* ConvolutionUtils.cpp:29-37 calc_out_size_for_convolution */
static int conv_out_size(int in, int k, int pad, int stride) {
    const long long t = (long long)in + 2LL * pad - k;
    if (t < 0) return 0;
    return (int)(t / stride + 1);
}

orc_net* orc_net_create(const dlb_layer_spec* specs, int n_layers) {
    if (!specs || n_layers <= 0) { fail("orc_net_create: empty layer list"); return NULL; }
    orc_net* net = (orc_net*)calloc(1, sizeof *net);
    net->n_layers = n_layers;
    net->layers = (orc_layer*)calloc((size_t)n_layers, sizeof(orc_layer));
    int64_t off = 0;
    for (int l = 0; l < n_layers; ++l) {
        orc_layer* L = &net->layers[l];
        L->s = specs[l];
        const dlb_layer_spec* s = &L->s;
        L->in_n = (int64_t)s->in_c * s->in_h * s->in_w;
        L->param_off = off;
        int ok = L->in_n > 0;
        switch (s->type) {
        case DLB_LAYER_FULL: /* NLayer.inl:116-129: in {1,1,cols}, out {1,1,rows} */
            L->out_c = 1; L->out_h = 1; L->out_w = s->out_c;
            L->bias_n = s->out_c; L->weight_n = (int64_t)s->out_c * L->in_n;
            ok = ok && s->out_c > 0 && s->activation >= DLB_ACT_SIGMOID && s->activation <= DLB_ACT_LINEAR;
            break;
        case DLB_LAYER_CONVOLUTION: /* CLayer.inl:29-48 */
            L->out_c = s->out_c;
            L->out_h = conv_out_size(s->in_h, s->win_h, s->pad_h, s->stride_h > 0 ? s->stride_h : 1);
            L->out_w = conv_out_size(s->in_w, s->win_w, s->pad_w, s->stride_w > 0 ? s->stride_w : 1);
            L->bias_n = (int64_t)L->out_c * L->out_h * L->out_w;
            L->weight_n = (int64_t)s->out_c * s->in_c * s->win_h * s->win_w;
            ok = ok && s->out_c > 0 && s->win_h > 0 && s->win_w > 0 && s->stride_h > 0 && s->stride_w > 0 &&
                 s->pad_h >= 0 && s->pad_w >= 0 && s->activation >= DLB_ACT_SIGMOID && s->activation <= DLB_ACT_LINEAR;
            break;
        case DLB_LAYER_PULL: /* PLayer.inl:27-33, 96-99: stride == window, no padding */
            L->out_c = s->in_c;
            L->out_h = s->win_h > 0 ? conv_out_size(s->in_h, s->win_h, 0, s->win_h) : 0;
            L->out_w = s->win_w > 0 ? conv_out_size(s->in_w, s->win_w, 0, s->win_w) : 0;
            L->bias_n = 0; L->weight_n = 0;
            ok = ok && s->win_h > 0 && s->win_w > 0 && s->pool_type >= DLB_POOL_MAX && s->pool_type <= DLB_POOL_AVERAGE;
            break;
        default: ok = 0;
        }
        L->out_n = (int64_t)L->out_c * L->out_h * L->out_w;
        L->keep_rate = (R)1;
        /* Net.inl:83-88 consistency of consecutive layers */
        if (ok && l > 0) {
            const orc_layer* P = &net->layers[l - 1];
            if (s->type == DLB_LAYER_FULL) ok = P->out_n == L->in_n;
            else ok = P->out_c == s->in_c && P->out_h == s->in_h && P->out_w == s->in_w;
        }
        if (!ok || L->out_n <= 0) {
            snprintf(g_err, sizeof g_err, "orc_net_create: invalid or inconsistent spec at layer %d", l);
            free(net->layers); free(net); return NULL;
        }
        off += L->bias_n + L->weight_n;
        net->act_total += L->out_n;
        net->in_total += L->in_n;
    }
    net->n_params = off;
    net->params = (R*)calloc((size_t)(off > 0 ? off : 1), sizeof(R));
    net->mask = (const R**)calloc((size_t)n_layers, sizeof(R*));
    return net;
}

/* ALayer<D>::set_keep_rate, ALayer.inl:154-157. */
int orc_net_set_keep_rate(orc_net* net, int layer, double keep_rate) {
    if (layer < 0 || layer >= net->n_layers) return fail("orc_net_set_keep_rate: no such layer");
    net->layers[layer].keep_rate = (R)keep_rate;
    return 0;
}

/* Number of mask elements one mini-batch consumes: the input sizes of the layers with keep_rate < 1, in layer order
 * (Net.inl:229-230 calls SetUpDropoutMask on every layer, ALayer.inl:28 skips those that keep everything). */
int64_t orc_net_mask_size(const orc_net* net) {
    int64_t n = 0;
    for (int l = 0; l < net->n_layers; ++l) if (net->layers[l].keep_rate < (R)1) n += net->layers[l].in_n;
    return n;
}

/* Points the layers at their masks inside one batch's block of `masks` (NULL: no dropout). */
static void bind_masks(orc_net* net, const R* masks) {
    for (int l = 0; l < net->n_layers; ++l) {
        net->mask[l] = NULL;
        if (masks && net->layers[l].keep_rate < (R)1) { net->mask[l] = masks; masks += net->layers[l].in_n; }
    }
}

void orc_net_destroy(orc_net* net) {
    if (net) free((void*)net->mask);
    if (!net) return;
    free(net->params); free(net->layers); free(net);
}

int64_t orc_net_param_count(const orc_net* net) { return net->n_params; }
int64_t orc_net_layer_in_size(const orc_net* net, int l) { return net->layers[l].in_n; }
int64_t orc_net_layer_out_size(const orc_net* net, int l) { return net->layers[l].out_n; }
int64_t orc_net_layer_param_offset(const orc_net* net, int l) { return net->layers[l].param_off; }

int orc_net_set_params(orc_net* net, const orc_real* p) { memcpy(net->params, p, (size_t)net->n_params * sizeof(R)); return 0; }
int orc_net_get_params(const orc_net* net, orc_real* p) { memcpy(p, net->params, (size_t)net->n_params * sizeof(R)); return 0; }

/* ---------------------------------------------------------------------------------------------------------------
 * Activations.  Value + auxiliary ("Derivatives") in place.
 *   sigmoid  ActivationFunctionHelper.cpp:84-95    tanh  :103-114    relu  :64-76
 *   linear   generic dual-number path :33-42 with the factory lambda ActivationFunctionFactory.h:38 -> (x, 1)
 *   softmax  ActivationFunction.inl:47-54 + ActivationFunctionHelper.cpp:44-48: aux = exp(x - max) (un-normalised)
 * ------------------------------------------------------------------------------------------------------------- */
static void act_and_aux(int id, R* v, R* d, int64_t n) {
    switch (id) {
    case DLB_ACT_SIGMOID:
        for (int64_t i = 0; i < n; ++i) { const R val = 1 / (1 + R_EXP(-v[i])); v[i] = val; if (d) d[i] = val * (1 - val); }
        break;
    case DLB_ACT_TANH:
        for (int64_t i = 0; i < n; ++i) { const R val = R_TANH(v[i]); v[i] = val; if (d) d[i] = 1 - val * val; }
        break;
    case DLB_ACT_RELU:
        for (int64_t i = 0; i < n; ++i) { const R x = v[i]; v[i] = x > 0 ? x : 0; if (d) d[i] = (R)(x > 0 ? 1 : 0); }
        break;
    case DLB_ACT_LINEAR:
        if (d) for (int64_t i = 0; i < n; ++i) d[i] = (R)1;
        break;
    case DLB_ACT_SOFTMAX: {
        R m = v[0];
        for (int64_t i = 1; i < n; ++i) if (m < v[i]) m = v[i]; /* std::max_element: first maximum */
        R sum = (R)0;
        for (int64_t i = 0; i < n; ++i) { v[i] = R_EXP(v[i] - m); }
        for (int64_t i = 0; i < n; ++i) sum += v[i];              /* BasicCollection::sum, BasicCollection.cpp:181-188 */
        const R factor = (R)1 / sum;
        for (int64_t i = 0; i < n; ++i) { if (d) d[i] = v[i]; v[i] = v[i] * factor; }
        break;
    }
    default: break;
    }
}

/* Gradient with respect to the activation's input.
 *   generic   AFunction::calc_in_grad, ActivationFunction.inl:102-106: out = aux (.) out_grad
 *   softmax   ActivationFunctionHelper.cpp:50-58 (evaluate_softmax_input_grad) */
static void act_in_grad(int id, const R* out_grad, const R* aux, R* res, int64_t n) {
    if (id == DLB_ACT_SOFTMAX) {
        R s = (R)0;
        for (int64_t i = 0; i < n; ++i) s += aux[i];
        const R ood = (R)1 / s;
        for (int64_t i = 0; i < n; ++i) res[i] = aux[i] * out_grad[i] * ood;
        R t = (R)0;
        for (int64_t i = 0; i < n; ++i) t += res[i];
        const R temp_sum = t * ood;
        for (int64_t i = 0; i < n; ++i) res[i] = res[i] - aux[i] * temp_sum;
    } else {
        for (int64_t i = 0; i < n; ++i) res[i] = aux[i] * out_grad[i];
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * Cost functions and their derivative.  The reference differentiates the lambdas of CostFunctionFactory.h:29-55
 * with forward-mode dual numbers (CostFunctionHelper.cpp:52-57, Dual.h); the dual arithmetic is expanded here:
 *   SQUARED_ERROR  0.5*diff*diff                 -> d = 0.5*diff + 0.5*diff  (Dual.h operator*=, :134-142)
 *   CROSS_ENTROPY  3 branches wrapped in Func::nan_to_num (Functions.h:65-78): when the value OR the derivative
 *                  is inf/nan the result is dual(+-max) or dual(0), whose dual part is 0.
 *   LINEAR         x                             -> d = 1
 * ------------------------------------------------------------------------------------------------------------- */
static int bad(R x) { return isinf(x) || isnan(x); }

static R cost_deriv(int id, R x, R ref) {
    switch (id) {
    case DLB_COST_SQUARED_ERROR: { const R diff = x - ref; const R h = (R)0.5 * diff; return (R)1 * h + (R)0.5 * diff; }
    case DLB_COST_CROSS_ENTROPY: {
        if (ref <= (R)0) {            /* -log(1 - x) */
            const R a = (R)1 - x;     /* dual (1 - x, -1) */
            const R da = (R)-1 * ((R)1 / a);
            const R val = -R_LOG(a), d = da * (R)-1;
            return (bad(val) || bad(d)) ? (R)0 : d;
        }
        if (ref >= (R)1) {            /* -log(x) */
            const R dl = (R)1 * ((R)1 / x);
            const R val = -R_LOG(x), d = dl * (R)-1;
            return (bad(val) || bad(d)) ? (R)0 : d;
        }
        /* -(ref*log(x) + (1-ref)*log(1-x)) */
        const R lx = R_LOG(x), dlx = (R)1 * ((R)1 / x);
        const R a = (R)1 - x, la = R_LOG(a), dla = (R)-1 * ((R)1 / a);
        const R omr = (R)1 - ref;
        const R v1 = ref * lx, d1 = dlx * ref + (R)0 * lx;       /* dual(ref,0) * log(x) */
        const R v2 = omr * la, d2 = dla * omr + (R)0 * la;       /* dual(1-ref,0) * log(1-x) */
        const R val = -(v1 + v2), d = (d1 + d2) * (R)-1;
        return (bad(val) || bad(d)) ? (R)0 : d;
    }
    case DLB_COST_LINEAR: return (R)1;
    default: return (R)NAN;
    }
}

static R nan_to_num(R v) { /* Functions.h:65-78 */
    if (isinf(v)) return v < (R)0 ? -R_MAX : R_MAX;
    if (isnan(v)) return (R)0;
    return v;
}

static R cost_value(int id, R x, R ref) { /* CostFunctionFactory.h:33-53 */
    switch (id) {
    case DLB_COST_SQUARED_ERROR: { const R diff = x - ref; return (R)0.5 * diff * diff; }
    case DLB_COST_CROSS_ENTROPY:
        if (ref <= (R)0) return nan_to_num(-R_LOG((R)1 - x));
        if (ref >= (R)1) return nan_to_num(-R_LOG(x));
        return nan_to_num(-(ref * R_LOG(x) + ((R)1 - ref) * R_LOG((R)1 - x)));
    case DLB_COST_LINEAR: return x;
    default: return (R)NAN;
    }
}

double orc_cost(const orc_real* o, const orc_real* r, int64_t n, int cost_id) { /* CostFunctionHelper.cpp:27-35 */
    R s = (R)0;
    for (int64_t i = 0; i < n; ++i) s += cost_value(cost_id, o[i], r[i]);
    return (double)s;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Layer forward ("act").
 * ------------------------------------------------------------------------------------------------------------- */

/* CLayer.inl:108-137 -> Tensor::convolve (Tensor.cpp:371-408): im2col row (c, ky, kx) . filter, zero padding
 * (build_receptive_field_matrix, Tensor.cpp:40-72), then `output += _biases` (untied), then activation. */
static void conv_act(const orc_layer* L, const R* params, const R* in, R* out, R* deriv) {
    const dlb_layer_spec* s = &L->s;
    const int C = s->in_c, H = s->in_h, W = s->in_w, KH = s->win_h, KW = s->win_w;
    const int64_t ksz = (int64_t)C * KH * KW, P = (int64_t)L->out_h * L->out_w;
    const R* bias = params + L->param_off;
    const R* filt = bias + L->bias_n;
    for (int f = 0; f < L->out_c; ++f) {
        const R* K = filt + f * ksz;
        for (int oy = 0; oy < L->out_h; ++oy) for (int ox = 0; ox < L->out_w; ++ox) {
            const int ty0 = oy * s->stride_h - s->pad_h, tx0 = ox * s->stride_w - s->pad_w;
            R acc = (R)0; /* std::inner_product(row, row + k_size, kernel_flat, Real(0)) */
            for (int c = 0; c < C; ++c) for (int ky = 0; ky < KH; ++ky) for (int kx = 0; kx < KW; ++kx) {
                const int ty = ty0 + ky, tx = tx0 + kx;
                const R x = (ty >= 0 && ty < H && tx >= 0 && tx < W) ? in[((int64_t)c * H + ty) * W + tx] : (R)0;
                acc = acc + x * K[((int64_t)c * KH + ky) * KW + kx];
            }
            const int64_t o = f * P + (int64_t)oy * L->out_w + ox;
            out[o] = acc + bias[o];
        }
    }
    act_and_aux(s->activation, out, deriv, L->out_n);
}

/* PLayer.inl:115-138 -> Tensor::min_max_pool<true> (Tensor.cpp:673-716) / scale_pool (Tensor.cpp:755-779). */
static void pool_act(const orc_layer* L, const R* in, R* out, int32_t* map) {
    const dlb_layer_spec* s = &L->s;
    const int H = s->in_h, W = s->in_w;
    for (int c = 0; c < L->out_c; ++c) for (int oy = 0; oy < L->out_h; ++oy) for (int ox = 0; ox < L->out_w; ++ox) {
        const int64_t o = ((int64_t)c * L->out_h + oy) * L->out_w + ox;
        if (s->pool_type == DLB_POOL_AVERAGE) {
            R acc = (R)0;
            for (int ky = 0; ky < s->win_h; ++ky) for (int kx = 0; kx < s->win_w; ++kx)
                acc += in[((int64_t)c * H + oy * s->win_h + ky) * W + ox * s->win_w + kx];
            out[o] = acc * ((R)1 / (R)((int64_t)s->win_h * s->win_w));
        } else {
            const int is_max = s->pool_type == DLB_POOL_MAX;
            R best = is_max ? -R_MAX : R_MAX;
            int32_t best_id = -1;
            for (int ky = 0; ky < s->win_h; ++ky) for (int kx = 0; kx < s->win_w; ++kx) {
                const int32_t id = (int32_t)(((int64_t)c * H + oy * s->win_h + ky) * W + ox * s->win_w + kx);
                const R cur = in[id];
                if (is_max ? (best < cur) : (best > cur)) { best = cur; best_id = id; } /* strict: first wins ties */
            }
            out[o] = best;
            if (map) map[o] = best_id;
        }
    }
}

/* NLayer.inl:131-142 -> Matrix::mul_add (Matrix.cpp:180-198): inner_product(row, x, 0) + b. */
static void full_act(const orc_layer* L, const R* params, const R* in, R* out, R* deriv) {
    const R* bias = params + L->param_off;
    const R* Wt = bias + L->bias_n;
    for (int r = 0; r < L->s.out_c; ++r) {
        const R* row = Wt + (int64_t)r * L->in_n;
        R acc = (R)0;
        for (int64_t k = 0; k < L->in_n; ++k) acc = acc + row[k] * in[k];
        out[r] = acc + bias[r];
    }
    act_and_aux(L->s.activation, out, deriv, L->out_n);
}

static void layer_act(const orc_net* net, int l, const R* in, R* out, R* deriv, int32_t* map) {
    const orc_layer* L = &net->layers[l];
    switch (L->s.type) {
    case DLB_LAYER_CONVOLUTION: conv_act(L, net->params, in, out, deriv); break;
    case DLB_LAYER_PULL: pool_act(L, in, out, map); break;
    default: full_act(L, net->params, in, out, deriv); break;
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * Layer backward ("backpropagate").  `grad` points at this layer's slice of the flat gradient; `scale` is the
 * reference's gradient_scale_factor (0 => overwrite, else grad = scale*grad + new).
 * ------------------------------------------------------------------------------------------------------------- */

/* NLayer.inl:144-171; Matrix.cpp:317-354 (outer products), Matrix.cpp:200-219 (transpose_mul). */
static void full_backprop(const orc_layer* L, const R* params, const R* deltas, const R* in, const R* deriv,
                          R* in_grad, R* grad, R scale, R* g) {
    const int rows = L->s.out_c;
    const int64_t cols = L->in_n;
    R* db = grad;
    R* dW = grad + L->bias_n;
    act_in_grad(L->s.activation, deltas, deriv, g, rows);
    if (scale != (R)0) {
        for (int r = 0; r < rows; ++r) db[r] = db[r] * scale + g[r];
        for (int r = 0; r < rows; ++r) { const R f = g[r]; R* row = dW + r * cols;
            for (int64_t c = 0; c < cols; ++c) row[c] = f * in[c] + scale * row[c]; }
    } else {
        for (int r = 0; r < rows; ++r) db[r] = g[r];
        for (int r = 0; r < rows; ++r) { const R f = g[r]; R* row = dW + r * cols;
            for (int64_t c = 0; c < cols; ++c) row[c] = f * in[c]; }
    }
    if (!in_grad) return;
    const R* Wt = params + L->param_off + L->bias_n;
    for (int64_t c = 0; c < cols; ++c) in_grad[c] = (R)0;
    for (int r = 0; r < rows; ++r) { const R sc = g[r]; const R* row = Wt + r * cols;
        for (int64_t c = 0; c < cols; ++c) in_grad[c] = row[c] * sc + in_grad[c]; }
}

/* CLayer.inl:139-197; Tensor::convolution_kernel_gradient (Tensor.cpp:466-498) and
 * Tensor::convolution_input_gradient (Tensor.cpp:500-546); zero factors are skipped as in the reference. */
static void conv_backprop(const orc_layer* L, const R* params, const R* deltas, const R* in, const R* deriv,
                          R* in_grad, R* grad, R scale, R* g) {
    const dlb_layer_spec* s = &L->s;
    const int C = s->in_c, H = s->in_h, W = s->in_w, KH = s->win_h, KW = s->win_w;
    const int64_t ksz = (int64_t)C * KH * KW, P = (int64_t)L->out_h * L->out_w;
    R* db = grad;
    R* dK_all = grad + L->bias_n;
    const R* filt = params + L->param_off + L->bias_n;
    act_in_grad(s->activation, deltas, deriv, g, L->out_n);
    if (scale != (R)0) for (int64_t i = 0; i < L->out_n; ++i) db[i] = db[i] * scale + g[i];
    else for (int64_t i = 0; i < L->out_n; ++i) db[i] = g[i];
    if (in_grad) for (int64_t i = 0; i < L->in_n; ++i) in_grad[i] = (R)0;
    for (int f = 0; f < L->out_c; ++f) {
        R* dK = dK_all + f * ksz;
        const R* K = filt + f * ksz;
        if (scale != (R)0) for (int64_t k = 0; k < ksz; ++k) dK[k] *= scale;
        else for (int64_t k = 0; k < ksz; ++k) dK[k] = (R)0;
        for (int64_t r = 0; r < P; ++r) {
            const R fac = g[f * P + r];
            if (fac == (R)0) continue;
            const int oy = (int)(r / L->out_w), ox = (int)(r % L->out_w);
            const int ty0 = oy * s->stride_h - s->pad_h, tx0 = ox * s->stride_w - s->pad_w;
            for (int c = 0; c < C; ++c) for (int ky = 0; ky < KH; ++ky) for (int kx = 0; kx < KW; ++kx) {
                const int ty = ty0 + ky, tx = tx0 + kx;
                const R x = (ty >= 0 && ty < H && tx >= 0 && tx < W) ? in[((int64_t)c * H + ty) * W + tx] : (R)0;
                const int64_t k = ((int64_t)c * KH + ky) * KW + kx;
                dK[k] = dK[k] + x * fac;
            }
        }
        if (!in_grad) continue;
        for (int64_t r = 0; r < P; ++r) {
            const R fac = g[f * P + r];
            if (fac == (R)0) continue;
            const int oy = (int)(r / L->out_w), ox = (int)(r % L->out_w);
            const int ty0 = oy * s->stride_h - s->pad_h, tx0 = ox * s->stride_w - s->pad_w;
            for (int c = 0; c < C; ++c) for (int ky = 0; ky < KH; ++ky) for (int kx = 0; kx < KW; ++kx) {
                const int ty = ty0 + ky, tx = tx0 + kx;
                if (ty < 0 || ty >= H || tx < 0 || tx >= W) continue;
                const int64_t t = ((int64_t)c * H + ty) * W + tx;
                in_grad[t] = in_grad[t] + K[((int64_t)c * KH + ky) * KW + kx] * fac;
            }
        }
    }
}

/* PLayer.inl:140-168 -> min_max_pool_input_gradient (Tensor.cpp:721-734, assignment through the index map after a
 * zero fill) / scale_pool_input_gradient (Tensor.cpp:781-802, assignment of g*scale to the whole window). */
static void pool_backprop(const orc_layer* L, const R* deltas, const int32_t* map, R* in_grad) {
    const dlb_layer_spec* s = &L->s;
    if (!in_grad) return;
    for (int64_t i = 0; i < L->in_n; ++i) in_grad[i] = (R)0;
    if (s->pool_type == DLB_POOL_AVERAGE) {
        const R sc = (R)1 / (R)((int64_t)s->win_h * s->win_w);
        for (int c = 0; c < L->out_c; ++c) for (int oy = 0; oy < L->out_h; ++oy) for (int ox = 0; ox < L->out_w; ++ox) {
            const R v = deltas[((int64_t)c * L->out_h + oy) * L->out_w + ox] * sc;
            for (int ky = 0; ky < s->win_h; ++ky) for (int kx = 0; kx < s->win_w; ++kx)
                in_grad[((int64_t)c * s->in_h + oy * s->win_h + ky) * s->in_w + ox * s->win_w + kx] = v;
        }
    } else {
        for (int64_t i = 0; i < L->out_n; ++i) in_grad[map[i]] = deltas[i];
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * Per-sample workspace and the forward/backward sweep of Net.inl:117-132 + 263-276.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
    R** input;   /* input[l]: input of layer l; input[L]: net output */
    R** deriv;
    int32_t** map;
    R* delta_a; R* delta_b; R* g;
    int64_t max_n;
} orc_ws;

static orc_ws* ws_create(const orc_net* net) {
    orc_ws* w = (orc_ws*)calloc(1, sizeof *w);
    const int L = net->n_layers;
    w->input = (R**)calloc((size_t)L + 1, sizeof(R*));
    w->deriv = (R**)calloc((size_t)L, sizeof(R*));
    w->map = (int32_t**)calloc((size_t)L, sizeof(int32_t*));
    int64_t mx = 1;
    for (int l = 0; l < L; ++l) {
        const orc_layer* Y = &net->layers[l];
        w->input[l] = (R*)calloc((size_t)Y->in_n, sizeof(R));
        w->deriv[l] = (R*)calloc((size_t)Y->out_n, sizeof(R));
        w->map[l] = (int32_t*)calloc((size_t)Y->out_n, sizeof(int32_t));
        if (Y->in_n > mx) mx = Y->in_n;
        if (Y->out_n > mx) mx = Y->out_n;
    }
    w->input[L] = (R*)calloc((size_t)net->layers[L - 1].out_n, sizeof(R));
    w->delta_a = (R*)calloc((size_t)mx, sizeof(R));
    w->delta_b = (R*)calloc((size_t)mx, sizeof(R));
    w->g = (R*)calloc((size_t)mx, sizeof(R));
    w->max_n = mx;
    return w;
}

static void ws_destroy(const orc_net* net, orc_ws* w) {
    for (int l = 0; l < net->n_layers; ++l) { free(w->input[l]); free(w->deriv[l]); free(w->map[l]); }
    free(w->input[net->n_layers]);
    free(w->input); free(w->deriv); free(w->map); free(w->delta_a); free(w->delta_b); free(w->g); free(w);
}

static void forward_traced(const orc_net* net, orc_ws* w, const R* input) {
    memcpy(w->input[0], input, (size_t)net->layers[0].in_n * sizeof(R)); /* gradient_cache[0].Input = input, Net.inl:123 */
    for (int l = 0; l < net->n_layers; ++l) {
        if (net->mask[l])   /* ApplyDropout(gradient_cache[l].Input, true): hadamard_prod_in_place with the keep mask, Net.inl:126, ALayer.inl:46-55 */
            for (int64_t i = 0; i < net->layers[l].in_n; ++i) w->input[l][i] = w->input[l][i] * net->mask[l][i];
        layer_act(net, l, w->input[l], w->input[l + 1], w->deriv[l], w->map[l]);
    }
}

/* deltas enter in w->delta_a (cost derivative); grads is the flat gradient buffer of one "thread". */
static void backward(const orc_net* net, orc_ws* w, R* grads, R scale, R* in_grads_out) {
    R* cur = w->delta_a; R* nxt = w->delta_b;
    int64_t in_off = net->in_total;
    for (int l = net->n_layers - 1; l >= 0; --l) {
        const orc_layer* Y = &net->layers[l];
        in_off -= Y->in_n;
        R* ig = l != 0 ? nxt : NULL; /* evaluate_input_gradient = layer_id != 0, Net.inl:272 */
        R* grad = grads + Y->param_off;
        switch (Y->s.type) {
        case DLB_LAYER_CONVOLUTION: conv_backprop(Y, net->params, cur, w->input[l], w->deriv[l], ig, grad, scale, w->g); break;
        case DLB_LAYER_PULL: pool_backprop(Y, cur, w->map[l], ig); break;
        default: full_backprop(Y, net->params, cur, w->input[l], w->deriv[l], ig, grad, scale, w->g); break;
        }
        if (ig && net->mask[l])   /* ApplyDropout(e_data.out(), true) on the input gradient, Net.inl:274-275 */
            for (int64_t i = 0; i < Y->in_n; ++i) ig[i] = ig[i] * net->mask[l][i];
        if (ig && in_grads_out) memcpy(in_grads_out + in_off, ig, (size_t)Y->in_n * sizeof(R));
        R* t = cur; cur = nxt; nxt = t;
    }
}

/* Masks for orc_net_trace (one block as in orc_net_learn_dropout); NULL clears them.  The caller keeps the array alive. */
int orc_net_bind_masks(orc_net* net, const orc_real* masks) { bind_masks(net, masks); return 0; }

int orc_net_trace(const orc_net* net, const orc_real* input, const orc_real* target, int cost_id, double scale,
                  orc_real* acts, orc_real* derivs, int32_t* index_map, orc_real* grads, orc_real* in_grads,
                  orc_real* cost_grad) {
    orc_ws* w = ws_create(net);
    forward_traced(net, w, input);
    int64_t off = 0;
    for (int l = 0; l < net->n_layers; ++l) {
        const orc_layer* Y = &net->layers[l];
        if (acts) memcpy(acts + off, w->input[l + 1], (size_t)Y->out_n * sizeof(R));
        if (derivs && Y->s.type != DLB_LAYER_PULL) memcpy(derivs + off, w->deriv[l], (size_t)Y->out_n * sizeof(R));
        if (index_map && Y->s.type == DLB_LAYER_PULL && Y->s.pool_type != DLB_POOL_AVERAGE)
            memcpy(index_map + off, w->map[l], (size_t)Y->out_n * sizeof(int32_t));
        off += Y->out_n;
    }
    const orc_layer* last = &net->layers[net->n_layers - 1];
    for (int64_t i = 0; i < last->out_n; ++i) w->delta_a[i] = cost_deriv(cost_id, w->input[net->n_layers][i], target[i]);
    if (cost_grad) memcpy(cost_grad, w->delta_a, (size_t)last->out_n * sizeof(R));
    R* gbuf = grads;
    R* tmp = NULL;
    if (!gbuf) { tmp = (R*)calloc((size_t)net->n_params + 1, sizeof(R)); gbuf = tmp; }
    backward(net, w, gbuf, (R)scale, in_grads);
    free(tmp);
    ws_destroy(net, w);
    return 0;
}

/* ALayer<D>::update, ALayer.inl:97-122; BasicCollection.cpp:110-135. */
static void update_params(orc_net* net, const R* grad, R l_rate, R reg_factor, R grad_scale_factor) {
    const R param_scale = (R)1 + reg_factor * l_rate;
    const int nontrivial_reg = reg_factor > (R)0;
    const R gts = l_rate * grad_scale_factor;
    for (int l = 0; l < net->n_layers; ++l) {
        const orc_layer* Y = &net->layers[l];
        R* p = net->params + Y->param_off;
        const R* g = grad + Y->param_off;
        for (int64_t i = 0; i < Y->bias_n; ++i) p[i] = p[i] + g[i] * gts;                 /* biases: add_scaled */
        p += Y->bias_n; g += Y->bias_n;
        if (nontrivial_reg) for (int64_t i = 0; i < Y->weight_n; ++i) p[i] = p[i] * param_scale + g[i] * gts;
        else for (int64_t i = 0; i < Y->weight_n; ++i) p[i] = p[i] + g[i] * gts;
    }
}

int orc_net_learn(orc_net* net, const orc_real* inputs, const orc_real* refs, int64_t n, int64_t batch, int64_t epochs,
                  double lr, int cost_id, double lambda, int threads, const int64_t* perms) {
    return orc_net_learn_dropout(net, inputs, refs, n, batch, epochs, lr, cost_id, lambda, threads, perms, NULL);
}

int orc_net_learn_dropout(orc_net* net, const orc_real* inputs, const orc_real* refs, int64_t n, int64_t batch, int64_t epochs,
                          double lr, int cost_id, double lambda, int threads, const int64_t* perms, const orc_real* masks) {
    if (n == 0 || batch == 0) return 0;                                  /* Net.inl:200-201 */
    const int64_t mask_n = orc_net_mask_size(net);
    if (mask_n > 0 && !masks) return fail("orc_net_learn: a layer drops inputs but no masks were supplied");
    int64_t batch_no = 0;
    if (!perms) return fail("orc_net_learn: permutations are required");
    const R reg_factor = (R)lambda / (R)n;                               /* Net.inl:203 */
    int64_t T = threads <= 0 ? 1 : threads;
    if (T > batch) T = batch;                                            /* Net.inl:210 */
    const int64_t P = net->n_params;
    const int64_t in_n = net->layers[0].in_n, out_n = net->layers[net->n_layers - 1].out_n;
    R** G = (R**)calloc((size_t)T, sizeof(R*));
    for (int64_t j = 0; j < T; ++j) G[j] = (R*)calloc((size_t)P + 1, sizeof(R));
    orc_ws* w = ws_create(net);
    for (int64_t e = 0; e < epochs; ++e) {
        const int64_t* perm = perms + e * n;
        int64_t start = 0;
        while (start < n) {
            bind_masks(net, mask_n > 0 ? masks + batch_no * mask_n : NULL);                         /* new masks per batch, Net.inl:228-230 */
            ++batch_no;
            const int64_t end = (double)(n - start) < 1.5 * (double)batch ? n : start + batch;     /* Net.inl:235-236 */
            const int64_t ipt = (end - start + T - 1) / T;                                         /* Net.inl:238 */
            const int64_t expected = (end - start + ipt - 1) / ipt;                                /* Net.inl:242 */
            for (int64_t j = 0; j < expected; ++j) {
                const int64_t cs = start + j * ipt, ce = cs + ipt < end ? cs + ipt : end;
                for (int64_t el = cs; el < ce; ++el) {
                    const R accum = el == cs ? (R)0 : (R)1;                                          /* Net.inl:259 */
                    const int64_t item = perm[el];
                    forward_traced(net, w, inputs + item * in_n);
                    for (int64_t i = 0; i < out_n; ++i)
                        w->delta_a[i] = cost_deriv(cost_id, w->input[net->n_layers][i], refs[item * out_n + i]);
                    backward(net, w, G[j], accum, NULL);
                }
            }
            /* two-level tree sum, Net.inl:279-300 */
            const int64_t half = expected / 2;
            for (int64_t i = 1; i < half; ++i) for (int64_t k = 0; k < P; ++k) G[0][k] = G[0][k] + G[i][k];
            for (int64_t i = half + 1; i < expected; ++i) for (int64_t k = 0; k < P; ++k) G[half][k] = G[half][k] + G[i][k];
            if (half != 0) for (int64_t k = 0; k < P; ++k) G[0][k] = G[0][k] + G[half][k];
            const R norm = (R)1 / (R)(end - start);                                                /* Net.inl:310 */
            start = end;
            update_params(net, G[0], -(R)lr, reg_factor, norm);                                    /* Net.inl:314-316 */
        }
    }
    bind_masks(net, NULL);                                                                         /* DisposeDropoutMask, Net.inl:322-323 */
    ws_destroy(net, w);
    for (int64_t j = 0; j < T; ++j) free(G[j]);
    free(G);
    return 0;
}

int orc_net_eval(const orc_net* net, const orc_real* inputs, int64_t n, int32_t* classes, orc_real* outputs) {
    orc_ws* w = ws_create(net);
    const int64_t in_n = net->layers[0].in_n, out_n = net->layers[net->n_layers - 1].out_n;
    for (int64_t i = 0; i < n; ++i) {
        memcpy(w->input[0], inputs + i * in_n, (size_t)in_n * sizeof(R));
        for (int l = 0; l < net->n_layers; ++l) {
            const R keep = net->layers[l].keep_rate;
            if (keep < (R)1)   /* ApplyDropout(eval_data.in(), false): input.mul(_keep_rate), Net.inl:143, ALayer.inl:53-54 */
                for (int64_t k = 0; k < net->layers[l].in_n; ++k) w->input[l][k] = w->input[l][k] * keep;
            layer_act(net, l, w->input[l], w->input[l + 1], NULL, NULL);
        }
        const R* o = w->input[net->n_layers];
        if (outputs) memcpy(outputs + i * out_n, o, (size_t)out_n * sizeof(R));
        if (classes) { int64_t b = 0; for (int64_t k = 1; k < out_n; ++k) if (o[b] < o[k]) b = k; classes[i] = (int32_t)b; }
    }
    ws_destroy(net, w);
    return 0;
}

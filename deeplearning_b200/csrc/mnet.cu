// mnet.cu — the recurrent stack (BASELINE.json configs[3]): MNet<D> of RMLayer<D> (Elman layers with full
// back-propagation through time and gradient-norm clipping), batched over the mini-batch on the device, Real = float | double.
//
// Reference semantics (per SAMPLE, sequentially, on the host):
//   RMLayer::act            h_t = f(W_in x_t + b + W_rec h_{t-1})                                  RMLayer.inl:165-187
//   RMLayer::backpropagate  a_t = (W_rec^T a_{t+1}) * f'(z_t) + dC/dh_t * f'(z_t);   db += a_t;  dW_in += a_t x_t^T;
//                           dW_rec += a_t h_{t-1}^T;   dC/dx_t = W_in^T a_t                          RMLayer.inl:189-233
//   RMLayer::update         clip: lr_eff = lr * threshold / |g| when |g| > threshold > 0;  b += lr_eff db;
//                           W = W (1 + reg lr_eff) + lr_eff dW                                       RMLayer.inl:235-260
//   MNet::learn             gradient SUM over the mini-batch, then update with lr = -learning_rate / n   MNet.inl:106-180
//
// Here every per-sample matrix-vector product becomes one contraction over the whole mini-batch (gemm.cuh, the library's
// functor-loader tiled kernel in the net's Real), with the element-wise steps of the recurrence in the epilogues.  All
// sequence tensors live TIME-MAJOR on the device, [T][B][dim]: step t of every sample is one contiguous [B][dim] matrix,
// and the contractions that do not depend on the recurrence run ONCE over all T * B rows:
//   forward    Z = X W_in^T + b                     1 contraction, M = T B
//              h_t = f(Z_t + h_{t-1} W_rec^T)       T - 1 contractions in sequence (the recurrence), activation in the epilogue
//   backward   a_t = (a_{t+1} W_rec + dC/dh_t) f'   T - 1 contractions in sequence, epilogue as written above
//              [dW_in | db] = A^T [X | 1]           1 contraction, K = T B (bias gradient through a ones column)
//              dW_rec = A[1:]^T H[:-1]              1 contraction, K = (T - 1) B
//              dC/dX = A W_in                       1 contraction, M = T B (only for layers above the first)
// Sums over the mini-batch and over time are therefore ordered by the contraction's tiling (fixed: run-to-run bit-identical),
// not by the reference's sample order: equal within float rounding (tests: ~1e-5 relative for float, ~1e-12 for double).
// Element-wise activations only (sigmoid, tanh, relu, linear); the reference's softmax RMLayer would need the full Jacobian.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dlb200.h"
#include "fused.cuh"

namespace dlb {
namespace {

template <class R>
struct EpiBiasStore {   // z = acc + b[j]
    R* out; size_t ld; const R* bias;
    __device__ __forceinline__ void operator()(int i, int j, R acc) const { out[(size_t)i * ld + j] = acc + bias[j]; }
};
template <class R>
struct EpiRecurrentAct {   // z = acc + Z (RMLayer.inl:181-182: rec_weights().mul_add(out[t-1], out[t], out[t])), h = f(z), d = f'(z)
    R* zh; R* deriv; size_t ld; int act;
    __device__ __forceinline__ void operator()(int i, int j, R acc) const {
        const size_t o = (size_t)i * ld + j;
        R d; const R h = apply_activation<R>(act, acc + zh[o], d);
        zh[o] = h; deriv[o] = d;
    }
};
template <class R>
struct EpiBptt {   // a_t = (a_{t+1} W_rec) f'(z_t) + dC/dh_t f'(z_t)   (calc_in_grad, then add_in_grad of the next iteration)
    R* a; const R* deriv; const R* gout; size_t ld;
    __device__ __forceinline__ void operator()(int i, int j, R acc) const {
        const size_t o = (size_t)i * ld + j;
        a[o] = acc * deriv[o] + gout[o] * deriv[o];
    }
};

template <class R>
__global__ void act_inplace_kernel(R* __restrict__ zh, R* __restrict__ deriv, size_t n, int act) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    R d; const R h = apply_activation<R>(act, zh[i], d);
    zh[i] = h; deriv[i] = d;
}
template <class R>
__global__ void mul_kernel(R* __restrict__ out, const R* __restrict__ a, const R* __restrict__ b, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * b[i];
}
// [B][T][dim] (the host's layout: one sequence after the other) <-> [T][B][dim]
template <class R>
__global__ void to_time_major_kernel(const R* __restrict__ src, R* __restrict__ dst, int B, int T, int dim) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)B * T * dim;
    if (i >= total) return;
    const int k = (int)(i % dim); const size_t r = i / dim; const int b = (int)(r % B), t = (int)(r / B);
    dst[i] = src[((size_t)b * T + t) * dim + k];
}
template <class R>
__global__ void to_batch_major_kernel(const R* __restrict__ src, R* __restrict__ dst, int B, int T, int dim) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)B * T * dim;
    if (i >= total) return;
    const int k = (int)(i % dim); const size_t r = i / dim; const int t = (int)(r % T), b = (int)(r / T);
    dst[i] = src[((size_t)t * B + b) * dim + k];
}
template <class R>
__global__ void cost_grad_all_kernel(const R* __restrict__ out, const R* __restrict__ ref, R* __restrict__ grad, size_t n, int cost_id) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) grad[i] = cost_derivative<R>(cost_id, out[i], ref[i]);
}
// One block: sum of squares of a layer's gradient (MLayerGradient::norm, MLayerGradient.h:66-74), fixed order
template <class R>
__global__ void grad_norm_kernel(const R* __restrict__ g, size_t n, double* __restrict__ norm_out) {
    __shared__ double part[1024];
    double acc = 0.0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) acc += (double)g[i] * (double)g[i];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *norm_out = sqrt(part[0]);
}
// RMLayer::update: bias [0, n_bias) is not regularised, the two weight matrices are
template <class R>
__global__ void rm_update_kernel(R* __restrict__ p, const R* __restrict__ g, size_t n, size_t n_bias, const double* __restrict__ norm, R lr, R reg, R clip) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    R eff = lr;
    const R gn = (R)*norm;
    if (clip > R(0) && gn > clip) eff *= clip / gn;
    if (i < n_bias || reg <= R(0)) p[i] += g[i] * eff;
    else p[i] = p[i] * (R(1) + reg * eff) + g[i] * eff;
}

// (Measured and not kept: running the chained [B x H] x [H x H] contractions of the recurrence as ONE launch of 32 x 32 tiles over the whole
// reduction instead of gemm()'s split + finalize pair: 4.29 ms against 2.85 ms per step at BASELINE configs[3] - the 2 x 2 register tile of the
// small-tile kernel costs more than the second launch saves.)

inline void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
#define MN_CUDA(x) cuda_check((x), #x)

struct LayerRt { int in_dim, out_dim, act; double clip; int64_t off; int64_t n_bias, n_win, n_wrec; };

}  // namespace
}  // namespace dlb

struct dlb_mnet {
    virtual ~dlb_mnet() = default;
    int real_bytes = 0, depth = 0;
    std::vector<dlb::LayerRt> layers;
    int64_t n_params = 0;
    virtual void set_params(const void* h) = 0;
    virtual void get_params(void* h) = 0;
    virtual void act_host(const void* x, int64_t n, void* out) = 0;
    virtual void gradient_sum_host(const void* x, const void* y, int64_t n, int cost_id, void* grads) = 0;
    virtual void learn_host(const void* x, const void* y, int64_t n, int cost_id, double lr, double reg) = 0;
    virtual void set_batch(const void* x, const void* y, int64_t n) = 0;
    virtual void learn_resident(int cost_id, double lr, double reg, int64_t steps) = 0;
};

namespace dlb {
namespace {

thread_local std::string g_mnet_error;

template <class R>
struct MNetT final : dlb_mnet {
    dlb_context* ctx = nullptr;
    cudaStream_t st = nullptr;
    int device = 0;
    R* params = nullptr; R* grads = nullptr; double* norms = nullptr;
    int cap = 0;                                  // batch capacity of the buffers below
    R* x_bm = nullptr; R* y_bm = nullptr;         // staged host batch, [B][T][dim]
    R* x_tm = nullptr; R* y_tm = nullptr;         // time-major input / reference
    std::vector<R*> H, D, A, G;                   // per layer [T][B][out]: outputs, derivatives, dC/dz; G[l] = dC/d(output of layer l)
    R* out_bm = nullptr;
    GemmWorkspace ws{};
    int resident_B = 0;

    ~MNetT() override {
        cudaSetDevice(device);
        free_batch();
        cudaFree(params); cudaFree(grads); cudaFree(norms); cudaFree(ws.ptr);
    }
    void free_batch() {
        drop_graph();                                 // (the captured step holds these buffers' addresses)
        for (R* p : {x_bm, y_bm, x_tm, y_tm, out_bm}) cudaFree(p);
        for (auto* v : {&H, &D, &A, &G}) { for (R* p : *v) cudaFree(p); v->clear(); }
        x_bm = y_bm = x_tm = y_tm = out_bm = nullptr; cap = 0;
    }
    template <class T> T* alloc(size_t n) { T* p = nullptr; MN_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T))); return p; }

    MNetT(dlb_context* c, int T, const dlb_rmlayer_spec* specs, int n_layers) {
        ctx = c; depth = T; real_bytes = (int)sizeof(R);
        st = static_cast<cudaStream_t>(dlb_context_stream(c));
        device = dlb_context_device(c);
        MN_CUDA(cudaSetDevice(device));
        int64_t off = 0;
        for (int l = 0; l < n_layers; ++l) {
            const auto& s = specs[l];
            if (s.in_dim <= 0 || s.out_dim <= 0) throw std::invalid_argument("Invalid layer dimensions");
            if (l > 0 && specs[l - 1].out_dim != s.in_dim) throw std::invalid_argument("Incompatible input size of the new layer.");
            if (s.activation == DLB_ACT_SOFTMAX) throw std::invalid_argument("RMLayer: element-wise activations only");
            LayerRt L{s.in_dim, s.out_dim, s.activation, s.clip_threshold, off, s.out_dim, (int64_t)s.out_dim * s.in_dim, (int64_t)s.out_dim * s.out_dim};
            off += L.n_bias + L.n_win + L.n_wrec;
            layers.push_back(L);
        }
        n_params = off;
        params = alloc<R>((size_t)n_params); grads = alloc<R>((size_t)n_params); norms = alloc<double>(layers.size());
        MN_CUDA(cudaMemsetAsync(params, 0, (size_t)n_params * sizeof(R), st));
        ws.bytes = 64u << 20; ws.ptr = alloc<unsigned char>(ws.bytes);
    }

    int in0() const { return layers.front().in_dim; }
    int outN() const { return layers.back().out_dim; }

    void ensure(int64_t B) {
        if (B <= 0 || B * (int64_t)depth >= (1LL << 24)) throw std::invalid_argument("Invalid batch size");
        if (B <= cap) return;
        MN_CUDA(cudaStreamSynchronize(st));
        free_batch();
        const size_t rows = (size_t)B * depth;
        x_bm = alloc<R>(rows * in0()); x_tm = alloc<R>(rows * in0());
        y_bm = alloc<R>(rows * outN()); y_tm = alloc<R>(rows * outN()); out_bm = alloc<R>(rows * outN());
        for (const auto& L : layers) {
            H.push_back(alloc<R>(rows * L.out_dim)); D.push_back(alloc<R>(rows * L.out_dim));
            A.push_back(alloc<R>(rows * L.out_dim)); G.push_back(alloc<R>(rows * L.out_dim));
        }
        cap = (int)B;
    }
    static dim3 grid_for(size_t n) { return dim3((unsigned)((n + 255) / 256)); }

    void upload(const void* x, const void* y, int B) {
        const size_t rows = (size_t)B * depth;
        MN_CUDA(cudaMemcpyAsync(x_bm, x, rows * in0() * sizeof(R), cudaMemcpyHostToDevice, st));
        DLB_LAUNCH((to_time_major_kernel<R>), grid_for(rows * in0()), dim3(256), st, (const R*)x_bm, x_tm, B, depth, in0());
        if (y) {
            MN_CUDA(cudaMemcpyAsync(y_bm, y, rows * outN() * sizeof(R), cudaMemcpyHostToDevice, st));
            DLB_LAUNCH((to_time_major_kernel<R>), grid_for(rows * outN()), dim3(256), st, (const R*)y_bm, y_tm, B, depth, outN());
        }
    }

    // MNet::act over the mini-batch (RMLayer.inl:165-187 per layer); traces (derivatives) are always kept: the backward needs them
    void forward(int B) {
        const int T = depth;
        for (size_t l = 0; l < layers.size(); ++l) {
            const auto& L = layers[l];
            TagScope tag_("rm_forward");
            const R* in = l == 0 ? x_tm : H[l - 1];
            const R* b = params + L.off; const R* Win = b + L.n_bias; const R* Wrec = Win + L.n_win;
            const int I = L.in_dim, Hd = L.out_dim;
            gemm<R>(st, T * B, Hd, I, DenseRowLoader<R>{in, (size_t)I, nullptr}, DenseRowLoader<R>{Win, (size_t)I, nullptr},
                    EpiBiasStore<R>{H[l], (size_t)Hd, b}, ws);
            const size_t step = (size_t)B * Hd;
            DLB_LAUNCH((act_inplace_kernel<R>), grid_for(step), dim3(256), st, H[l], D[l], step, L.act);
            for (int t = 1; t < T; ++t)
                gemm<R>(st, B, Hd, Hd, DenseRowLoader<R>{H[l] + (size_t)(t - 1) * step, (size_t)Hd, nullptr}, DenseRowLoader<R>{Wrec, (size_t)Hd, nullptr},
                        EpiRecurrentAct<R>{H[l] + (size_t)t * step, D[l] + (size_t)t * step, (size_t)Hd, L.act}, ws);
        }
    }

    // MNet::add_gradient summed over the mini-batch (MNet.inl:123-146, RMLayer.inl:189-233): gradient SUM into `grads`
    void backward(int B, int cost_id) {
        const int T = depth;
        const size_t nl = layers.size();
        {
            TagScope tag_("rm_cost");
            const size_t n = (size_t)T * B * outN();
            DLB_LAUNCH((cost_grad_all_kernel<R>), grid_for(n), dim3(256), st, (const R*)H[nl - 1], (const R*)y_tm, G[nl - 1], n, cost_id);
        }
        for (int l = (int)nl - 1; l >= 0; --l) {
            const auto& L = layers[(size_t)l];
            TagScope tag_("rm_backward");
            const R* in = l == 0 ? x_tm : H[(size_t)l - 1];
            R* gb = grads + L.off; R* gWin = gb + L.n_bias; R* gWrec = gWin + L.n_win;
            const R* Win = params + L.off + L.n_bias; const R* Wrec = Win + L.n_win;
            const int I = L.in_dim, Hd = L.out_dim;
            const size_t step = (size_t)B * Hd;
            R* a = A[(size_t)l]; const R* d = D[(size_t)l]; const R* go = G[(size_t)l];
            DLB_LAUNCH((mul_kernel<R>), grid_for(step), dim3(256), st, a + (size_t)(T - 1) * step, go + (size_t)(T - 1) * step, d + (size_t)(T - 1) * step, step);
            for (int t = T - 2; t >= 0; --t)   // (a_{t+1} W_rec)[b][j] = sum_k a_{t+1}[b][k] W_rec[k][j]
                gemm<R>(st, B, Hd, Hd, DenseRowLoader<R>{a + (size_t)(t + 1) * step, (size_t)Hd, nullptr}, DenseColLoader<R>{Wrec, (size_t)Hd, nullptr},
                        EpiBptt<R>{a + (size_t)t * step, d + (size_t)t * step, go + (size_t)t * step, (size_t)Hd}, ws);
            // [dW_in | db][n][k] = sum_{t,b} a[t,b][n] * [x | 1][t,b][k]
            gemm<R>(st, Hd, I + 1, T * B, DenseColLoader<R>{a, (size_t)Hd, nullptr}, DenseColOnesLoader<R>{in, (size_t)I, nullptr, nullptr},
                    EpiWeightAndBias<R>{gb, gWin, I}, ws);
            if (T > 1)   // dW_rec[n][k] = sum_{t >= 1, b} a[t,b][n] h[t-1,b][k]
                gemm<R>(st, Hd, Hd, (T - 1) * B, DenseColLoader<R>{a + step, (size_t)Hd, nullptr}, DenseColLoader<R>{H[(size_t)l], (size_t)Hd, nullptr},
                        EpiStore<R>{gWrec, (size_t)Hd}, ws);
            else MN_CUDA(cudaMemsetAsync(gWrec, 0, (size_t)L.n_wrec * sizeof(R), st));
            if (l > 0)   // dC/dx[t,b][k] = sum_n a[t,b][n] W_in[n][k]
                gemm<R>(st, T * B, I, Hd, DenseRowLoader<R>{a, (size_t)Hd, nullptr}, DenseColLoader<R>{Win, (size_t)I, nullptr},
                        EpiStore<R>{G[(size_t)l - 1], (size_t)I}, ws);
        }
    }

    void update(int B, double lr, double reg) {
        TagScope tag_("rm_update");
        const R scale = -static_cast<R>(lr) / static_cast<R>(B);   // MNet.inl:177
        for (size_t l = 0; l < layers.size(); ++l) {
            const auto& L = layers[l];
            const size_t n = (size_t)(L.n_bias + L.n_win + L.n_wrec);
            DLB_LAUNCH((grad_norm_kernel<R>), dim3(1), dim3(1024), st, (const R*)(grads + L.off), n, norms + l);
            DLB_LAUNCH((rm_update_kernel<R>), grid_for(n), dim3(256), st, params + L.off, (const R*)(grads + L.off), n, (size_t)L.n_bias,
                       (const double*)(norms + l), scale, static_cast<R>(reg), static_cast<R>(L.clip));
        }
    }

    void set_params(const void* h) override { MN_CUDA(cudaSetDevice(device)); MN_CUDA(cudaMemcpyAsync(params, h, (size_t)n_params * sizeof(R), cudaMemcpyHostToDevice, st)); MN_CUDA(cudaStreamSynchronize(st)); }
    void get_params(void* h) override { MN_CUDA(cudaSetDevice(device)); MN_CUDA(cudaMemcpyAsync(h, params, (size_t)n_params * sizeof(R), cudaMemcpyDeviceToHost, st)); MN_CUDA(cudaStreamSynchronize(st)); }
    void act_host(const void* x, int64_t n, void* out) override {
        MN_CUDA(cudaSetDevice(device));
        ensure(n); upload(x, nullptr, (int)n); forward((int)n);
        const size_t total = (size_t)n * depth * outN();
        DLB_LAUNCH((to_batch_major_kernel<R>), grid_for(total), dim3(256), st, (const R*)H.back(), out_bm, (int)n, depth, outN());
        MN_CUDA(cudaMemcpyAsync(out, out_bm, total * sizeof(R), cudaMemcpyDeviceToHost, st));
        MN_CUDA(cudaStreamSynchronize(st));
    }
    void gradient_sum_host(const void* x, const void* y, int64_t n, int cost_id, void* g) override {
        MN_CUDA(cudaSetDevice(device));
        ensure(n); upload(x, y, (int)n); forward((int)n); backward((int)n, cost_id);
        MN_CUDA(cudaMemcpyAsync(g, grads, (size_t)n_params * sizeof(R), cudaMemcpyDeviceToHost, st));
        MN_CUDA(cudaStreamSynchronize(st));
    }
    void learn_host(const void* x, const void* y, int64_t n, int cost_id, double lr, double reg) override {
        MN_CUDA(cudaSetDevice(device));
        ensure(n); upload(x, y, (int)n); forward((int)n); backward((int)n, cost_id); update((int)n, lr, reg);
        resident_B = (int)n;
    }
    void set_batch(const void* x, const void* y, int64_t n) override {
        MN_CUDA(cudaSetDevice(device));
        ensure(n); upload(x, y, (int)n); resident_B = (int)n;
    }
    // A step on the resident mini-batch is ~400 dependent launches of 2-10 us (the recurrence), with no host decision in between: the
    // second step of a (batch, cost, rate, regularisation) combination is captured into a CUDA graph and every later one replays it
    // (same kernels, same order: bit-identical to the eager step).  DLB_MNET_GRAPH=0 keeps the eager launches (A/B measurements).
    struct StepGraph { cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; int B = 0, cost = -1; double lr = 0, reg = 0; long long launches = 0; } sg;
    void drop_graph() {
        if (sg.exec) cudaGraphExecDestroy(sg.exec);
        if (sg.graph) cudaGraphDestroy(sg.graph);
        sg = StepGraph{};
    }
    void learn_resident(int cost_id, double lr, double reg, int64_t steps) override {
        MN_CUDA(cudaSetDevice(device));
        if (resident_B <= 0) throw std::invalid_argument("no resident mini-batch (dlb_mnet_set_batch)");
        auto eager = [&] { forward(resident_B); backward(resident_B, cost_id); update(resident_B, lr, reg); };
        int64_t s = 0;
        static const bool graphs_on = [] { const char* e = std::getenv("DLB_MNET_GRAPH"); return !(e && e[0] == '0'); }();
        if (graphs_on && !g_timing_on && steps >= 2) {
            if (!sg.exec || sg.B != resident_B || sg.cost != cost_id || sg.lr != lr || sg.reg != reg) {
                drop_graph();
                eager(); ++s;                                      // first step eager (one-time kernel attributes are set outside a capture)
                const long long before = g_launch_count.load();
                MN_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                try { eager(); } catch (...) { cudaGraph_t junk = nullptr; cudaStreamEndCapture(st, &junk); if (junk) cudaGraphDestroy(junk); throw; }
                MN_CUDA(cudaStreamEndCapture(st, &sg.graph));
                sg.launches = g_launch_count.load() - before;
                g_launch_count.store(before);
                MN_CUDA(cudaGraphInstantiate(&sg.exec, sg.graph, 0));
                sg.B = resident_B; sg.cost = cost_id; sg.lr = lr; sg.reg = reg;
            }
            for (; s < steps; ++s) { MN_CUDA(cudaGraphLaunch(sg.exec, st)); g_launch_count.fetch_add(sg.launches, std::memory_order_relaxed); }
        }
        for (; s < steps; ++s) eager();
    }
};

template <class F>
int mnet_guarded(F&& f) noexcept {
    try { f(); return 0; }
    catch (const std::invalid_argument& e) { g_mnet_error = e.what(); return 2; }
    catch (const std::exception& e) { g_mnet_error = e.what(); return 1; }
    catch (...) { g_mnet_error = "unknown error"; return 1; }
}

}  // namespace
}  // namespace dlb

extern "C" {

const char* dlb_mnet_last_error(void) { return dlb::g_mnet_error.c_str(); }

int dlb_mnet_create(dlb_context* ctx, int depth, const dlb_rmlayer_spec* layers, int n_layers, int real_bytes, dlb_mnet** out) {
    return dlb::mnet_guarded([&] {
        if (!ctx || !layers || !out || n_layers <= 0 || depth <= 0) throw std::invalid_argument("null or empty argument");
        if (real_bytes == 4) *out = new dlb::MNetT<float>(ctx, depth, layers, n_layers);
        else if (real_bytes == 8) *out = new dlb::MNetT<double>(ctx, depth, layers, n_layers);
        else throw std::invalid_argument("real_bytes must be 4 or 8");
    });
}
int dlb_mnet_destroy(dlb_mnet* net) { delete net; return 0; }
int64_t dlb_mnet_param_count(const dlb_mnet* net) { return net ? net->n_params : -1; }
int64_t dlb_mnet_layer_param_offset(const dlb_mnet* net, int layer) {
    return net && layer >= 0 && layer < (int)net->layers.size() ? net->layers[(size_t)layer].off : -1;
}
int dlb_mnet_set_params(dlb_mnet* net, const void* host_params) { return dlb::mnet_guarded([&] { if (!net || !host_params) throw std::invalid_argument("null argument"); net->set_params(host_params); }); }
int dlb_mnet_get_params(dlb_mnet* net, void* host_params) { return dlb::mnet_guarded([&] { if (!net || !host_params) throw std::invalid_argument("null argument"); net->get_params(host_params); }); }
int dlb_mnet_act_host(dlb_mnet* net, const void* inputs, int64_t n, void* outputs) {
    return dlb::mnet_guarded([&] { if (!net || !inputs || !outputs) throw std::invalid_argument("null argument"); net->act_host(inputs, n, outputs); });
}
int dlb_mnet_gradient_sum_host(dlb_mnet* net, const void* inputs, const void* refs, int64_t n, int cost_id, void* host_grads) {
    return dlb::mnet_guarded([&] { if (!net || !inputs || !refs || !host_grads) throw std::invalid_argument("null argument"); net->gradient_sum_host(inputs, refs, n, cost_id, host_grads); });
}
int dlb_mnet_learn_host(dlb_mnet* net, const void* inputs, const void* refs, int64_t n, int cost_id, double learning_rate, double reg_factor) {
    return dlb::mnet_guarded([&] { if (!net || !inputs || !refs) throw std::invalid_argument("null argument"); net->learn_host(inputs, refs, n, cost_id, learning_rate, reg_factor); });
}
int dlb_mnet_set_batch(dlb_mnet* net, const void* inputs, const void* refs, int64_t n) {
    return dlb::mnet_guarded([&] { if (!net || !inputs || !refs) throw std::invalid_argument("null argument"); net->set_batch(inputs, refs, n); });
}
int dlb_mnet_learn_resident(dlb_mnet* net, int cost_id, double learning_rate, double reg_factor, int64_t n_steps) {
    return dlb::mnet_guarded([&] { if (!net) throw std::invalid_argument("null argument"); net->learn_resident(cost_id, learning_rate, reg_factor, n_steps); });
}

}  // extern "C"

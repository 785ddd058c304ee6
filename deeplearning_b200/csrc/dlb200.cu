// dlb200.cu — libdlb200.so: context, the batched Net<GpuDC> and the extern "C" boundary (include/dlb200.h).
//
// Host-side structure of one training step (Net<D>::learn's per-batch body, Net.inl:226-317, batched):
//   forward   : one kernel chain per layer over the whole mini-batch, traces (Derivatives / IndexMapping) kept
//   cost      : derivative of the cost w.r.t. the net output for every sample
//   backward  : per layer  g = dC/dz, bias gradient (batch sum), weight/filter gradient (batch contraction),
//               input gradient for every layer but the first (Net.inl:272)
//   allreduce : (data parallel only) sum of the flat gradient over ranks
//   update    : one launch over the flat parameter buffer (ALayer.inl:97-122)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <numeric>
#include <random>
#include <ranges>
#include <type_traits>
#include <vector>

#ifndef DLB_CUDA_EMU
#include <dlfcn.h>
#endif

#include "../../include/dlb200.h"
#include "fused.cuh"
#include "mega.h"
#include "conv28.h"
#include "tc_gemm.cuh"

namespace dlb {

std::atomic<long long> g_launch_count{0};
bool g_timing_on = false;
thread_local const char* g_tag = "";
thread_local int g_tag_layer = -1;

#ifndef DLB_CUDA_EMU
// Per-kernel device timing for bench.py's roofline pass: every launch is bracketed by two events on the launching
// stream; dlb_timing_report() synchronises once and aggregates by "<layer>:<operator tag>:<kernel text>".
namespace {
struct TimedLaunch { cudaEvent_t e0, e1; std::string key; };
std::vector<TimedLaunch> g_timed;
std::vector<cudaEvent_t> g_event_pool;
cudaEvent_t take_event() {
    if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    cudaEvent_t e; DLB_CUDA(cudaEventCreate(&e)); return e;
}
}  // namespace
void timing_begin(cudaStream_t stream, const char* kernel_text) {
    TimedLaunch t;
    t.e0 = take_event(); t.e1 = take_event();
    std::string k(kernel_text);
    const size_t lt = k.find('<');
    if (lt != std::string::npos) k = k.substr(0, lt);
    const size_t ns = k.rfind("::");
    if (ns != std::string::npos) k = k.substr(ns + 2);
    while (!k.empty() && (k.front() == '(' || k.front() == ' ')) k.erase(k.begin());
    while (!k.empty() && (k.back() == ')' || k.back() == ' ')) k.pop_back();
    t.key = std::to_string(g_tag_layer) + ":" + g_tag + ":" + k;
    DLB_CUDA(cudaEventRecord(t.e0, stream));
    g_timed.push_back(std::move(t));
}
void timing_end(cudaStream_t stream) { DLB_CUDA(cudaEventRecord(g_timed.back().e1, stream)); }
void allow_big_smem(const void* kernel, size_t bytes) {
    static std::vector<std::pair<const void*, size_t>> done;
    for (auto& d : done) if (d.first == kernel && d.second >= bytes) return;
    DLB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    done.emplace_back(kernel, bytes);
}
#else
void allow_big_smem(const void*, size_t) {}
void timing_begin(cudaStream_t, const char*) {}
void timing_end(cudaStream_t) {}
#endif

thread_local std::string g_last_error;

struct arg_error : std::invalid_argument { using std::invalid_argument::invalid_argument; };

template <class F>
int guarded(F&& f) noexcept {
    try { f(); return 0; }
    catch (const arg_error& e) { g_last_error = e.what(); return 2; }
    catch (const cuda_error& e) { g_last_error = e.what(); return 3; }
    catch (const std::exception& e) { g_last_error = e.what(); return 1; }
    catch (...) { g_last_error = "unknown error"; return 1; }
}

#define DLB_REQUIRE(cond, msg) do { if (!(cond)) throw ::dlb::arg_error(msg); } while (0)

// ---------------------------------------------------------------------------------------------------------------
// NCCL, resolved at run time: whichever libnccl.so.2 the process already has (torch bundles one) is used, else
// the system library is opened.  Nothing is linked at build time, so a single-GPU process never needs NCCL.
// ---------------------------------------------------------------------------------------------------------------
#ifndef DLB_CUDA_EMU
struct NcclApi {
    typedef struct { char internal[128]; } unique_id;
    typedef void* comm_t;
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;

    static NcclApi& get() {
        static NcclApi api = load();
        return api;
    }
    static NcclApi load() {
        NcclApi a;
        void* h = RTLD_DEFAULT;
        if (!dlsym(h, "ncclAllReduce")) {
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!h) { a.why = std::string("libnccl.so.2 not found: ") + dlerror(); return a; }
        }
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
        a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString;
        if (!a.ok) a.why = "NCCL symbols missing";
        return a;
    }
    void check(int rc, const char* what) const {
        if (rc != 0) throw std::runtime_error(std::string(what) + ": " + (GetErrorString ? GetErrorString(rc) : "nccl error"));
    }
};
constexpr int kNcclFloat = 7, kNcclDouble = 8, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values (nccl.h)
#endif

}  // namespace dlb

using namespace dlb;

// =================================================================================================================
// Context
// =================================================================================================================
struct dlb_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    GemmWorkspace ws;
    std::vector<cudaStream_t> aux_streams;   // streams of this context's nets that carry results back to the host (dlb_context_synchronize waits for them too)

    LaunchCtx lc() const { return LaunchCtx{stream, ws}; }
    void ensure_ws(size_t bytes) {
        if (bytes <= ws.bytes) return;
        if (ws.ptr) { DLB_CUDA(cudaStreamSynchronize(stream)); DLB_CUDA(cudaFree(ws.ptr)); ws.ptr = nullptr; ws.bytes = 0; }
        DLB_CUDA(cudaMalloc(&ws.ptr, bytes));
        ws.bytes = bytes;
    }
};

namespace {

void use_device(const dlb_context* ctx) { DLB_CUDA(cudaSetDevice(ctx->device)); }

template <class T> T* dev_alloc(size_t n) {
    T* p = nullptr;
    DLB_CUDA(cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}
template <class T> void dev_free(T*& p) {
    if (p) { cudaFree(p); p = nullptr; }
}

struct LayerRt {
    dlb_layer_spec s{};
    ConvGeom cg{};
    int out_c = 0, out_h = 0, out_w = 0;
    int64_t in_n = 0, out_n = 0, param_off = 0, bias_n = 0, weight_n = 0;
    double keep_rate = 1.0;   // ALayer::_keep_rate: one minus the dropout rate of the layer's input (ALayer.h:57-59)
};

// Validates the specs the way the reference's constructors and script loader would (CLayer.inl:29-40,
// PLayer.inl:27-33, Net.inl:83-88) and lays out the flat parameter buffer.
std::vector<LayerRt> build_layers(const dlb_layer_spec* specs, int n, int64_t& n_params) {
    DLB_REQUIRE(specs && n > 0, "Invalid collection of layer dimensions.");
    std::vector<LayerRt> out((size_t)n);
    int64_t off = 0;
    for (int l = 0; l < n; ++l) {
        LayerRt& L = out[(size_t)l];
        L.s = specs[l];
        const auto& s = L.s;
        L.in_n = (int64_t)s.in_c * s.in_h * s.in_w;
        L.param_off = off;
        DLB_REQUIRE(s.in_c > 0 && s.in_h > 0 && s.in_w > 0, "Invalid input size of a layer");
        switch (s.type) {
        case DLB_LAYER_FULL:
            DLB_REQUIRE(s.out_c > 0, "Can't parse output dimensions of NLayer");
            DLB_REQUIRE(s.activation >= DLB_ACT_SIGMOID && s.activation <= DLB_ACT_LINEAR, "Unknown activation function");
            L.out_c = 1; L.out_h = 1; L.out_w = s.out_c;
            L.bias_n = s.out_c; L.weight_n = (int64_t)s.out_c * L.in_n;
            break;
        case DLB_LAYER_CONVOLUTION:
            DLB_REQUIRE(s.out_c > 0, "Can't parse number of filters of CLayer");
            DLB_REQUIRE(s.win_h > 0 && s.win_w > 0, "Can't parse filter window size of CLayer");
            DLB_REQUIRE(s.stride_h > 0 && s.stride_w > 0 && s.pad_h >= 0 && s.pad_w >= 0, "Invalid paddings/strides of CLayer");
            DLB_REQUIRE(s.activation >= DLB_ACT_SIGMOID && s.activation <= DLB_ACT_LINEAR, "Unknown activation function");
            L.cg = make_conv_geom(s);
            L.out_c = s.out_c; L.out_h = L.cg.OH; L.out_w = L.cg.OW;
            L.bias_n = (int64_t)L.out_c * L.out_h * L.out_w;
            L.weight_n = (int64_t)s.out_c * L.cg.K;
            break;
        case DLB_LAYER_PULL:
            DLB_REQUIRE(s.win_h > 0 && s.win_w > 0, "Can't parse window size of PLayer");
            DLB_REQUIRE(s.pool_type >= DLB_POOL_MAX && s.pool_type <= DLB_POOL_AVERAGE, "Unsupported pool type");
            L.out_c = s.in_c;
            L.out_h = conv_out_size(s.in_h, s.win_h, 0, s.win_h);
            L.out_w = conv_out_size(s.in_w, s.win_w, 0, s.win_w);
            break;
        default: throw arg_error("Unknown layer type");
        }
        L.out_n = (int64_t)L.out_c * L.out_h * L.out_w;
        DLB_REQUIRE(L.out_n > 0, "A layer has an empty output");
        DLB_REQUIRE(L.in_n < (1LL << 30) && L.out_n < (1LL << 30), "Layer too large");
        if (l > 0) {
            const LayerRt& P = out[(size_t)l - 1];
            const bool ok = s.type == DLB_LAYER_FULL ? P.out_n == L.in_n
                                                     : (P.out_c == s.in_c && P.out_h == s.in_h && P.out_w == s.in_w);
            DLB_REQUIRE(ok, "Inconsistent input/output dimensions");
        }
        off += L.bias_n + L.weight_n;
    }
    n_params = off;
    return out;
}

// One keep mask over n inputs: Vector::fill_with_random_selection_map (Vector.cpp:208-233) restated with the same std::
// calls - uniform values in (-1, 1) as sort keys (BasicCollection.cpp:327-328), the floor(n * keep_rate) smallest are
// kept (ALayer.inl:31, product taken in Real) - so that with the same generator state the mask is the reference's, bit for
// bit (same libstdc++).
template <class R>
void draw_keep_mask(std::mt19937& gen, size_t n, R keep, R* m, std::vector<R>& vals, std::vector<int>& idx) {
    const size_t selected = static_cast<size_t>(static_cast<R>(n) * keep);
    if (selected >= n) { std::fill(m, m + n, R(1)); return; }                     // Vector.cpp:211-215
    vals.resize(n); idx.resize(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::uniform_real_distribution<R> dist(R(-1), R(1));
    std::generate(vals.begin(), vals.end(), [&]() { return dist(gen); });
    const R* v = vals.data();
    std::ranges::sort(idx, [v](const auto a, const auto b) { return v[a] < v[b]; });
    for (size_t i = 0; i < selected; ++i) m[idx[i]] = R(1);
    for (size_t i = selected; i < n; ++i) m[idx[i]] = R(0);
}

}  // namespace

// =================================================================================================================
// Net
// =================================================================================================================
struct dlb_net {
    dlb_context* ctx = nullptr;
    int real_bytes = 4;
    std::vector<LayerRt> layers;
    int64_t n_params = 0;
    std::mt19937 rng{std::random_device{}()};  // Net<D>::ran_gen, Net.inl:392-396
    std::mt19937 layer_rng{std::random_device{}()};  // ALayer<D>::ran_gen (ALayer.inl:131-135): draws the keep masks
    int rank = 0, world = 1;
#ifndef DLB_CUDA_EMU
    NcclApi::comm_t comm = nullptr;
#endif
    virtual ~dlb_net() {}
    virtual void set_params(const void* h) = 0;
    virtual void get_params(void* h) = 0;
    virtual void* device_params() = 0;
    virtual void* device_grads() = 0;
    virtual void set_dataset(const void* x, const void* y, int64_t n, bool on_device) = 0;
    virtual void learn(int64_t batch, int64_t epochs, double lr, int cost, double lambda, dlb_epoch_callback cb, void* user) = 0;
    virtual int64_t train_steps(int64_t batch, int64_t n_steps, double lr, int cost, double lambda) = 0;
    virtual void step_indices(const int64_t* idx, int64_t n, double lr, int cost, double reg_factor) = 0;
    virtual void step_host(const void* x, const void* y, int64_t n, double lr, int cost, double reg_factor, double* cost_out, bool sync) = 0;
    virtual void forward_backward_host(const void* x, const void* y, int64_t n, int cost, double scale) = 0;
    virtual void read(int what, int layer, void* dst, int64_t bytes) = 0;
    virtual void write_gradient(const void* src) = 0;
    virtual void eval_host(const void* x, int64_t n, int32_t* classes, void* outputs) = 0;
    virtual void evaluate_host(const void* x, const void* y, int64_t n, int cost, double l2_reg_factor, double* cost_out, double* correct_out) = 0;
    virtual std::string describe_plan() = 0;
    virtual void enable_graphs(bool on) = 0;
    virtual void enable_tensor_cores(int mode) = 0;
    virtual void enable_step_kernel(int mode) = 0;
    virtual void set_keep_rate(int layer, double keep_rate) = 0;
    virtual int64_t mask_size() const = 0;
    virtual void set_dropout_masks(const void* masks_host) = 0;
    virtual void p2p_export(void* handle64) = 0;
    virtual void p2p_connect(int rank, int world, const void* handles) = 0;
};

namespace {

template <class R>
struct NetT final : dlb_net {
    R* params = nullptr;
    R* grads = nullptr;
    int cap = 0;                    // batch capacity of the workspace below (grow-only, like the reference's buffers)
    std::vector<R*> out, deriv, in_grad;   // per layer [cap][out_n] / [cap][out_n] / [cap][in_n]; in_grad[L] = cost gradient
    std::vector<int*> imap;
    R* g = nullptr;                 // [cap][max out_n] activation-gradient scratch
    R* stage_x = nullptr;           // [cap][in_n(0)]   staged host batch
    R* stage_y = nullptr;           // [cap][out_n(L-1)]
    int* gather_dev = nullptr;      // [cap]
    R* data_x = nullptr; R* data_y = nullptr; int64_t n_items = 0; bool own_data = false;
    // The epoch permutation: two device buffers and two pinned host buffers used alternately, so that the next epoch's permutation is
    // uploaded behind the kernels of the current one WITHOUT draining the stream (two synchronisations per epoch used to leave a bubble of
    // tens of microseconds - at 4 steps per epoch and, data parallel, a different one on every rank, i.e. a wait at the exchange barrier).
    int* perm_dev = nullptr; int64_t perm_cap = 0;   // the buffer of the current epoch
    int* perm_bufs[2] = {nullptr, nullptr}; int* perm_host[2] = {nullptr, nullptr}; int perm_sel = 0;
#ifndef DLB_CUDA_EMU
    cudaEvent_t perm_ev[2] = {nullptr, nullptr}; bool perm_ev_pending[2] = {false, false};
#endif
    void ensure_perm(int64_t n) {
        if (perm_cap >= n) return;
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
        drop_graphs();
        for (int i = 0; i < 2; ++i) {
            dev_free(perm_bufs[i]); perm_bufs[i] = dev_alloc<int>((size_t)n);
#ifndef DLB_CUDA_EMU
            if (perm_host[i]) cudaFreeHost(perm_host[i]);
            DLB_CUDA(cudaMallocHost((void**)&perm_host[i], (size_t)n * sizeof(int)));
            if (!perm_ev[i]) DLB_CUDA(cudaEventCreateWithFlags(&perm_ev[i], cudaEventDisableTiming));
            perm_ev_pending[i] = false;
#else
            std::free(perm_host[i]); perm_host[i] = static_cast<int*>(std::malloc((size_t)n * sizeof(int)));
#endif
        }
        perm_dev = perm_bufs[0]; perm_sel = 0; perm_cap = n;
    }
    // Uploads the permutation of the next epoch into the other buffer (stream-ordered behind everything that reads the current one) and
    // resets the device-side batch cursor.
    void upload_permutation(const std::vector<std::size_t>& index_map) {
        perm_sel ^= 1;
        int* h = perm_host[perm_sel];
#ifndef DLB_CUDA_EMU
        if (perm_ev_pending[perm_sel]) DLB_CUDA(cudaEventSynchronize(perm_ev[perm_sel]));   // the upload that last read this pinned buffer (two epochs ago) has run
#endif
        for (int64_t i = 0; i < n_items; ++i) h[i] = (int)index_map[(size_t)i];
        perm_dev = perm_bufs[perm_sel];
        DLB_CUDA(cudaMemcpyAsync(perm_dev, h, (size_t)n_items * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        DLB_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(int), ctx->stream));
#ifndef DLB_CUDA_EMU
        DLB_CUDA(cudaEventRecord(perm_ev[perm_sel], ctx->stream)); perm_ev_pending[perm_sel] = true;
#endif
    }
    double* cost_dev = nullptr;
    int last_B = 0;
    UpdateSegments segs{};

    // ---- dropout (ALayer.inl:25-55): a layer with keep_rate < 1 sees its input through a 0/1 keep mask that is redrawn for
    // every mini-batch and shared by the batch's samples; its input gradient passes through the same mask; inference scales
    // the input by keep_rate instead.  Off (keep_rate = 1 everywhere) the buffers below do not exist.
    std::vector<R*> masked_in;         // [0]: [cap][in_n] masked (or scaled) copy of the batch when layer 0 drops inputs; deeper layers mask in place
    std::vector<int64_t> mask_off;     // per layer offset into mask_dev, -1: the layer keeps everything
    R* mask_dev = nullptr; int64_t mask_n = 0;
    bool masks_ready = false;
    std::vector<R> mask_host, mask_vals; std::vector<int> mask_idx;

    // ---- train mode (fused fast path, fused.cuh) ----
    struct TrainPlan {
        bool ok = false;
        bool has_cp = false;     // layers 0/1 are a fused CLayer + PLayer stage
        int first_fc = 0;        // index of the first NLayer
        ConvPoolPlan cp;
        int head_nmax = 16;
    } plan;
    // ---- the step megakernel (mega.cu): LeNet-shaped float nets run the whole step as ONE cooperative kernel ----
    struct MegaState {
        bool ok = false;            // the net has the shape the kernel covers
        bool enabled = [] { const char* e = std::getenv("DLB_MEGA"); return e && e[0] == '1'; }();   // dlb_net_enable_step_kernel; DLB_MEGA=1 turns it on for every net
        bool split = [] { const char* e = std::getenv("DLB_MEGA_SPLIT"); return e && e[0] == '1'; }();   // one launch per phase (debugging / per-phase timing)
        bool multi = [] { const char* e = std::getenv("DLB_MEGA_MULTI"); return !(e && e[0] == '0'); }();   // train_steps: consecutive steps in one launch
        mega::Shape shape{};
#ifndef DLB_CUDA_EMU
        mega::Runner* runner = nullptr;
#endif
    } mg;
    // ---- the large-batch kernels of the first LeNet stage (conv28.cu): dense FFMA2 forward / backward, 1-byte window selectors ----
    struct Conv28State {
        bool ok = false;
        int mode = [] { const char* e = std::getenv("DLB_CONV28"); return e ? std::atoi(e) : -1; }();            // 0 never, 1 always, otherwise from min_batch up
        int min_batch = [] { const char* e = std::getenv("DLB_CONV28_MIN"); return e ? std::atoi(e) : 128; }();
#ifndef DLB_CUDA_EMU
        conv28::Stage stage{};
#endif
    } c28;
    bool conv28_for(int64_t B) const { return c28.ok && c28.mode != 0 && (c28.mode == 1 || B >= c28.min_batch); }
    R* cp_partial = nullptr; int cp_grid = 0;          // [cp_grid][cp.n_par] per-CTA partial gradients of the conv stage
    int cp_rows = 0;                                    // rows of cp_partial the last backward wrote (the reduction sums these)
    bool imap1_holds_selectors = false;                 // the last fused forward left one-byte window selectors in imap[1] (conv28.cu)
    R* head_partial = nullptr; int head_chunks = 0;     // [head_chunks][n * (in_n + 1)]
    // Transposed copy W^T [in_n][out_n] of each hidden NLayer's weights (float, tensor-core steps): the input-gradient contraction
    // dx[b][k] = sum_n g[b][n] W[n][k] reads W along its strided index; from W^T both of its operands are reduction-contiguous, i.e. the
    // K-major plan of tc_gemm (TMA tiles, only the lo parts computed) instead of the transposing one.  Rebuilt at the start of every step
    // on the head's side stream, beside the forward pass (1.15 MB at LeNet: ~3 us, off the critical path).  DLB_DX_WT=0 disables it,
    // DLB_DX_WT_MIN sets the smallest mini-batch it is used for.
    std::vector<R*> wT;
    cudaEvent_t wt_fork_ev = nullptr, wt_done_ev = nullptr;
    bool wt_enabled = [] { const char* e = std::getenv("DLB_DX_WT"); return !(e && e[0] == '0'); }();
    // Measured (B200, LeNet): 78.5 -> 72.0 us for the contraction at mini-batch 8 192 (step 16.55 -> 16.94 M samples/s), but at mini-batch
    // 256 the extra kernel costs more than the contraction gains (63.6 -> 64.4 us per step): used from 1 024 samples per GPU up.
    int wt_min_batch = [] { const char* e = std::getenv("DLB_DX_WT_MIN"); return e ? std::atoi(e) : 1024; }();
    int head_chunk_cap = 0;
    int* d_cursor = nullptr;                            // device-side offset of the current batch in the epoch permutation
    static constexpr int kHeadChunk = 32;

    // CUDA graphs of the whole step, one per distinct (share, batch, hyper-parameters): the per-step launch sequence is
    // identical for every batch of an epoch because the batch position lives in d_cursor on the device.
    struct StepKey {
        int64_t mine, global; double lr, reg; int cost; const void* x; int flags; const void* gather;   // (gather: which of the two permutation buffers)
        bool operator==(const StepKey& o) const {
            return mine == o.mine && global == o.global && lr == o.lr && reg == o.reg && cost == o.cost && x == o.x && flags == o.flags && gather == o.gather;
        }
    };
    struct StepGraph {
        StepKey key{}; int seen = 0; long long launches = 0;
#ifndef DLB_CUDA_EMU
        cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
#endif
    };
    std::vector<StepGraph> step_graphs;

    // Runs `body` (a fixed launch sequence on fixed buffers) eagerly the first time a key is seen (this also sets kernel
    // attributes), captures it into a CUDA graph the second time, and replays the graph from then on.
    template <class F>
    void run_graphed(const StepKey& key, F&& body) {
#ifndef DLB_CUDA_EMU
        if (ws_stamp == ctx->ws.ptr && !g_timing_on && graphs_enabled) {
            StepGraph* sg = nullptr;
            for (auto& g : step_graphs) if (g.key == key) { sg = &g; break; }
            if (!sg) { if (step_graphs.size() >= 16) drop_graphs(); step_graphs.push_back(StepGraph{}); sg = &step_graphs.back(); sg->key = key; }
            if (sg->exec) {
                DLB_CUDA(cudaGraphLaunch(sg->exec, ctx->stream));
                g_launch_count.fetch_add(sg->launches, std::memory_order_relaxed);
                last_B = (int)key.mine;
                return;
            }
            if (sg->seen++ >= 1) {
                const long long before = g_launch_count.load();
                DLB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
                try { body(); } catch (...) { cudaGraph_t junk = nullptr; cudaStreamEndCapture(ctx->stream, &junk); if (junk) cudaGraphDestroy(junk); throw; }
                DLB_CUDA(cudaStreamEndCapture(ctx->stream, &sg->graph));
                sg->launches = g_launch_count.load() - before;
                g_launch_count.store(before);
                DLB_CUDA(cudaGraphInstantiate(&sg->exec, sg->graph, 0));
                DLB_CUDA(cudaGraphLaunch(sg->exec, ctx->stream));
                g_launch_count.fetch_add(sg->launches, std::memory_order_relaxed);
                return;
            }
        }
        ws_stamp = ctx->ws.ptr;
#endif
        body();
    }

    // Side streams: after the head kernel the step forks - the head's and every hidden NLayer's weight gradient run beside
    // the input-gradient -> conv-backward chain - and joins again before the reduce/update.  Inside a CUDA-graph capture
    // the same event fork/join becomes parallel graph branches.
    cudaStream_t side_head = nullptr, side_dw = nullptr;
    std::vector<cudaEvent_t> fork_events;
    GemmWorkspace ws_side;            // split-partial scratch of the weight-gradient contractions on side_dw
    // Host-buffer steps (dlb_net_step_host*): the next batch is copied on its own stream into the other half of a double
    // buffer while the current step computes.
    cudaStream_t side_copy = nullptr;
    // ... and the batch cost travels back on its own stream from a per-half device slot, so that the main stream carries
    // nothing but the step between two steps
    cudaStream_t side_d2h = nullptr;
    double* cost_slot[2] = {nullptr, nullptr};
    cudaEvent_t ev_cost_read[2] = {nullptr, nullptr};
    bool cost_read_pending[2] = {false, false};
    double* cur_cost_dev = nullptr;        // where the cost kernel of the step being enqueued writes
    R* hstage_x[2] = {nullptr, nullptr}; R* hstage_y[2] = {nullptr, nullptr};
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    bool hbuf_used[2] = {false, false};
    int hbuf = 0; int64_t hcap = 0;
    // Peer-memory data parallelism (dlb_net_set_distributed_p2p): [publish 0 | publish 1 | flags] in one IPC-shared allocation
    unsigned char* p2p_region = nullptr; size_t p2p_publish_bytes = 0;
    std::vector<void*> p2p_peer_regions;      // opened peer mappings (index = rank; own entry = p2p_region)
    unsigned long long p2p_step = 0;          // host mirror of *d_p2p_step (both advance once per step)
    unsigned long long* d_p2p_step = nullptr;
    bool p2p_on = false;
    // this rank's slot of the given parity inside its own region: [parity][kMaxPeers slots][publish bytes] | flags
    R* publish_buf(int parity) { return reinterpret_cast<R*>(p2p_region + ((size_t)parity * kMaxPeers + (size_t)rank) * p2p_publish_bytes); }
    unsigned long long* p2p_prof = nullptr;   // DLB_P2P_PROF=1 (development): device counters of the exchange kernel, printed when the net is destroyed
    // DLB_P2P_PUSH=1: push the slot to the peers instead of letting them pull it.  Measured equal (8 GPUs: 158.3 vs 154.0 us per step): the
    // exchange kernel's time is the wait for the slowest rank at its barrier (20-30 us at 2 GPUs, DLB_P2P_PROF=1), not the 1.2 MB per peer.
    bool p2p_push = [] { const char* e = std::getenv("DLB_P2P_PUSH"); return e && e[0] == '1'; }();
    bool overlap_enabled = true;
    // DLB_HEAD_FOLD=1: the head sums the split partials of the layer in front of it instead of a separate finalize kernel.  Measured SLOWER
    // (70.1 vs 64.1 us per step at mini-batch 256: 30 dependent partial loads per element inside a one-warp-per-sample kernel), so it is off.
    bool head_folds_finalize = [] { const char* e = std::getenv("DLB_HEAD_FOLD"); return e && e[0] == '1'; }();
    bool pdl_enabled = [] { const char* e = std::getenv("DLB_PDL"); return !(e && e[0] == '0'); }();   // programmatic dependent launches (DLB_PDL=0: off, for A/B runs)
    const void* ws_stamp = nullptr;   // the context scratch the captured graphs point into
    bool graphs_enabled = true;
    // fp32 NLayer contractions: 0 = packed-FMA path (gemm.cuh), 1 = tcgen05 3xTF32 (tc_gemm.cuh) for mini-batches that fill
    // at least one 128-row tensor-core tile (tc_min_batch samples per GPU) and the packed-FMA path below, 2 = tcgen05 always
    int tc_mode = 1;
    int tc_min_batch = 128;
    bool tc_for(int B) const { return tc_mode == 2 || (tc_mode == 1 && B >= tc_min_batch); }
    void drop_graphs() {
#ifndef DLB_CUDA_EMU
        for (auto& g : step_graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); if (g.graph) cudaGraphDestroy(g.graph); }
#endif
        step_graphs.clear();
    }

    int L() const { return (int)layers.size(); }
    int64_t in0() const { return layers.front().in_n; }
    int64_t outN() const { return layers.back().out_n; }

    NetT(dlb_context* c, std::vector<LayerRt> ls, int64_t np) {
        ctx = c; real_bytes = (int)sizeof(R); layers = std::move(ls); n_params = np;
        use_device(ctx);
        params = dev_alloc<R>((size_t)n_params);
        grads = dev_alloc<R>((size_t)n_params);
        DLB_CUDA(cudaMemsetAsync(params, 0, (size_t)n_params * sizeof(R), ctx->stream));
        DLB_CUDA(cudaMemsetAsync(grads, 0, (size_t)n_params * sizeof(R), ctx->stream));
        cost_dev = dev_alloc<double>(1);
        out.assign((size_t)L(), nullptr); deriv.assign((size_t)L(), nullptr); imap.assign((size_t)L(), nullptr);
        in_grad.assign((size_t)L() + 1, nullptr);
        segs.count = 0;
        for (const auto& Y : layers) {
            if (Y.bias_n + Y.weight_n == 0) continue;
            DLB_REQUIRE(segs.count + 2 <= kMaxSegments, "Too many parameter containers");
            segs.end[segs.count] = Y.param_off + Y.bias_n; segs.regularised[segs.count++] = 0;           // biases
            segs.end[segs.count] = Y.param_off + Y.bias_n + Y.weight_n; segs.regularised[segs.count++] = 1;  // weights / filters
        }
        if (segs.count == 0) { segs.count = 1; segs.end[0] = 0; segs.regularised[0] = 0; }
        d_cursor = dev_alloc<int>(1);
        DLB_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(int), ctx->stream));
        layout_masks();
        build_plan();
#ifndef DLB_CUDA_EMU
        if (plan.ok) {
            DLB_CUDA(cudaStreamCreateWithFlags(&side_head, cudaStreamNonBlocking));
            DLB_CUDA(cudaStreamCreateWithFlags(&side_dw, cudaStreamNonBlocking));
            fork_events.resize((size_t)L() + 4);
            for (auto& e : fork_events) DLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ws_side.bytes = 64u << 20;
            DLB_CUDA(cudaMalloc(&ws_side.ptr, ws_side.bytes));
            DLB_CUDA(cudaEventCreateWithFlags(&wt_fork_ev, cudaEventDisableTiming));
            DLB_CUDA(cudaEventCreateWithFlags(&wt_done_ev, cudaEventDisableTiming));
            wT.assign((size_t)L(), nullptr);
            if constexpr (std::is_same_v<R, float>) {
                for (int l = std::max(plan.first_fc, 1); l < L() - 1; ++l)
                    if (layers[(size_t)l].s.type == DLB_LAYER_FULL) wT[(size_t)l] = dev_alloc<R>((size_t)layers[(size_t)l].in_n * layers[(size_t)l].out_n);
            }
        }
        DLB_CUDA(cudaStreamCreateWithFlags(&side_copy, cudaStreamNonBlocking));
        DLB_CUDA(cudaStreamCreateWithFlags(&side_d2h, cudaStreamNonBlocking));
        ctx->aux_streams.push_back(side_d2h);
        for (int i = 0; i < 2; ++i) { cost_slot[i] = dev_alloc<double>(1); DLB_CUDA(cudaEventCreateWithFlags(&ev_cost_read[i], cudaEventDisableTiming)); }
        for (int i = 0; i < 2; ++i) {
            DLB_CUDA(cudaEventCreateWithFlags(&ev_copied[i], cudaEventDisableTiming));
            DLB_CUDA(cudaEventCreateWithFlags(&ev_consumed[i], cudaEventDisableTiming));
        }
#endif
    }

    // Which nets take the fused path:  [CLayer + PLayer(max|min) covered by plan_conv_pool]  NLayer*  NLayer(head)
    // with element-wise activations on all but the head, a head of at most 32 outputs whose weights fit in shared
    // memory, and at least two layers of parameters after the optional conv stage's position 0.
    void build_plan() {
        plan = TrainPlan{};
        for (const auto& Y : layers) if (Y.keep_rate < 1.0) return;   // dropout runs through the layer-by-layer operators
        int idx = 0;
        if (layers[0].s.type == DLB_LAYER_CONVOLUTION) {
            if (L() < 3) return;
            plan.cp = plan_conv_pool(layers[0].s, layers[1].s, sizeof(R));
            if (!plan.cp.ok) return;
            plan.has_cp = true;
            idx = 2;
        }
        plan.first_fc = idx;
        if (L() - idx < 2 - (plan.has_cp ? 1 : 0)) { if (!(plan.has_cp && L() - idx >= 1)) return; }
        for (int l = idx; l < L(); ++l) {
            if (layers[(size_t)l].s.type != DLB_LAYER_FULL) return;
            if (l < L() - 1 && layers[(size_t)l].s.activation == DLB_ACT_SOFTMAX) return;
        }
        if (L() - 1 == 0) return;   // a single NLayer: the head would have to read the gathered data set
        const auto& H = layers.back();
        if (H.out_n > 32 || (size_t)H.out_n * (size_t)H.in_n * sizeof(R) > 96 * 1024) return;
        plan.head_nmax = H.out_n <= 10 ? 10 : H.out_n <= 16 ? 16 : 32;
        plan.ok = true;
        plan_mega();
        c28.ok = false;
#ifndef DLB_CUDA_EMU
        if constexpr (std::is_same_v<R, float>) {
            if (plan.has_cp) {
                const auto& c = layers[0].s; const auto& p = layers[1].s;
                if (conv28::eligible(c.in_c, c.in_h, c.in_w, c.win_h, c.win_w, c.pad_h, c.pad_w, c.stride_h, c.stride_w, c.out_c, c.activation, p.win_h,
                                     p.win_w, p.pool_type)) {
                    c28.stage = conv28::Stage{c.out_c, c.activation, p.pool_type == DLB_POOL_MAX ? 1 : 0};
                    c28.ok = true;
                }
            }
        }
#endif
    }

    // Which nets the step megakernel covers (mega.h): CLayer 5x5 over one 28x28 channel, stride 1, no padding, an even number
    // of filters -> PLayer 2x2 max|min -> NLayer (<= 104 outputs, element-wise activation) -> NLayer head (<= 16 outputs).
    void plan_mega() {
        mg.ok = false;
#ifndef DLB_CUDA_EMU
        if constexpr (std::is_same_v<R, float>) {
            if (L() != 4 || !plan.has_cp) return;
            const auto& c = layers[0].s; const auto& p = layers[1].s;
            if (c.in_c != 1 || c.in_h != 28 || c.in_w != 28 || c.win_h != 5 || c.win_w != 5 || c.stride_h != 1 || c.stride_w != 1 || c.pad_h != 0 || c.pad_w != 0) return;
            if (c.out_c < 2 || (c.out_c & 1) || c.out_c > 64 || c.activation == DLB_ACT_SOFTMAX) return;
            if (p.win_h != 2 || p.win_w != 2 || (p.pool_type != DLB_POOL_MAX && p.pool_type != DLB_POOL_MIN)) return;
            if (layers[2].out_n > 104 || (layers[2].out_n & 3) || layers[2].s.activation == DLB_ACT_SOFTMAX || layers[3].out_n > 16) return;
            // 16-byte copies / vector loads of the hidden layer's parameters: its bias and weight blocks must start 16-byte aligned in the
            // flat parameter buffer (F * (576 + 25) and the hidden width multiples of 4; LeNet: 12 020 and 100)
            if ((layers[2].param_off & 3) || ((layers[2].param_off + layers[2].bias_n) & 3)) return;
            mg.shape = mega::Shape{c.out_c, (int)layers[2].out_n, (int)layers[3].out_n};
            mg.ok = true;
        }
#endif
    }

    ~NetT() override {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        free_workspace();
        dev_free(params); dev_free(grads); dev_free(cost_dev);
        for (int i = 0; i < 2; ++i) {
            dev_free(perm_bufs[i]);
#ifndef DLB_CUDA_EMU
            if (perm_host[i]) cudaFreeHost(perm_host[i]);
            if (perm_ev[i]) cudaEventDestroy(perm_ev[i]);
#else
            std::free(perm_host[i]);
#endif
        }
        drop_graphs();
#ifndef DLB_CUDA_EMU
        if (side_head) { cudaStreamSynchronize(side_head); cudaStreamDestroy(side_head); }
        if (side_dw) { cudaStreamSynchronize(side_dw); cudaStreamDestroy(side_dw); }
        for (auto& e : fork_events) cudaEventDestroy(e);
        if (ws_side.ptr) cudaFree(ws_side.ptr);
        if (wt_fork_ev) cudaEventDestroy(wt_fork_ev);
        if (wt_done_ev) cudaEventDestroy(wt_done_ev);
        if (side_copy) { cudaStreamSynchronize(side_copy); cudaStreamDestroy(side_copy); }
        if (side_d2h) {
            cudaStreamSynchronize(side_d2h);
            auto& aux = ctx->aux_streams;
            aux.erase(std::remove(aux.begin(), aux.end(), side_d2h), aux.end());
            cudaStreamDestroy(side_d2h);
        }
        for (int i = 0; i < 2; ++i) { if (ev_cost_read[i]) cudaEventDestroy(ev_cost_read[i]); dev_free(cost_slot[i]); }
        for (int i = 0; i < 2; ++i) { if (ev_copied[i]) cudaEventDestroy(ev_copied[i]); if (ev_consumed[i]) cudaEventDestroy(ev_consumed[i]); }
#endif
        for (int i = 0; i < 2; ++i) { dev_free(hstage_x[i]); dev_free(hstage_y[i]); }
#ifndef DLB_CUDA_EMU
        for (size_t r = 0; r < p2p_peer_regions.size(); ++r)
            if (p2p_peer_regions[r] && p2p_peer_regions[r] != p2p_region) cudaIpcCloseMemHandle(p2p_peer_regions[r]);
#endif
#ifndef DLB_CUDA_EMU
        if (p2p_prof) {
            unsigned long long h[4] = {0, 0, 0, 0};
            cudaMemcpy(h, p2p_prof, sizeof h, cudaMemcpyDeviceToHost);
            if (h[0]) std::fprintf(stderr, "[dlb p2p rank %d] %llu exchange kernels: block 0 passed the barrier after %.2f us, finished its data after %.2f us (averages)\n", rank,
                                   h[0], 1e-3 * (double)h[1] / (double)h[0], 1e-3 * (double)h[2] / (double)h[0]);
            cudaFree(p2p_prof);
        }
#endif
        if (p2p_region) cudaFree(p2p_region);
        dev_free(d_p2p_step);
        dev_free(cp_partial); dev_free(head_partial); dev_free(d_cursor); dev_free(mask_dev);
        free_mega();
        if (own_data) { dev_free(data_x); dev_free(data_y); }
#ifndef DLB_CUDA_EMU
        if (comm) NcclApi::get().CommDestroy(comm);
#endif
    }

    void free_mega() {
#ifndef DLB_CUDA_EMU
        if (mg.runner) { mega::runner_destroy(mg.runner); mg.runner = nullptr; }
#endif
    }

    // From mini-batch 128 up: below that the kernel's parameters drift from the oracle's several times faster than the graph path's
    // (measured at mini-batch 37: 1.5e-5 against 1.4e-6 after three steps), so small batches keep the graph path.
    static constexpr int kMegaMinBatch = 128;
    bool mega_for(int64_t B) const { return mg.ok && mg.enabled && plan.ok && B >= kMegaMinBatch && (world <= 1); }

    // `n_steps` whole training steps (forward, backward, ordered sums, update) in the megakernel.
    void mega_step(int B, const R* x, const R* refs, const int* gather_base, const int* cursor, int advance, int cost, double lr, double reg_factor,
                   int64_t batch_items, double* cost_dst, int n_steps = 1) {
#ifndef DLB_CUDA_EMU
        if constexpr (std::is_same_v<R, float>) {
            if (!mg.runner) mg.runner = mega::runner_create(mg.shape, ctx->device);
            mega::StepIn in{};
            in.stream = ctx->stream;
            in.x = x; in.refs = refs; in.gather_base = gather_base; in.cursor = cursor; in.cursor_rw = advance ? d_cursor : nullptr; in.advance = advance;
            in.B = B; in.cap = cap;
            in.act0 = layers[0].s.activation; in.is_max = layers[1].s.pool_type == DLB_POOL_MAX; in.act1 = layers[2].s.activation; in.act2 = layers[3].s.activation;
            in.cost_id = cost;
            in.params = params; in.grads = grads; in.n_params = n_params;
            in.off_b0 = layers[0].param_off; in.off_k0 = in.off_b0 + layers[0].bias_n;
            in.off_b1 = layers[2].param_off; in.off_w1 = in.off_b1 + layers[2].bias_n;
            in.off_b2 = layers[3].param_off; in.off_w2 = in.off_b2 + layers[3].bias_n;
            in.pooled = out[1]; in.dpool = in_grad[2];
            in.lr = lr; in.reg_factor = reg_factor; in.batch_items = batch_items; in.cost_out = cost_dst;
            in.n_steps = n_steps;
            in.world = 1; in.rank = 0;
            in.split_phases = mg.split ? 1 : 0;
            mega::runner_step(mg.runner, in);
            last_B = B;
        }
#else
        (void)B; (void)x; (void)refs; (void)gather_base; (void)cursor; (void)advance; (void)cost; (void)lr; (void)reg_factor; (void)batch_items; (void)cost_dst; (void)n_steps;
#endif
    }

    void free_workspace() {
        for (auto& p : out) dev_free(p);
        for (auto& p : deriv) dev_free(p);
        for (auto& p : in_grad) dev_free(p);
        for (auto& p : imap) dev_free(p);
        for (auto& p : masked_in) dev_free(p);
        dev_free(g); dev_free(stage_x); dev_free(stage_y); dev_free(gather_dev);
        cap = 0;
    }

    void ensure_capacity(int64_t B) {
        DLB_REQUIRE(B > 0 && B < (1LL << 24), "Invalid batch size");
        for (const auto& Y : layers)
            DLB_REQUIRE(B * std::max(Y.in_n, Y.out_n) < (1LL << 31), "batch * layer size exceeds the 32-bit index range");
        if (B <= cap) return;
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
        drop_graphs();
        free_workspace();
        const size_t b = (size_t)B;
        int64_t max_out = 1;
        for (int l = 0; l < L(); ++l) {
            const auto& Y = layers[(size_t)l];
            out[(size_t)l] = dev_alloc<R>(b * (size_t)Y.out_n);
            deriv[(size_t)l] = dev_alloc<R>(b * (size_t)Y.out_n);
            imap[(size_t)l] = dev_alloc<int>(b * (size_t)Y.out_n);
            if (l > 0) in_grad[(size_t)l] = dev_alloc<R>(b * (size_t)Y.in_n);
            max_out = std::max(max_out, Y.out_n);
        }
        in_grad[(size_t)L()] = dev_alloc<R>(b * (size_t)outN());
        masked_in.assign((size_t)L(), nullptr);
        if (layers[0].keep_rate < 1.0) masked_in[0] = dev_alloc<R>(b * (size_t)layers[0].in_n);
        g = dev_alloc<R>(b * (size_t)max_out);
        stage_x = dev_alloc<R>(b * (size_t)in0());
        stage_y = dev_alloc<R>(b * (size_t)outN());
        gather_dev = dev_alloc<int>(b);
        cap = (int)B;
        if (plan.ok) {
            if (plan.has_cp && !cp_partial) {
                cp_grid = kNumSMs;
                cp_partial = dev_alloc<R>((size_t)cp_grid * (size_t)plan.cp.n_par);
                // window positions the pooling never covers are never written and must read as zero
                DLB_CUDA(cudaMemsetAsync(cp_partial, 0, (size_t)cp_grid * (size_t)plan.cp.n_par * sizeof(R), ctx->stream));
            }
            const int need = (int)ceil_div<int64_t>(B, kHeadChunk);
            if (need > head_chunk_cap) {
                drop_graphs();
                dev_free(head_partial);
                const auto& H = layers.back();
                head_partial = dev_alloc<R>((size_t)need * (size_t)H.out_n * (size_t)(H.in_n + 1));
                head_chunk_cap = need;
            }
        }
        // Split-partial scratch: enough for ~2 waves of partial tiles of the largest gradient, at least 64 MiB.
        size_t need = 64u << 20;
        for (const auto& Y : layers) need = std::max(need, (size_t)(Y.weight_n + Y.bias_n) * sizeof(R) * 8);
        ctx->ensure_ws(need);
    }

    // ---- dropout plumbing ----------------------------------------------------------------------------------------
    void layout_masks() {
        mask_off.assign((size_t)L(), -1);
        mask_n = 0;
        for (int l = 0; l < L(); ++l)
            if (layers[(size_t)l].keep_rate < 1.0) { mask_off[(size_t)l] = mask_n; mask_n += layers[(size_t)l].in_n; }
    }

    void set_keep_rate(int layer, double keep_rate) override {
        DLB_REQUIRE(layer >= 0 && layer < L(), "Invalid layer index");
        DLB_REQUIRE(keep_rate > 0.0, "keep_rate must be positive");
        use_device(ctx);
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
        layers[(size_t)layer].keep_rate = keep_rate;
        drop_graphs();
        free_workspace();              // the masked-input buffers follow the keep rates
        build_plan();
        layout_masks();
        dev_free(mask_dev);
        if (mask_n > 0) mask_dev = dev_alloc<R>((size_t)mask_n);
        masks_ready = false;
    }

    int64_t mask_size() const override { return mask_n; }

    void upload_masks(const R* host) {
        DLB_CUDA(cudaMemcpyAsync(mask_dev, host, (size_t)mask_n * sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));   // the host array may be pageable and short-lived
        masks_ready = true;
    }

    void set_dropout_masks(const void* masks_host) override {
        DLB_REQUIRE(mask_n > 0, "No layer of this net drops inputs");
        DLB_REQUIRE(masks_host, "null masks");
        use_device(ctx);
        upload_masks(static_cast<const R*>(masks_host));
    }

    // SetUpDropoutMask of every layer, in layer order, once per mini-batch (Net.inl:228-230, ALayer.inl:26-35) on ALayer's generator.
    void draw_masks() {
        if (mask_n == 0) return;
        mask_host.resize((size_t)mask_n);
        for (int l = 0; l < L(); ++l)
            if (mask_off[(size_t)l] >= 0)
                draw_keep_mask<R>(layer_rng, (size_t)layers[(size_t)l].in_n, static_cast<R>(layers[(size_t)l].keep_rate),
                                  mask_host.data() + mask_off[(size_t)l], mask_vals, mask_idx);
        upload_masks(mask_host.data());
    }

    // ---- parameter / data movement ---------------------------------------------------------------------------
    void set_params(const void* h) override {
        use_device(ctx);
        DLB_CUDA(cudaMemcpyAsync(params, h, (size_t)n_params * sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    void get_params(void* h) override {
        use_device(ctx);
        DLB_CUDA(cudaMemcpyAsync(h, params, (size_t)n_params * sizeof(R), cudaMemcpyDeviceToHost, ctx->stream));
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    void* device_params() override { return params; }
    void* device_grads() override { return grads; }

    void set_dataset(const void* x, const void* y, int64_t n, bool on_device) override {
        DLB_REQUIRE(x && y && n > 0 && n < (1LL << 31), "Incompatible collection of training and reference items.");
        use_device(ctx);
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (own_data) { dev_free(data_x); dev_free(data_y); }
        if (on_device) {
            data_x = (R*)x; data_y = (R*)y; own_data = false;
        } else {
            data_x = dev_alloc<R>((size_t)n * (size_t)in0());
            data_y = dev_alloc<R>((size_t)n * (size_t)outN());
            own_data = true;
            DLB_CUDA(cudaMemcpyAsync(data_x, x, (size_t)n * (size_t)in0() * sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
            DLB_CUDA(cudaMemcpyAsync(data_y, y, (size_t)n * (size_t)outN() * sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
            DLB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        n_items = n;
    }

    // ---- forward / backward over one batch --------------------------------------------------------------------
    // x: [rows][in0] device; gather (nullable): batch slot -> row of x.
    void forward(int B, const R* x, const int* gather, bool traces) {
        imap1_holds_selectors = false;
        LaunchCtx lc = ctx->lc();
        lc.tc = std::is_same_v<R, float> && tc_for(B);
        lc.gather_rows = x == data_x ? n_items : (long long)cap;
        const R* in = x;
        const int* ga = gather;
        for (int l = 0; l < L(); ++l) {
            const auto& Y = layers[(size_t)l];
            g_tag_layer = l;
            R* y = out[(size_t)l];
            R* d = traces ? deriv[(size_t)l] : nullptr;
            const R* bias = params + Y.param_off;
            const R* w = bias + Y.bias_n;
            if (Y.keep_rate < 1.0) {   // ApplyDropout(input, trainingMode): Net.inl:126 (keep mask) / Net.inl:143 (times keep_rate)
                DLB_REQUIRE(!traces || masks_ready, "A layer drops inputs but no keep masks have been set up");
                TagScope tag_("dropout");
                const size_t total = (size_t)B * (size_t)Y.in_n;
                // like the reference, in place on the previous layer's output (gradient_cache[l].Input IS that tensor); only the
                // resident data set (layer 0) is masked into a copy
                R* dst = l == 0 ? masked_in[0] : out[(size_t)l - 1];
                DLB_LAUNCH((dropout_apply_kernel<R>), dim3((unsigned)ceil_div<size_t>(total, 256)), dim3(256), lc.stream, in, ga,
                           traces ? (const R*)(mask_dev + mask_off[(size_t)l]) : (const R*)nullptr, static_cast<R>(Y.keep_rate),
                           dst, total, (int)Y.in_n);
                in = dst; ga = nullptr;
            }
            switch (Y.s.type) {
            case DLB_LAYER_CONVOLUTION: conv_forward<R>(lc, B, Y.cg, Y.s.activation, in, ga, bias, w, y, d); break;
            case DLB_LAYER_FULL: full_forward<R>(lc, B, (int)Y.in_n, (int)Y.out_n, Y.s.activation, in, ga, bias, w, y, d); break;
            default:
                DLB_REQUIRE(ga == nullptr, "A pooling first layer cannot read a gathered data set");
                pool_forward<R>(lc, B, Y.s, in, y, traces ? imap[(size_t)l] : nullptr);
            }
            in = y; ga = nullptr;
        }
    }

    // Deltas enter in in_grad[L] (cost derivative).  scale: gradient_scale_factor on the previous gradient content.
    void backward(int B, const R* x, const int* gather, R scale) {
        LaunchCtx lc = ctx->lc();
        lc.tc = std::is_same_v<R, float> && tc_for(B);
        lc.gather_rows = x == data_x ? n_items : (long long)cap;
        const R* delta = in_grad[(size_t)L()];
        for (int l = L() - 1; l >= 0; --l) {
            const auto& Y = layers[(size_t)l];
            g_tag_layer = l;
            const R* in = l == 0 ? x : out[(size_t)l - 1];
            const int* ga = l == 0 ? gather : nullptr;
            if (Y.keep_rate < 1.0 && l == 0) { in = masked_in[0]; ga = nullptr; }   // the layer saw (and differentiates against) the masked input; l > 0: out[l-1] was masked in place
            R* dx = l == 0 ? nullptr : in_grad[(size_t)l];   // evaluate_input_gradient = layer_id != 0 (Net.inl:272)
            R* db = grads + Y.param_off;
            R* dw = db + Y.bias_n;
            const R* w = params + Y.param_off + Y.bias_n;
            switch (Y.s.type) {
            case DLB_LAYER_CONVOLUTION:
                conv_backward<R>(lc, B, Y.cg, Y.s.activation, in, ga, delta, deriv[(size_t)l], w, g, db, dw, dx, scale); break;
            case DLB_LAYER_FULL:
                full_backward<R>(lc, B, (int)Y.in_n, (int)Y.out_n, Y.s.activation, in, ga, delta, deriv[(size_t)l], w, g, db, dw, dx, scale); break;
            default:
                if (dx) pool_backward<R>(lc, B, Y.s, delta, imap[(size_t)l], dx);
            }
            if (dx && Y.keep_rate < 1.0) {   // ApplyDropout(e_data.out(), true) on the input gradient, Net.inl:274-275
                TagScope tag_("dropout");
                const size_t total = (size_t)B * (size_t)Y.in_n;
                DLB_LAUNCH((dropout_apply_kernel<R>), dim3((unsigned)ceil_div<size_t>(total, 256)), dim3(256), lc.stream, (const R*)dx,
                           (const int*)nullptr, (const R*)(mask_dev + mask_off[(size_t)l]), R(1), dx, total, (int)Y.in_n);
            }
            delta = dx;
        }
    }

    void forward_backward(int B, const R* x, const R* refs, const int* gather, int cost, R scale) {
        forward(B, x, gather, true);
        g_tag_layer = L();
        cost_gradient<R>(ctx->lc(), B, (int)outN(), cost, out[(size_t)L() - 1], refs, gather, in_grad[(size_t)L()]);
        backward(B, x, gather, scale);
        last_B = B;
    }

    // ---- train mode ---------------------------------------------------------------------------------------------
    // One fused forward + backward over `B` samples.  x / refs are the bases the (optional) gather indexes into;
    // `cursor` (device, nullable) is added to gather_base on the device, so that the launch sequence is identical
    // for every batch of an epoch (CUDA-graph friendly).
    static const int* resolve_host_gather(const int* gather_base, const int* cursor) {
        DLB_REQUIRE(cursor == nullptr, "the batch cost of a cursor-addressed batch is not available here");
        return gather_base;
    }

    // want_cost: also sum the batch cost into cost_dev - on the head's side stream, off the critical path.
    void train_forward_backward(int B, const R* x, const R* refs, const int* gather_base, const int* cursor, int cost, bool want_cost = false) {
        const LaunchCtx lc = ctx->lc();
        cudaStream_t st = ctx->stream;
        const int Ln = L();
#ifndef DLB_CUDA_EMU
        const bool fork = overlap_enabled && side_head != nullptr && !g_timing_on;
#else
        const bool fork = false;
#endif
        bool forked_dw = false;
        int l = 0;
        bool wt_ready = false, wt_forked = false;
#ifndef DLB_CUDA_EMU
        if constexpr (std::is_same_v<R, float>) {
            if (wt_enabled && tc_for(B) && B >= wt_min_batch && !wT.empty()) {
                cudaStream_t ts = st;
                if (fork) {
                    DLB_CUDA(cudaEventRecord(wt_fork_ev, st));
                    DLB_CUDA(cudaStreamWaitEvent(side_head, wt_fork_ev, 0));
                    ts = side_head; wt_forked = true;
                }
                for (int t = std::max(plan.first_fc, 1); t < Ln - 1; ++t) {
                    if (!wT[(size_t)t]) continue;
                    const auto& T = layers[(size_t)t];
                    g_tag_layer = t;
                    TagScope tag_("wT");
                    DLB_LAUNCH((transpose_kernel<R>), dim3((unsigned)ceil_div<size_t>(T.in_n, 32), (unsigned)ceil_div<size_t>(T.out_n, 32)), dim3(256), ts,
                               (const R*)(params + T.param_off + T.bias_n), wT[(size_t)t], (int)T.out_n, (int)T.in_n);
                    wt_ready = true;
                }
                if (wt_forked && wt_ready) DLB_CUDA(cudaEventRecord(wt_done_ev, side_head));
            }
        }
#endif
        if (plan.has_cp) {
            g_tag_layer = 0;
            const auto& Y = layers[0];
            bool done = false;
#ifndef DLB_CUDA_EMU
            if constexpr (std::is_same_v<R, float>) {
                if (conv28_for(B)) {   // (the index-map buffer holds the one-byte window selectors here)
                    conv28::forward(st, c28.stage, x, gather_base, cursor, params + Y.param_off, params + Y.param_off + Y.bias_n, out[1],
                                    reinterpret_cast<unsigned char*>(imap[1]), B);
                    imap1_holds_selectors = true;
                    done = true;
                }
            }
#endif
            if (!done) {
                const int grid = std::min((int)ceil_div(B, plan.cp.spc), kNumSMs);
                conv_pool_forward<R>(st, plan.cp, grid, x, gather_base, cursor, params + Y.param_off, params + Y.param_off + Y.bias_n,
                                     out[1], imap[1], B);
                imap1_holds_selectors = false;
            }
            l = 2;
        }
        // hidden NLayers
        const R* head_partials = nullptr; int head_partial_splits = 0;
        for (; l < Ln - 1; ++l) {
            const auto& Y = layers[(size_t)l];
            g_tag_layer = l;
            TagScope tag_("full_fwd");
            const R* bias = params + Y.param_off;
#ifndef DLB_CUDA_EMU
            if constexpr (std::is_same_v<R, float>) {
                if (tc_for(B)) {   // y = act(X W^T + b): both operands reduction-contiguous
                    const TcOperand A{l == 0 ? x : out[(size_t)l - 1], (long long)Y.in_n, 0, l == 0 ? gather_base : nullptr, l == 0 ? cursor : nullptr, -1};
                    const TcOperand Bo{bias + Y.bias_n, (long long)Y.in_n, 0, nullptr, nullptr, -1};
                    // behind the fused conv + pool kernel (which triggers early) the contraction is launched programmatically
                    // the last hidden layer in front of a head with at most 128 inputs leaves a split reduction unfinalised: the head sums the
                    // partials itself (fused.cuh, HeadInputPartials) - one dependent kernel less
                    const bool defer = head_folds_finalize && l == Ln - 2 && Y.out_n <= 128;
                    int splits = 1;
                    tc_gemm<kRedContig, kRedContig>(st, B, (int)Y.out_n, (int)Y.in_n, A, Bo, EpiFullForward<float>{out[(size_t)l], nullptr, bias, (int)Y.out_n, Y.s.activation}, lc.ws,
                                                    /*pdl*/ plan.has_cp && l == 2 && pdl_enabled, defer ? &splits : nullptr);
                    if (defer && splits > 1) { head_partials = static_cast<const R*>(lc.ws.ptr); head_partial_splits = splits; }
                    else if (defer) head_partial_splits = 0;   // (one split: the contraction's own epilogue wrote the activations)
                    continue;
                }
            }
#endif
            if (l == 0)
                gemm<R>(st, B, (int)Y.out_n, (int)Y.in_n, DenseRowCursorLoader<R>{x, (size_t)Y.in_n, gather_base, cursor},
                        DenseRowLoader<R>{bias + Y.bias_n, (size_t)Y.in_n, nullptr},
                        EpiFullForward<R>{out[(size_t)l], nullptr, bias, (int)Y.out_n, Y.s.activation}, lc.ws);
            else
                gemm<R>(st, B, (int)Y.out_n, (int)Y.in_n, DenseRowLoader<R>{out[(size_t)l - 1], (size_t)Y.in_n, nullptr},
                        DenseRowLoader<R>{bias + Y.bias_n, (size_t)Y.in_n, nullptr},
                        EpiFullForward<R>{out[(size_t)l], nullptr, bias, (int)Y.out_n, Y.s.activation}, lc.ws);
        }
        // head: forward + cost derivative + activation gradient + input gradient (already times the previous NLayer's derivative)
        {
            const int h = Ln - 1;
            const auto& Y = layers[(size_t)h];
            g_tag_layer = h;
            const auto& Pv = layers[(size_t)h - 1];
            const int prev_act = Pv.s.type == DLB_LAYER_FULL ? Pv.s.activation : 0;
            const R* bias = params + Y.param_off;
            const size_t smem = (size_t)Y.out_n * (size_t)Y.in_n * sizeof(R);
            const int grid = std::min((int)ceil_div(B, 8), 4 * kNumSMs);
            const HeadInputPartials<R> hp{head_partial_splits > 1 ? head_partials : nullptr, head_partial_splits, params + Pv.param_off, Pv.s.activation, out[(size_t)h - 1]};
#ifndef DLB_CUDA_EMU
            // the transposed weights of the hidden NLayers (side stream, since the start of the step) are joined here, in front of the head:
            // the input-gradient contraction that needs them is launched programmatically right behind the head kernel
            if (wt_forked && wt_ready) DLB_CUDA(cudaStreamWaitEvent(st, wt_done_ev, 0));
#endif
            {
                TagScope tag_("head");
                if (plan.head_nmax == 10)
                    DLB_LAUNCH_SMEM((full_head_kernel<R, 10>), dim3((unsigned)grid), dim3(256), smem, st, (const R*)out[(size_t)h - 1], bias,
                                    bias + Y.bias_n, refs, gather_base, cursor, out[(size_t)h], in_grad[(size_t)Ln], in_grad[(size_t)h], B,
                                    (int)Y.in_n, (int)Y.out_n, Y.s.activation, cost, prev_act, hp);
                else if (plan.head_nmax == 16)
                    DLB_LAUNCH_SMEM((full_head_kernel<R, 16>), dim3((unsigned)grid), dim3(256), smem, st, (const R*)out[(size_t)h - 1], bias,
                                    bias + Y.bias_n, refs, gather_base, cursor, out[(size_t)h], in_grad[(size_t)Ln], in_grad[(size_t)h], B,
                                    (int)Y.in_n, (int)Y.out_n, Y.s.activation, cost, prev_act, hp);
                else
                    DLB_LAUNCH_SMEM((full_head_kernel<R, 32>), dim3((unsigned)grid), dim3(256), smem, st, (const R*)out[(size_t)h - 1], bias,
                                    bias + Y.bias_n, refs, gather_base, cursor, out[(size_t)h], in_grad[(size_t)Ln], in_grad[(size_t)h], B,
                                    (int)Y.in_n, (int)Y.out_n, Y.s.activation, cost, prev_act, hp);
            }
            {
                TagScope tag_("head_dW");
                head_chunks = (int)ceil_div(B, kHeadChunk);
                const int total = (int)Y.out_n * (int)(Y.in_n + 1);
                cudaStream_t hs = st;
#ifndef DLB_CUDA_EMU
                if (fork) {
                    DLB_CUDA(cudaEventRecord(fork_events[(size_t)Ln], st));
                    DLB_CUDA(cudaStreamWaitEvent(side_head, fork_events[(size_t)Ln], 0));
                    hs = side_head;
                }
#endif
                DLB_LAUNCH((batch_outer_partial_kernel<R>), dim3((unsigned)ceil_div(total, 128), (unsigned)head_chunks), dim3(128), hs,
                           (const R*)in_grad[(size_t)Ln], (const R*)out[(size_t)h - 1], head_partial, B, (int)Y.out_n, (int)Y.in_n, kHeadChunk);
                if (want_cost) {   // the batch cost only needs the head's outputs: it rides the same side stream
                    TagScope tag2_("cost_sum");
                    DLB_LAUNCH((cost_sum_kernel<R>), dim3(1), dim3(256), hs, (const R*)out[(size_t)h], refs, resolve_host_gather(gather_base, cursor),
                               B, (int)Y.out_n, cost, cur_cost_dev ? cur_cost_dev : cost_dev);
                }
            }
        }
        // hidden NLayers backward: in_grad[l + 1] holds g_l = dC/dz_l
        for (l = Ln - 2; l >= plan.first_fc; --l) {
            const auto& Y = layers[(size_t)l];
            g_tag_layer = l;
            const R* gl = in_grad[(size_t)l + 1];
            R* db = grads + Y.param_off;
            bool done_tc = false;
            // the weight gradient of this layer only needs g_l: it runs on the side stream while the main stream goes on with dX
            cudaStream_t ws_stream = st;
            GemmWorkspace ws_for_dw = lc.ws;
#ifndef DLB_CUDA_EMU
            if (fork) {
                DLB_CUDA(cudaEventRecord(fork_events[(size_t)l], st));
                DLB_CUDA(cudaStreamWaitEvent(side_dw, fork_events[(size_t)l], 0));
                ws_stream = side_dw; ws_for_dw = ws_side; forked_dw = true;
            }
            if constexpr (std::is_same_v<R, float>) {
                if (tc_for(B)) {
                    {   // dW[n][k] = sum_b g[b][n] x[b][k], db[n] = sum_b g[b][n] (column of ones): both operands outer-contiguous
                        TagScope tag_("full_dW");
                        const TcOperand A{gl, (long long)Y.out_n, 1, nullptr, nullptr, -1};
                        const TcOperand Bo{l == 0 ? x : out[(size_t)l - 1], (long long)Y.in_n, 1, l == 0 ? gather_base : nullptr,
                                           l == 0 ? cursor : nullptr, (int)Y.in_n};
                        tc_gemm<kOuterContig, kOuterContig>(ws_stream, (int)Y.out_n, (int)Y.in_n + 1, B, A, Bo, EpiWeightAndBias<float>{db, db + Y.bias_n, (int)Y.in_n}, ws_for_dw);
                    }
                    if (l > 0) {   // dx[b][k] = sum_n g[b][n] W[n][k]: A reduction-contiguous, B (= W) outer-contiguous
                        TagScope tag_("full_dX");
                        const auto& Pv = layers[(size_t)l - 1];
                        const int prev_act = Pv.s.type == DLB_LAYER_FULL ? Pv.s.activation : 0;
                        const TcOperand A{gl, (long long)Y.out_n, 0, nullptr, nullptr, -1};
                        const EpiStoreTimesDeriv<float> epi{in_grad[(size_t)l], (size_t)Y.in_n, out[(size_t)l - 1], prev_act};
                        const bool pdl = /*the head kernel is the previous kernel on this stream and triggers early*/ l == Ln - 2 && pdl_enabled;
                        if (wt_ready && wT[(size_t)l]) {   // W^T [in_n][out_n]: both operands reduction-contiguous (K-major plan, TMA where the rows are 16-byte multiples)
                            const TcOperand Bo{wT[(size_t)l], (long long)Y.out_n, 0, nullptr, nullptr, -1};
                            tc_gemm<kRedContig, kRedContig>(st, B, (int)Y.in_n, (int)Y.out_n, A, Bo, epi, lc.ws, pdl);
                        } else {
                            const TcOperand Bo{params + Y.param_off + Y.bias_n, (long long)Y.in_n, 1, nullptr, nullptr, -1};
                            tc_gemm<kRedContig, kOuterContig>(st, B, (int)Y.in_n, (int)Y.out_n, A, Bo, epi, lc.ws, pdl);
                        }
                    }
                    done_tc = true;
                }
            }
#endif
            if (done_tc) continue;
            {
                TagScope tag_("full_dW");
                if (l == 0)
                    gemm<R>(ws_stream, (int)Y.out_n, (int)Y.in_n + 1, B, DenseColLoader<R>{gl, (size_t)Y.out_n, nullptr},
                            DenseColOnesLoader<R>{x, (size_t)Y.in_n, gather_base, cursor}, EpiWeightAndBias<R>{db, db + Y.bias_n, (int)Y.in_n}, ws_for_dw);
                else
                    gemm<R>(ws_stream, (int)Y.out_n, (int)Y.in_n + 1, B, DenseColLoader<R>{gl, (size_t)Y.out_n, nullptr},
                            DenseColOnesLoader<R>{out[(size_t)l - 1], (size_t)Y.in_n, nullptr, nullptr},
                            EpiWeightAndBias<R>{db, db + Y.bias_n, (int)Y.in_n}, ws_for_dw);
            }
            if (l > 0) {
                TagScope tag_("full_dX");
                const auto& Pv = layers[(size_t)l - 1];
                const int prev_act = Pv.s.type == DLB_LAYER_FULL ? Pv.s.activation : 0;
                gemm<R>(st, B, (int)Y.in_n, (int)Y.out_n, DenseRowLoader<R>{gl, (size_t)Y.out_n, nullptr},
                        DenseColLoader<R>{params + Y.param_off + Y.bias_n, (size_t)Y.in_n, nullptr},
                        EpiStoreTimesDeriv<R>{in_grad[(size_t)l], (size_t)Y.in_n, out[(size_t)l - 1], prev_act}, lc.ws);
            }
        }
        if (plan.has_cp) {
            g_tag_layer = 0;
            bool done = false;
#ifndef DLB_CUDA_EMU
            if constexpr (std::is_same_v<R, float>) {
                if (conv28_for(B)) {
                    cp_rows = conv28::backward(st, c28.stage, x, gather_base, cursor, (const float*)in_grad[2], reinterpret_cast<const unsigned char*>(imap[1]),
                                               (const float*)out[1], cp_partial, cp_grid, B);
                    done = true;
                }
            }
#endif
            if (!done) {
                const int grid = std::min(B, cp_grid);
                if (grid < cp_grid)   // CTAs that do not run this time must not leave stale partial sums behind
                    DLB_CUDA(cudaMemsetAsync(cp_partial + (size_t)grid * plan.cp.n_par, 0, (size_t)(cp_grid - grid) * plan.cp.n_par * sizeof(R), st));
                conv_pool_backward<R>(st, plan.cp, grid, x, gather_base, cursor, (const R*)in_grad[2], (const int*)imap[1], (const R*)out[1], cp_partial, B);
                cp_rows = cp_grid;
            }
        }
#ifndef DLB_CUDA_EMU
        if (fork) {   // join: the reduce/update that follows reads every gradient
            DLB_CUDA(cudaEventRecord(fork_events[(size_t)Ln + 1], side_head));
            DLB_CUDA(cudaStreamWaitEvent(st, fork_events[(size_t)Ln + 1], 0));
            if (forked_dw) {
                DLB_CUDA(cudaEventRecord(fork_events[(size_t)Ln + 2], side_dw));
                DLB_CUDA(cudaStreamWaitEvent(st, fork_events[(size_t)Ln + 2], 0));
            }
        }
#endif
        (void)forked_dw;
        last_B = B;
    }

    // Sums the partial gradients into the flat gradient buffer and (optionally) updates the parameters in the same pass.
    void train_reduce_update(bool do_update, double lr, double reg_factor, int64_t batch_items, int advance, bool publish = false) {
        ReduceSegments rs{};
        rs.count = 0;
        for (int l = 0; l < L(); ++l) {
            const auto& Y = layers[(size_t)l];
            if (Y.bias_n + Y.weight_n == 0) continue;
            const bool conv_stage = plan.has_cp && l == 0;
            const bool head = l == L() - 1;
            for (int part = 0; part < 2; ++part) {
                const int i = rs.count++;
                rs.end[i] = Y.param_off + Y.bias_n + (part ? Y.weight_n : 0);
                rs.regularised[i] = part ? 1 : 0;
                rs.partial[i] = nullptr; rs.chunks[i] = 0; rs.stride[i] = 0; rs.local_off[i] = part ? Y.bias_n : 0;
                if (conv_stage) { rs.partial[i] = cp_partial; rs.chunks[i] = cp_rows > 0 ? cp_rows : cp_grid; rs.stride[i] = plan.cp.n_par; }
                else if (head) { rs.partial[i] = head_partial; rs.chunks[i] = head_chunks; rs.stride[i] = Y.out_n * (Y.in_n + 1); }
            }
        }
        const R l_rate = -static_cast<R>(lr);
        const R reg = static_cast<R>(reg_factor);
        const R grad_norm = static_cast<R>(1) / static_cast<R>(batch_items);
        g_tag_layer = -1;
        TagScope tag_("reduce_update");
        long long n_partial = 0;
        for (int i = 0; i < rs.count; ++i)
            if (rs.partial[i]) n_partial += rs.end[i] - (i == 0 ? 0 : rs.end[i - 1]);
        const int partial_blocks = (int)ceil_div<long long>(n_partial, 16);
        DLB_LAUNCH((reduce_update_kernel<R>), dim3((unsigned)(partial_blocks + ceil_div<size_t>((size_t)n_params, 256))), dim3(256), ctx->stream,
                   params, grads, publish ? publish_buf(0) : (R*)nullptr, (size_t)kMaxPeers * p2p_publish_bytes / sizeof(R), (const unsigned long long*)d_p2p_step,
                   (size_t)n_params, rs, partial_blocks, n_partial, do_update ? 1 : 0, static_cast<R>(1) + reg * l_rate,
                   l_rate * grad_norm, reg > static_cast<R>(0) ? 1 : 0, advance ? d_cursor : (int*)nullptr, advance);
    }

    // Data-parallel tail of a step whose local gradient sum is (or will be put) in `grads`: peer-memory kernel or NCCL.
    // local_reduce: for the fused path, runs the partial-sum reduction into the published buffer first.
    void distributed_finish(bool fused_partials, double lr, double reg_factor, int64_t global, int advance) {
#ifndef DLB_CUDA_EMU
        if (p2p_on) {
            const int parity = (int)(p2p_step & 1);
            ++p2p_step;
            if (fused_partials) train_reduce_update(false, lr, reg_factor, global, 0, true);
            else DLB_CUDA(cudaMemcpyAsync(publish_buf(parity), grads, (size_t)n_params * sizeof(R), cudaMemcpyDeviceToDevice, ctx->stream));
            PeerTable pt{};
            for (int r = 0; r < world; ++r) pt.region[r] = static_cast<const unsigned char*>(p2p_peer_regions[(size_t)r]);
            pt.publish_bytes = p2p_publish_bytes;
            const R l_rate = -static_cast<R>(lr);
            const R reg = static_cast<R>(reg_factor);
            const R grad_norm = static_cast<R>(1) / static_cast<R>(global);
            g_tag_layer = -1;
            TagScope tag_("p2p_reduce_update");
            DLB_LAUNCH((p2p_reduce_update_kernel<R>), dim3(2 * kNumSMs), dim3(256), ctx->stream, params, grads, (size_t)n_params, pt, world, rank,
                       d_p2p_step, reinterpret_cast<unsigned int*>(d_p2p_step + 1), segs, static_cast<R>(1) + reg * l_rate, l_rate * grad_norm,
                       reg > static_cast<R>(0) ? 1 : 0, advance ? d_cursor : (int*)nullptr, advance, p2p_push ? 1 : 0, p2p_prof);
            return;
        }
#endif
        if (fused_partials) train_reduce_update(false, lr, reg_factor, global, 0);
        allreduce_gradient();
        update(lr, reg_factor, global, advance);
    }

    void allreduce_gradient() {
#ifndef DLB_CUDA_EMU
        if (world <= 1) return;
        auto& api = NcclApi::get();
        api.check(api.AllReduce(grads, grads, (size_t)n_params, sizeof(R) == 4 ? kNcclFloat : kNcclDouble, kNcclSum, comm, ctx->stream),
                  "ncclAllReduce");
#endif
    }

    // ALayer<D>::update for every layer (Net.inl:309-316): l_rate = -lr, grad_scale_factor = 1 / B_actual.
    void update(double lr, double reg_factor, int64_t batch_items, int advance = 0) {
        if (advance) {   // keep the device-side cursor in step (the unfused path reads its gather slice from the host-computed pointer)
            ReduceSegments none{}; none.count = 1; none.end[0] = 0;
            DLB_LAUNCH((reduce_update_kernel<R>), dim3(1), dim3(256), ctx->stream, params, grads, (R*)nullptr, (size_t)0,
                       (const unsigned long long*)nullptr, (size_t)0, none, 0, (long long)0, 0, R(0), R(0), 0, d_cursor, advance);
        }
        const R l_rate = -static_cast<R>(lr);
        const R reg = static_cast<R>(reg_factor);
        const R grad_norm = static_cast<R>(1) / static_cast<R>(batch_items);
        const R param_scale = static_cast<R>(1) + reg * l_rate;
        const R grad_total_scale = l_rate * grad_norm;
        g_tag_layer = -1;
        sgd_update<R>(ctx->lc(), params, grads, (size_t)n_params, segs, param_scale, grad_total_scale, reg > static_cast<R>(0));
    }

    // One step on this rank's share [first, first + mine) of a global batch of `global` items.
    // gather_base + (*cursor if cursor) addresses this rank's slice; `advance` is added to the cursor at the end of the step.
    void step_device(const R* x, const R* refs, const int* gather_base, const int* cursor, int advance, int64_t mine, int64_t global,
                     double lr, int cost, double reg_factor) {
        draw_masks();   // one set of keep masks per (global) mini-batch; every rank draws the same ones from the same seed
        if (mega_for(mine)) {
            ensure_capacity(mine);
            mega_step((int)mine, x, refs, gather_base, cursor, advance, cost, lr, reg_factor, global, nullptr);
            return;
        }
        if (plan.ok && mine > 0) {
            ensure_capacity(mine);
            auto body = [&] {
                train_forward_backward((int)mine, x, refs, gather_base, cursor, cost);
                if (world <= 1) train_reduce_update(true, lr, reg_factor, global, advance);
                else distributed_finish(true, lr, reg_factor, global, advance);
            };
            // Only paths whose buffers are fixed replay a graph: the resident-data path (cursor-relative gather).
            if (cursor && x == data_x) { run_graphed(StepKey{mine, global, lr, reg_factor, cost, x, p2p_on ? -1 : 0, gather_base}, body); return; }
            body();
            return;
        }
        if (mine > 0) {
            ensure_capacity(mine);
            forward_backward((int)mine, x, refs, gather_base, cost, R(0));
        } else {
            DLB_CUDA(cudaMemsetAsync(grads, 0, (size_t)n_params * sizeof(R), ctx->stream));
        }
        if (world > 1) distributed_finish(false, lr, reg_factor, global, advance);
        else update(lr, reg_factor, global, advance);
    }

    void rank_slice(int64_t count, int64_t& first, int64_t& mine) const {
        const int64_t per = (count + world - 1) / world;
        first = std::min<int64_t>(count, (int64_t)rank * per);
        mine = std::min<int64_t>(count, first + per) - first;
    }

    // The mini-batch [start, end) of the current epoch permutation (perm_dev); d_cursor == start on the device.
    void run_batch(int64_t start, int64_t end, double lr, int cost, double reg_factor) {
        int64_t first, mine;
        rank_slice(end - start, first, mine);
        if (plan.ok) step_device(data_x, data_y, perm_dev + first, d_cursor, (int)(end - start), mine, end - start, lr, cost, reg_factor);
        else step_device(data_x, data_y, perm_dev + start + first, nullptr, 0, mine, end - start, lr, cost, reg_factor);
    }

    void learn(int64_t batch, int64_t epochs, double lr, int cost, double lambda, dlb_epoch_callback cb, void* user) override {
        DLB_REQUIRE(data_x && data_y, "Incompatible collection of training and reference items.");
        if (n_items == 0 || batch == 0) return;   // Net.inl:200-201
        DLB_REQUIRE(batch > 0 && epochs >= 0, "Invalid batch size / epoch count");
        use_device(ctx);
        const R reg_factor = static_cast<R>(lambda) / static_cast<R>(n_items);   // Net.inl:203
        std::vector<std::size_t> index_map((size_t)n_items);
        std::iota(index_map.begin(), index_map.end(), 0);                         // Net.inl:164-171
        ensure_perm(n_items);
        for (int64_t e = 0; e < epochs; ++e) {
            std::ranges::shuffle(index_map, rng);                                 // Net.inl:158-161, 223
            upload_permutation(index_map);
            int64_t start = 0;
            while (start < n_items) {
                // Net.inl:235-236: fewer than 1.5 * batch_size remaining items are taken as one batch
                const int64_t end = (double)(n_items - start) < 1.5 * (double)batch ? n_items : start + batch;
                if (mega_for(end - start) && mg.multi && end - start == batch) {   // all full mini-batches of the epoch in ONE launch of the step kernel
                    int64_t k = 1;
                    while ((double)(n_items - (start + k * batch)) >= 1.5 * (double)batch) ++k;
                    ensure_capacity(batch);
                    mega_step((int)batch, data_x, data_y, perm_dev, d_cursor, (int)batch, cost, lr, (double)reg_factor, batch, nullptr, (int)k);
                    start += k * batch;
                    continue;
                }
                run_batch(start, end, lr, cost, (double)reg_factor);
                start = end;
            }
            if (cb) { DLB_CUDA(cudaStreamSynchronize(ctx->stream)); cb(e, 0.5 * (double)reg_factor, user); }   // Net.inl:319
        }
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
    }

    // Net::learn's loop made resumable: runs the next `n_steps` mini-batches of the reference's epoch sequence
    // (reshuffling when an epoch is exhausted) and returns the number of items consumed by all ranks.
    std::vector<std::size_t> ts_index_map;
    int64_t ts_cursor = -1;
    int64_t train_steps(int64_t batch, int64_t n_steps, double lr, int cost, double lambda) override {
        DLB_REQUIRE(data_x && data_y, "Incompatible collection of training and reference items.");
        DLB_REQUIRE(batch > 0 && n_steps >= 0, "Invalid batch size / step count");
        use_device(ctx);
        const R reg_factor = static_cast<R>(lambda) / static_cast<R>(n_items);
        if (perm_cap < n_items) { ensure_perm(n_items); ts_cursor = -1; }
        if ((int64_t)ts_index_map.size() != n_items) { ts_index_map.resize((size_t)n_items); std::iota(ts_index_map.begin(), ts_index_map.end(), 0); ts_cursor = -1; }
        int64_t consumed = 0;
        for (int64_t s = 0; s < n_steps; ++s) {
            if (ts_cursor < 0 || ts_cursor >= n_items) {
                std::ranges::shuffle(ts_index_map, rng);
                upload_permutation(ts_index_map);
                ts_cursor = 0;
            }
            const int64_t start = ts_cursor;
            const int64_t end = (double)(n_items - start) < 1.5 * (double)batch ? n_items : start + batch;
            if (mega_for(end - start) && mg.multi && end - start == batch) {
                // consecutive full mini-batches of this epoch run inside ONE launch of the step kernel (the batch position lives in d_cursor)
                int64_t k = 1;
                while (s + k < n_steps && (double)(n_items - (start + k * batch)) >= 1.5 * (double)batch) ++k;
                ensure_capacity(batch);
                mega_step((int)batch, data_x, data_y, perm_dev, d_cursor, (int)batch, cost, lr, (double)reg_factor, batch, nullptr, (int)k);
                consumed += k * batch;
                ts_cursor = start + k * batch;
                s += k - 1;
                continue;
            }
            run_batch(start, end, lr, cost, (double)reg_factor);
            consumed += end - start;
            ts_cursor = end;
        }
        return consumed;
    }

    void step_indices(const int64_t* idx, int64_t n, double lr, int cost, double reg_factor) override {
        DLB_REQUIRE(data_x && data_y && idx && n > 0, "Invalid input.");
        use_device(ctx);
        int64_t first, mine;
        rank_slice(n, first, mine);
        ensure_capacity(std::max<int64_t>(mine, 1));
        std::vector<int> tmp((size_t)std::max<int64_t>(mine, 1));
        for (int64_t i = 0; i < mine; ++i) {
            DLB_REQUIRE(idx[first + i] >= 0 && idx[first + i] < n_items, "Item index out of range");
            tmp[(size_t)i] = (int)idx[first + i];
        }
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
        DLB_CUDA(cudaMemcpyAsync(gather_dev, tmp.data(), (size_t)mine * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));   // tmp is pageable and goes out of scope
        step_device(data_x, data_y, gather_dev, nullptr, 0, mine, n, lr, cost, reg_factor);
    }

    void stage_host_batch(const void* x, const void* y, int64_t n) {
        ensure_capacity(n);
        DLB_CUDA(cudaMemcpyAsync(stage_x, x, (size_t)n * (size_t)in0() * sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
        if (y) DLB_CUDA(cudaMemcpyAsync(stage_y, y, (size_t)n * (size_t)outN() * sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
    }

    // One step on host buffers.  The copy of this batch runs on the copy stream into the half of the staging double buffer
    // that the step before the previous one used, so it overlaps the previous step's kernels; the step itself (a CUDA graph
    // per buffer half on the fused path) and the device->host copy of the batch cost are enqueued on the main stream.
    // sync == false leaves everything in flight (dlb_net_step_host_async); the caller synchronises the context.
    void step_host(const void* x, const void* y, int64_t n, double lr, int cost, double reg_factor, double* cost_out, bool sync) override {
        DLB_REQUIRE(x && y && n > 0, "Invalid input.");
        use_device(ctx);
        ensure_capacity(n);
        if (n > hcap) {
            DLB_CUDA(cudaStreamSynchronize(ctx->stream));
#ifndef DLB_CUDA_EMU
            DLB_CUDA(cudaStreamSynchronize(side_copy));
#endif
            drop_graphs();
            for (int i = 0; i < 2; ++i) {
                dev_free(hstage_x[i]); dev_free(hstage_y[i]);
                hstage_x[i] = dev_alloc<R>((size_t)n * (size_t)in0());
                hstage_y[i] = dev_alloc<R>((size_t)n * (size_t)outN());
                hbuf_used[i] = false;
            }
            hcap = n;
        }
        const int buf = hbuf; hbuf ^= 1;
        R* sx = hstage_x[buf]; R* sy = hstage_y[buf];
        cudaStream_t cs = ctx->stream;
#ifndef DLB_CUDA_EMU
        cs = side_copy;
        if (hbuf_used[buf]) DLB_CUDA(cudaStreamWaitEvent(cs, ev_consumed[buf], 0));   // the step that last read this half is done
        if (cost_read_pending[buf]) DLB_CUDA(cudaStreamWaitEvent(cs, ev_cost_read[buf], 0));   // ... and its cost has left its slot
        cur_cost_dev = cost_slot[buf];
#endif
        DLB_CUDA(cudaMemcpyAsync(sx, x, (size_t)n * (size_t)in0() * sizeof(R), cudaMemcpyHostToDevice, cs));
        DLB_CUDA(cudaMemcpyAsync(sy, y, (size_t)n * (size_t)outN() * sizeof(R), cudaMemcpyHostToDevice, cs));
#ifndef DLB_CUDA_EMU
        DLB_CUDA(cudaEventRecord(ev_copied[buf], cs));
        DLB_CUDA(cudaStreamWaitEvent(ctx->stream, ev_copied[buf], 0));
#endif
        const bool want_cost = cost_out != nullptr;
        draw_masks();
        auto body = [&] {
            if (plan.ok) train_forward_backward((int)n, sx, sy, nullptr, nullptr, cost, want_cost);
            else forward_backward((int)n, sx, sy, nullptr, cost, R(0));
            if (want_cost && !plan.ok) {
                g_tag_layer = L();
                TagScope tag_("cost_sum");
                DLB_LAUNCH((cost_sum_kernel<R>), dim3(1), dim3(256), ctx->stream, (const R*)out[(size_t)L() - 1], (const R*)sy,
                           (const int*)nullptr, (int)n, (int)outN(), cost, cur_cost_dev ? cur_cost_dev : cost_dev);
            }
            if (world > 1) distributed_finish(plan.ok, lr, reg_factor, n * world, 0);
            else if (plan.ok) train_reduce_update(true, lr, reg_factor, n, 0);
            else update(lr, reg_factor, n);
        };
        if (mega_for(n)) mega_step((int)n, sx, sy, nullptr, nullptr, 0, cost, lr, reg_factor, n, want_cost ? (cur_cost_dev ? cur_cost_dev : cost_dev) : nullptr);
        else if (plan.ok) run_graphed(StepKey{n, n * world, lr, reg_factor, cost, sx, want_cost ? 1 : 2}, body);
        else body();
#ifndef DLB_CUDA_EMU
        DLB_CUDA(cudaEventRecord(ev_consumed[buf], ctx->stream));
        hbuf_used[buf] = true;
        if (cost_out) {   // off the main stream: the copy waits for this step, the next user of this half waits for the copy
            DLB_CUDA(cudaStreamWaitEvent(side_d2h, ev_consumed[buf], 0));
            DLB_CUDA(cudaMemcpyAsync(cost_out, cur_cost_dev, sizeof(double), cudaMemcpyDeviceToHost, side_d2h));
            DLB_CUDA(cudaEventRecord(ev_cost_read[buf], side_d2h));
            cost_read_pending[buf] = true;
        }
        cur_cost_dev = nullptr;
        if (sync) { DLB_CUDA(cudaStreamSynchronize(ctx->stream)); DLB_CUDA(cudaStreamSynchronize(side_d2h)); }
#else
        if (cost_out) DLB_CUDA(cudaMemcpyAsync(cost_out, cost_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (sync) DLB_CUDA(cudaStreamSynchronize(ctx->stream));
#endif
    }

    void forward_backward_host(const void* x, const void* y, int64_t n, int cost, double scale) override {
        DLB_REQUIRE(x && y && n > 0, "Invalid input.");
        use_device(ctx);
        stage_host_batch(x, y, n);
        forward_backward((int)n, stage_x, stage_y, nullptr, cost, static_cast<R>(scale));
    }

    void read(int what, int layer, void* dst, int64_t bytes) override {
        use_device(ctx);
        const void* src = nullptr;
        size_t need = 0;
        const size_t b = (size_t)last_B;
        auto check_layer = [&](int lo) { DLB_REQUIRE(layer >= lo && layer < L(), "Invalid layer index"); };
#ifndef DLB_CUDA_EMU
        if (what > 100) {   // development only (tools/mega_debug.py): the megakernel's scratch buffers
            DLB_REQUIRE(mg.runner != nullptr, "the megakernel has not run");
            mega::runner_debug_read(mg.runner, what, dst, (size_t)bytes, ctx->stream);
            return;
        }
#endif
        switch (what) {
        case DLB_BUF_OUTPUT: check_layer(0); src = out[(size_t)layer]; need = b * (size_t)layers[(size_t)layer].out_n * sizeof(R); break;
        case DLB_BUF_DERIVATIVES: check_layer(0); src = deriv[(size_t)layer]; need = b * (size_t)layers[(size_t)layer].out_n * sizeof(R); break;
        case DLB_BUF_INDEX_MAP: check_layer(0); src = imap[(size_t)layer]; need = b * (size_t)layers[(size_t)layer].out_n * sizeof(int); break;
        case DLB_BUF_INPUT_GRAD: check_layer(1); src = in_grad[(size_t)layer]; need = b * (size_t)layers[(size_t)layer].in_n * sizeof(R); break;
        case DLB_BUF_COST_GRAD: src = in_grad[(size_t)L()]; need = b * (size_t)outN() * sizeof(R); break;
        case DLB_BUF_GRADIENT: src = grads; need = (size_t)n_params * sizeof(R); break;
        default: throw arg_error("Unknown buffer id");
        }
        DLB_REQUIRE(what == DLB_BUF_GRADIENT || last_B > 0, "No forward/backward pass has run yet");
        DLB_REQUIRE((size_t)bytes >= need, "Destination buffer too small");
        if (what == DLB_BUF_INDEX_MAP && layer == 1 && imap1_holds_selectors) {
            // the large-batch first stage keeps one byte per window (position 0..3 inside the 2 x 2 window, 255 = none): expand it to the
            // reference's Trace.IndexMapping - the flat index of the extremum inside the PLayer's input (PLayer.inl:115-138)
            const auto& P = layers[1].s;
            const int OW = P.in_w, POW = P.in_w / 2, NPOS = (P.in_h / 2) * POW, F = P.in_c;
            const size_t n = b * (size_t)layers[1].out_n;
            std::vector<unsigned char> sel(n);
            DLB_CUDA(cudaMemcpyAsync(sel.data(), src, n, cudaMemcpyDeviceToHost, ctx->stream));
            DLB_CUDA(cudaStreamSynchronize(ctx->stream));
            int* out_map = static_cast<int*>(dst);
            for (size_t i = 0; i < n; ++i) {
                const int w = (int)(i % (size_t)NPOS), f = (int)((i / (size_t)NPOS) % (size_t)F);
                const int py = w / POW, px = w - py * POW, q = sel[i];
                out_map[i] = q > 3 ? -1 : (f * P.in_h + 2 * py + (q >> 1)) * OW + 2 * px + (q & 1);
            }
            return;
        }
        DLB_CUDA(cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, ctx->stream));
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
    }

    void write_gradient(const void* src) override {
        use_device(ctx);
        DLB_CUDA(cudaMemcpyAsync(grads, src, (size_t)n_params * sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
    }

    // Net<D>::evaluate_cost_function_and_correct_answers, Net.inl:411-439: Cost = sum_i cost(act(x_i), ref_i) / n +
    // l2_reg_factor * squared_weights_sum(), CorrectAnswers = #(argmax match) / n.
    void evaluate_host(const void* x, const void* y, int64_t n, int cost, double l2_reg_factor, double* cost_out, double* correct_out) override {
        DLB_REQUIRE(x && y && n > 0, "Invalid input.");
        use_device(ctx);
        const int64_t chunk = std::min<int64_t>(n, 4096);
        ensure_capacity(chunk);
        const R* hx = static_cast<const R*>(x);
        const R* hy = static_cast<const R*>(y);
        double* acc = dev_alloc<double>(2);
        int* correct = reinterpret_cast<int*>(acc + 1);
        double total_cost = 0.0;
        DLB_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), ctx->stream));
        for (int64_t s0 = 0; s0 < n; s0 += chunk) {
            const int64_t m = std::min(chunk, n - s0);
            DLB_CUDA(cudaMemcpyAsync(stage_x, hx + s0 * in0(), (size_t)m * (size_t)in0() * sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
            DLB_CUDA(cudaMemcpyAsync(stage_y, hy + s0 * outN(), (size_t)m * (size_t)outN() * sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
            forward((int)m, stage_x, nullptr, false);
            const R* out_last = out[(size_t)L() - 1];
            DLB_LAUNCH((cost_sum_kernel<R>), dim3(1), dim3(256), ctx->stream, out_last, (const R*)stage_y, (const int*)nullptr, (int)m, (int)outN(), cost, cost_dev);
            DLB_LAUNCH((count_correct_kernel<R>), dim3((unsigned)ceil_div<int64_t>(m, 128)), dim3(128), ctx->stream, out_last, (const R*)stage_y, (int)m, (int)outN(), correct);
            double c = 0.0;
            DLB_CUDA(cudaMemcpyAsync(&c, cost_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            DLB_CUDA(cudaStreamSynchronize(ctx->stream));
            total_cost += c;
        }
        for (const auto& Y : layers)   // squared_weights_sum: NLayer weights and CLayer filters, not the biases
            if (Y.weight_n > 0)
                DLB_LAUNCH((sum_of_squares_kernel<R>), dim3(64), dim3(256), ctx->stream, (const R*)(params + Y.param_off + Y.bias_n), (size_t)Y.weight_n, acc);
        double sq = 0.0; int ok = 0;
        DLB_CUDA(cudaMemcpyAsync(&sq, acc, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        DLB_CUDA(cudaMemcpyAsync(&ok, correct, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
        dev_free(acc);
        if (cost_out) *cost_out = total_cost / (double)n + l2_reg_factor * sq;
        if (correct_out) *correct_out = (double)ok / (double)n;
    }

    std::string describe_plan() override {
        if (!plan.ok) return "unfused: every layer runs the generic operators (layers.cuh)";
        if (mega_for(kMegaMinBatch)) return "fused: one cooperative kernel per step (mega.cu) from mini-batch 128 up: conv_pool(5x5) + hidden NLayer on tcgen05 + head + update";
        std::string d = "fused:";
        if (plan.has_cp)
            d += " conv_pool(" + std::to_string(plan.cp.KH) + "x" + std::to_string(plan.cp.KW) + ", fpt " + std::to_string(plan.cp.fpt) +
                 ", wpl " + std::to_string(plan.cp.wpl) + ")";
        d += " + " + std::to_string(L() - 1 - plan.first_fc) + " hidden NLayer(s) + head(nmax " + std::to_string(plan.head_nmax) + ")";
        if (c28.ok && c28.mode != 0)
            d += "; conv stage " + (c28.mode == 1 ? std::string("always") : "from mini-batch " + std::to_string(c28.min_batch)) + ": dense FFMA2 kernels (conv28.cu)";
        return d;
    }
    void enable_graphs(bool on) override { graphs_enabled = on; if (!on) drop_graphs(); }
    void enable_tensor_cores(int mode) override { tc_mode = mode; drop_graphs(); }
    void enable_step_kernel(int mode) override { mg.enabled = mode != 0; }

    // Allocates the shared region [parity 0: kMaxPeers slots | parity 1: kMaxPeers slots | flags] and returns its CUDA IPC handle (64 bytes).
    void p2p_export(void* handle64) override {
#ifdef DLB_CUDA_EMU
        (void)handle64; throw std::runtime_error("peer memory is not available");
#else
        use_device(ctx);
        if (!p2p_region) {
            p2p_publish_bytes = (((size_t)n_params * sizeof(R)) + 255) / 256 * 256;
            const size_t total = 2 * (size_t)kMaxPeers * p2p_publish_bytes + (kMaxPeers + 1) * sizeof(unsigned long long);
            DLB_CUDA(cudaMalloc((void**)&p2p_region, total));
            DLB_CUDA(cudaMemset(p2p_region, 0, total));
            d_p2p_step = dev_alloc<unsigned long long>(2);   // [step counter | blocks-done counter of the peer kernel]
            DLB_CUDA(cudaMemset(d_p2p_step, 0, 2 * sizeof(unsigned long long)));
        }
        cudaIpcMemHandle_t h;
        DLB_CUDA(cudaIpcGetMemHandle(&h, p2p_region));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        std::memcpy(handle64, &h, sizeof h);
#endif
    }

    // handles: world x 64 bytes, index = rank (every rank's p2p_export result, gathered by the launcher's own plumbing).
    void p2p_connect(int rank_, int world_, const void* handles) override {
#ifdef DLB_CUDA_EMU
        (void)rank_; (void)world_; (void)handles; throw std::runtime_error("peer memory is not available");
#else
        DLB_REQUIRE(p2p_region, "call dlb_net_p2p_export first");
        DLB_REQUIRE(world_ >= 1 && world_ <= kMaxPeers && rank_ >= 0 && rank_ < world_ && handles, "Invalid rank / world size");
        use_device(ctx);
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
        drop_graphs();
        p2p_peer_regions.assign((size_t)world_, nullptr);
        for (int r = 0; r < world_; ++r) {
            if (r == rank_) { p2p_peer_regions[(size_t)r] = p2p_region; continue; }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, static_cast<const unsigned char*>(handles) + 64 * (size_t)r, sizeof h);
            DLB_CUDA(cudaIpcOpenMemHandle(&p2p_peer_regions[(size_t)r], h, cudaIpcMemLazyEnablePeerAccess));
        }
        rank = rank_; world = world_; p2p_on = world_ > 1; p2p_step = 0;
        if (std::getenv("DLB_P2P_PROF") && !p2p_prof) { p2p_prof = dev_alloc<unsigned long long>(4); DLB_CUDA(cudaMemset(p2p_prof, 0, 4 * sizeof(unsigned long long))); }
        DLB_CUDA(cudaMemset(d_p2p_step, 0, 2 * sizeof(unsigned long long)));
        DLB_CUDA(cudaMemset(p2p_region + 2 * (size_t)kMaxPeers * p2p_publish_bytes, 0, (kMaxPeers + 1) * sizeof(unsigned long long)));
#endif
    }

    void eval_host(const void* x, int64_t n, int32_t* classes, void* outputs) override {
        DLB_REQUIRE(x && n >= 0, "Invalid input.");
        use_device(ctx);
        const int64_t chunk = std::min<int64_t>(std::max<int64_t>(n, 1), 4096);
        ensure_capacity(chunk);
        const R* hx = static_cast<const R*>(x);
        for (int64_t s = 0; s < n; s += chunk) {
            const int64_t m = std::min(chunk, n - s);
            DLB_CUDA(cudaMemcpyAsync(stage_x, hx + s * in0(), (size_t)m * (size_t)in0() * sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
            forward((int)m, stage_x, nullptr, false);
            const R* y = out[(size_t)L() - 1];
            if (classes) {
                DLB_LAUNCH((argmax_rows_kernel<R>), dim3((unsigned)ceil_div<int64_t>(m, 128)), dim3(128), ctx->stream, y, gather_dev, (int)m, (int)outN());
                DLB_CUDA(cudaMemcpyAsync(classes + s, gather_dev, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            }
            if (outputs)
                DLB_CUDA(cudaMemcpyAsync(static_cast<R*>(outputs) + s * outN(), y, (size_t)m * (size_t)outN() * sizeof(R), cudaMemcpyDeviceToHost, ctx->stream));
            DLB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
};

// Scratch big enough for the op-level entry points.
void ensure_op_ws(dlb_context* ctx) { ctx->ensure_ws(64u << 20); }

template <class R>
void op_sgd_update(dlb_context* ctx, R* p, const R* gr, int64_t n, int regularised, double lr, double reg_factor, double gsf) {
    DLB_REQUIRE(p && gr && n >= 0, "Collections must be of the same size");
    use_device(ctx);
    UpdateSegments segs{}; segs.count = 1; segs.end[0] = n; segs.regularised[0] = regularised ? 1 : 0;
    const R l_rate = static_cast<R>(lr);   // ALayer::update receives -learning_rate from Net (Net.inl:315); callers pass l_rate as is
    const R reg = static_cast<R>(reg_factor);
    sgd_update<R>(ctx->lc(), p, gr, (size_t)n, segs, static_cast<R>(1) + reg * l_rate, l_rate * static_cast<R>(gsf), reg > static_cast<R>(0));
}

void check_spec(const dlb_layer_spec* s, int type) {
    DLB_REQUIRE(s && s->type == type, "Layer spec of the wrong type");
    int64_t np = 0;
    (void)build_layers(s, 1, np);
}

}  // namespace

// =================================================================================================================
// extern "C"
// =================================================================================================================
extern "C" {

int dlb_version(void) { return 100; }
const char* dlb_last_error(void) { return g_last_error.c_str(); }

int dlb_context_create(int device, void* external_stream, dlb_context** out) {
    return guarded([&] {
        DLB_REQUIRE(out, "null output pointer");
        int count = 0;
        DLB_CUDA(cudaGetDeviceCount(&count));
        DLB_REQUIRE(device >= 0 && device < count, "No such CUDA device (libdlb200 has no CPU fallback)");
        DLB_CUDA(cudaSetDevice(device));
        auto ctx = std::make_unique<dlb_context>();
        ctx->device = device;
        if (external_stream) { ctx->stream = (cudaStream_t)external_stream; ctx->own_stream = false; }
        else { DLB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
        *out = ctx.release();
    });
}

int dlb_context_destroy(dlb_context* ctx) {
    return guarded([&] {
        if (!ctx) return;
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->ws.ptr) cudaFree(ctx->ws.ptr);
        if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
    });
}

int dlb_context_synchronize(dlb_context* ctx) {
    return guarded([&] {
        DLB_REQUIRE(ctx, "null context");
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (cudaStream_t s : ctx->aux_streams) DLB_CUDA(cudaStreamSynchronize(s));
    });
}
void* dlb_context_stream(dlb_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int dlb_context_device(dlb_context* ctx) { return ctx ? ctx->device : -1; }
int64_t dlb_context_launch_count(dlb_context*) { return (int64_t)g_launch_count.load(); }

int dlb_malloc(dlb_context* ctx, size_t bytes, void** out) {
    return guarded([&] { DLB_REQUIRE(ctx && out, "null argument"); use_device(ctx); DLB_CUDA(cudaMalloc(out, bytes ? bytes : 1)); });
}
int dlb_free(dlb_context* ctx, void* ptr) {
    return guarded([&] { DLB_REQUIRE(ctx, "null context"); use_device(ctx); if (ptr) { DLB_CUDA(cudaStreamSynchronize(ctx->stream)); DLB_CUDA(cudaFree(ptr)); } });
}
int dlb_memcpy_h2d(dlb_context* ctx, void* dst, const void* src, size_t bytes) {
    return guarded([&] { DLB_REQUIRE(ctx, "null context"); use_device(ctx); DLB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream)); });
}
int dlb_memcpy_d2h(dlb_context* ctx, void* dst, const void* src, size_t bytes) {
    return guarded([&] {
        DLB_REQUIRE(ctx, "null context"); use_device(ctx);
        DLB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        DLB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
int dlb_memcpy_d2d(dlb_context* ctx, void* dst, const void* src, size_t bytes) {
    return guarded([&] { DLB_REQUIRE(ctx, "null context"); use_device(ctx); DLB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream)); });
}
int dlb_memset(dlb_context* ctx, void* dst, int value, size_t bytes) {
    return guarded([&] { DLB_REQUIRE(ctx, "null context"); use_device(ctx); DLB_CUDA(cudaMemsetAsync(dst, value, bytes, ctx->stream)); });
}
int dlb_host_alloc_pinned(size_t bytes, void** out) { return guarded([&] { DLB_REQUIRE(out, "null argument"); DLB_CUDA(cudaMallocHost(out, bytes ? bytes : 1)); }); }
int dlb_host_free_pinned(void* ptr) { return guarded([&] { if (ptr) DLB_CUDA(cudaFreeHost(ptr)); }); }

#define DLB_DEFINE_OPS(SFX, REAL)                                                                                                  \
    int dlb_conv_forward_##SFX(dlb_context* ctx, const dlb_layer_spec* spec, int batch, const REAL* x, const int32_t* gather,      \
                               const REAL* bias, const REAL* filters, REAL* y, REAL* deriv) {                                     \
        return guarded([&] {                                                                                                       \
            DLB_REQUIRE(ctx && x && bias && filters && y && batch > 0, "Unexpected size of the input tensor");                     \
            check_spec(spec, DLB_LAYER_CONVOLUTION); use_device(ctx); ensure_op_ws(ctx);                                           \
            conv_forward<REAL>(ctx->lc(), batch, make_conv_geom(*spec), spec->activation, x, gather, bias, filters, y, deriv);     \
        });                                                                                                                        \
    }                                                                                                                              \
    int dlb_conv_backward_##SFX(dlb_context* ctx, const dlb_layer_spec* spec, int batch, const REAL* x, const int32_t* gather,     \
                                const REAL* delta, const REAL* deriv, const REAL* filters, REAL* g, REAL* dbias, REAL* dfilters,   \
                                REAL* dx, double scale) {                                                                          \
        return guarded([&] {                                                                                                       \
            DLB_REQUIRE(ctx && x && delta && deriv && filters && g && dbias && dfilters && batch > 0,                              \
                        "Unexpected size of the input tensor of derivatives");                                                     \
            check_spec(spec, DLB_LAYER_CONVOLUTION); use_device(ctx); ensure_op_ws(ctx);                                           \
            conv_backward<REAL>(ctx->lc(), batch, make_conv_geom(*spec), spec->activation, x, gather, delta, deriv, filters, g,    \
                                dbias, dfilters, dx, static_cast<REAL>(scale));                                                    \
        });                                                                                                                        \
    }                                                                                                                              \
    int dlb_pool_forward_##SFX(dlb_context* ctx, const dlb_layer_spec* spec, int batch, const REAL* x, REAL* y, int32_t* map) {    \
        return guarded([&] {                                                                                                       \
            DLB_REQUIRE(ctx && x && y && batch > 0, "Unexpected size of the input tensor");                                        \
            check_spec(spec, DLB_LAYER_PULL); use_device(ctx);                                                                     \
            pool_forward<REAL>(ctx->lc(), batch, *spec, x, y, map);                                                                \
        });                                                                                                                        \
    }                                                                                                                              \
    int dlb_pool_backward_##SFX(dlb_context* ctx, const dlb_layer_spec* spec, int batch, const REAL* delta, const int32_t* map,    \
                                REAL* dx) {                                                                                        \
        return guarded([&] {                                                                                                       \
            DLB_REQUIRE(ctx && delta && dx && batch > 0, "Unexpected size of the input tensor of derivatives");                    \
            check_spec(spec, DLB_LAYER_PULL);                                                                                      \
            DLB_REQUIRE(map || spec->pool_type == DLB_POOL_AVERAGE, "Invalid index mapping");                                      \
            use_device(ctx);                                                                                                       \
            pool_backward<REAL>(ctx->lc(), batch, *spec, delta, map, dx);                                                          \
        });                                                                                                                        \
    }                                                                                                                              \
    int dlb_full_forward_##SFX(dlb_context* ctx, const dlb_layer_spec* spec, int batch, const REAL* x, const int32_t* gather,      \
                               const REAL* bias, const REAL* weights, REAL* y, REAL* deriv) {                                     \
        return guarded([&] {                                                                                                       \
            DLB_REQUIRE(ctx && x && bias && weights && y && batch > 0, "Incompatible matrix-vector dimensionality");               \
            check_spec(spec, DLB_LAYER_FULL); use_device(ctx); ensure_op_ws(ctx);                                                  \
            full_forward<REAL>(ctx->lc(), batch, spec->in_c * spec->in_h * spec->in_w, spec->out_c, spec->activation, x, gather,   \
                               bias, weights, y, deriv);                                                                           \
        });                                                                                                                        \
    }                                                                                                                              \
    int dlb_full_backward_##SFX(dlb_context* ctx, const dlb_layer_spec* spec, int batch, const REAL* x, const int32_t* gather,     \
                                const REAL* delta, const REAL* deriv, const REAL* weights, REAL* g, REAL* dbias, REAL* dweights,   \
                                REAL* dx, double scale) {                                                                          \
        return guarded([&] {                                                                                                       \
            DLB_REQUIRE(ctx && x && delta && deriv && weights && g && dbias && dweights && batch > 0, "Invalid input");            \
            check_spec(spec, DLB_LAYER_FULL); use_device(ctx); ensure_op_ws(ctx);                                                  \
            full_backward<REAL>(ctx->lc(), batch, spec->in_c * spec->in_h * spec->in_w, spec->out_c, spec->activation, x, gather,  \
                                delta, deriv, weights, g, dbias, dweights, dx, static_cast<REAL>(scale));                          \
        });                                                                                                                        \
    }                                                                                                                              \
    int dlb_cost_gradient_##SFX(dlb_context* ctx, int cost_id, int batch, int n, const REAL* out, const REAL* ref,                 \
                                const int32_t* gather, REAL* grad) {                                                               \
        return guarded([&] {                                                                                                       \
            DLB_REQUIRE(ctx && out && ref && grad && batch > 0 && n > 0, "Inconsistent input.");                                   \
            DLB_REQUIRE(cost_id >= DLB_COST_SQUARED_ERROR && cost_id <= DLB_COST_LINEAR, "Unknown cost function");                 \
            use_device(ctx);                                                                                                       \
            cost_gradient<REAL>(ctx->lc(), batch, n, cost_id, out, ref, gather, grad);                                             \
        });                                                                                                                        \
    }                                                                                                                              \
    int dlb_sgd_update_##SFX(dlb_context* ctx, REAL* params, const REAL* grads, int64_t n, int regularised, double l_rate,         \
                             double reg_factor, double grad_scale_factor) {                                                        \
        return guarded([&] { DLB_REQUIRE(ctx, "null context"); op_sgd_update<REAL>(ctx, params, grads, n, regularised,             \
                                                                                 l_rate, reg_factor, grad_scale_factor); });       \
    }

DLB_DEFINE_OPS(f32, float)
DLB_DEFINE_OPS(f64, double)

int dlb_net_create(dlb_context* ctx, const dlb_layer_spec* specs, int n_layers, int real_bytes, dlb_net** out) {
    return guarded([&] {
        DLB_REQUIRE(ctx && out, "null argument");
        DLB_REQUIRE(real_bytes == 4 || real_bytes == 8, "real_bytes must be 4 (float) or 8 (double)");
        int64_t np = 0;
        auto layers = build_layers(specs, n_layers, np);
        if (real_bytes == 4) *out = new NetT<float>(ctx, std::move(layers), np);
        else *out = new NetT<double>(ctx, std::move(layers), np);
    });
}
int dlb_net_destroy(dlb_net* net) { return guarded([&] { delete net; }); }
int64_t dlb_net_param_count(const dlb_net* net) { return net ? net->n_params : -1; }
int dlb_net_layer_count(const dlb_net* net) { return net ? (int)net->layers.size() : -1; }
int64_t dlb_net_layer_in_size(const dlb_net* net, int l) { return (net && l >= 0 && l < (int)net->layers.size()) ? net->layers[(size_t)l].in_n : -1; }
int64_t dlb_net_layer_out_size(const dlb_net* net, int l) { return (net && l >= 0 && l < (int)net->layers.size()) ? net->layers[(size_t)l].out_n : -1; }
int64_t dlb_net_layer_param_offset(const dlb_net* net, int l) { return (net && l >= 0 && l < (int)net->layers.size()) ? net->layers[(size_t)l].param_off : -1; }
int dlb_net_set_params(dlb_net* net, const void* h) { return guarded([&] { DLB_REQUIRE(net && h, "null argument"); net->set_params(h); }); }
int dlb_net_get_params(dlb_net* net, void* h) { return guarded([&] { DLB_REQUIRE(net && h, "null argument"); net->get_params(h); }); }
void* dlb_net_device_params(dlb_net* net) { return net ? net->device_params() : nullptr; }
void* dlb_net_device_grads(dlb_net* net) { return net ? net->device_grads() : nullptr; }
int dlb_net_set_dataset(dlb_net* net, const void* x, const void* y, int64_t n, int on_device) {
    return guarded([&] { DLB_REQUIRE(net, "null net"); net->set_dataset(x, y, n, on_device != 0); });
}
int dlb_net_reset_random_generator(dlb_net* net, unsigned seed) {
    return guarded([&] { DLB_REQUIRE(net, "null net"); net->rng.seed(seed); net->layer_rng.seed(seed); });   // Net.inl:620-624 seeds both
}
int dlb_net_set_keep_rate(dlb_net* net, int layer, double keep_rate) { return guarded([&] { DLB_REQUIRE(net, "null net"); net->set_keep_rate(layer, keep_rate); }); }
int64_t dlb_net_mask_size(const dlb_net* net) { return net ? net->mask_size() : -1; }
int dlb_net_set_dropout_masks(dlb_net* net, const void* masks_host) { return guarded([&] { DLB_REQUIRE(net, "null net"); net->set_dropout_masks(masks_host); }); }
int dlb_net_learn(dlb_net* net, int64_t batch, int64_t epochs, double lr, int cost, double lambda, dlb_epoch_callback cb, void* user) {
    return guarded([&] {
        DLB_REQUIRE(net, "null net");
        DLB_REQUIRE(cost >= DLB_COST_SQUARED_ERROR && cost <= DLB_COST_LINEAR, "Unknown cost function");
        net->learn(batch, epochs, lr, cost, lambda, cb, user);
    });
}
int dlb_net_step_indices(dlb_net* net, const int64_t* idx, int64_t n, double lr, int cost, double reg_factor) {
    return guarded([&] { DLB_REQUIRE(net, "null net"); net->step_indices(idx, n, lr, cost, reg_factor); });
}
int dlb_net_step_host(dlb_net* net, const void* x, const void* y, int64_t n, double lr, int cost, double reg_factor, double* cost_out) {
    return guarded([&] { DLB_REQUIRE(net, "null net"); net->step_host(x, y, n, lr, cost, reg_factor, cost_out, true); });
}
int dlb_net_step_host_async(dlb_net* net, const void* x, const void* y, int64_t n, double lr, int cost, double reg_factor, double* cost_out) {
    return guarded([&] { DLB_REQUIRE(net, "null net"); net->step_host(x, y, n, lr, cost, reg_factor, cost_out, false); });
}
int dlb_net_forward_backward_host(dlb_net* net, const void* x, const void* y, int64_t n, int cost, double scale) {
    return guarded([&] { DLB_REQUIRE(net, "null net"); net->forward_backward_host(x, y, n, cost, scale); });
}
int dlb_net_read(dlb_net* net, int what, int layer, void* dst, int64_t bytes) {
    return guarded([&] { DLB_REQUIRE(net && dst, "null argument"); net->read(what, layer, dst, bytes); });
}
int dlb_net_write_gradient(dlb_net* net, const void* src) { return guarded([&] { DLB_REQUIRE(net && src, "null argument"); net->write_gradient(src); }); }
int dlb_net_eval_host(dlb_net* net, const void* x, int64_t n, int32_t* classes, void* outputs) {
    return guarded([&] { DLB_REQUIRE(net, "null net"); net->eval_host(x, n, classes, outputs); });
}

int dlb_net_describe_plan(dlb_net* net, char* buf, int64_t cap) {
    return guarded([&] { DLB_REQUIRE(net && buf && cap > 0, "null argument"); std::snprintf(buf, (size_t)cap, "%s", net->describe_plan().c_str()); });
}
int dlb_net_enable_tensor_cores(dlb_net* net, int mode) {
    return guarded([&] {
        DLB_REQUIRE(net, "null net");
        DLB_REQUIRE(mode >= 0 && mode <= 2, "tensor-core mode must be 0 (off), 1 (auto) or 2 (always)");
        net->enable_tensor_cores(mode);
    });
}

// C[M][N] = sum_k A(m, k) * B(n, k) in fp32.  a_transposed / b_transposed: the operand is stored [K][M] / [K][N]
// (leading dimension lda / ldb) instead of [M][K] / [N][K].  use_tensor_cores: tcgen05 3xTF32 (tc_gemm.cuh) or FMA.
int dlb_matmul_f32(dlb_context* ctx, int M, int N, int K, const float* A, int64_t lda, int a_transposed, const float* B, int64_t ldb,
                   int b_transposed, float* C, int use_tensor_cores) {
    return guarded([&] {
        DLB_REQUIRE(ctx && A && B && C && M > 0 && N > 0 && K > 0, "Incompatible matrix-vector dimensionality");
        use_device(ctx); ensure_op_ws(ctx);
        TagScope tag_("matmul");
        if (use_tensor_cores) {
#ifdef DLB_CUDA_EMU
            throw std::runtime_error("tensor cores are not available");
#else
            const TcOperand oa{A, lda, a_transposed ? kOuterContig : kRedContig, nullptr, nullptr, -1, {}};
            const TcOperand ob{B, ldb, b_transposed ? kOuterContig : kRedContig, nullptr, nullptr, -1, {}};
            const EpiStore<float> ep{C, (size_t)N};
            if (!a_transposed && !b_transposed) tc_gemm<kRedContig, kRedContig>(ctx->stream, M, N, K, oa, ob, ep, ctx->ws);
            else if (!a_transposed && b_transposed) tc_gemm<kRedContig, kOuterContig>(ctx->stream, M, N, K, oa, ob, ep, ctx->ws);
            else if (a_transposed && !b_transposed) tc_gemm<kOuterContig, kRedContig>(ctx->stream, M, N, K, oa, ob, ep, ctx->ws);
            else tc_gemm<kOuterContig, kOuterContig>(ctx->stream, M, N, K, oa, ob, ep, ctx->ws);
#endif
        } else {
            auto run = [&](auto al, auto bl) { gemm<float>(ctx->stream, M, N, K, al, bl, EpiStore<float>{C, (size_t)N}, ctx->ws); };
            if (!a_transposed && !b_transposed) run(DenseRowLoader<float>{A, (size_t)lda, nullptr}, DenseRowLoader<float>{B, (size_t)ldb, nullptr});
            else if (!a_transposed && b_transposed) run(DenseRowLoader<float>{A, (size_t)lda, nullptr}, DenseColLoader<float>{B, (size_t)ldb, nullptr});
            else if (a_transposed && !b_transposed) run(DenseColLoader<float>{A, (size_t)lda, nullptr}, DenseRowLoader<float>{B, (size_t)ldb, nullptr});
            else run(DenseColLoader<float>{A, (size_t)lda, nullptr}, DenseColLoader<float>{B, (size_t)ldb, nullptr});
        }
    });
}

// The same contraction for Real = double: the DFMA micro-kernels of gemm.cuh (tcgen05 has no fp64 kind, and mma.sync.m8n8k4.f64
// measures the same peak as the DFMA pipe on this part: tools/peaks.cu).
int dlb_matmul_f64(dlb_context* ctx, int M, int N, int K, const double* A, int64_t lda, int a_transposed, const double* B, int64_t ldb,
                   int b_transposed, double* C) {
    return guarded([&] {
        DLB_REQUIRE(ctx && A && B && C && M > 0 && N > 0 && K > 0, "Incompatible matrix-vector dimensionality");
        use_device(ctx); ensure_op_ws(ctx);
        TagScope tag_("matmul");
        auto run = [&](auto al, auto bl) { gemm<double>(ctx->stream, M, N, K, al, bl, EpiStore<double>{C, (size_t)N}, ctx->ws); };
        if (!a_transposed && !b_transposed) run(DenseRowLoader<double>{A, (size_t)lda, nullptr}, DenseRowLoader<double>{B, (size_t)ldb, nullptr});
        else if (!a_transposed && b_transposed) run(DenseRowLoader<double>{A, (size_t)lda, nullptr}, DenseColLoader<double>{B, (size_t)ldb, nullptr});
        else if (a_transposed && !b_transposed) run(DenseColLoader<double>{A, (size_t)lda, nullptr}, DenseRowLoader<double>{B, (size_t)ldb, nullptr});
        else run(DenseColLoader<double>{A, (size_t)lda, nullptr}, DenseColLoader<double>{B, (size_t)ldb, nullptr});
    });
}

#ifdef DLB_TC_PROFILE
// Development builds only (-DDLB_TC_PROFILE): reads and clears the phase clocks of tc_gemm_kernel (not part of the ABI).
extern "C" __attribute__((visibility("default"))) int dlb_debug_tc_profile(unsigned long long* out16) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, tc::g_tc_prof, 16 * sizeof(unsigned long long));
    unsigned long long zero[16] = {};
    cudaMemcpyToSymbol(tc::g_tc_prof, zero, sizeof(zero));
    return 0;
}
#endif

int dlb_net_enable_step_kernel(dlb_net* net, int mode) {
    return guarded([&] {
        DLB_REQUIRE(net, "null net");
        DLB_REQUIRE(mode == 0 || mode == 1, "step-kernel mode must be 0 (off) or 1 (on)");
        net->enable_step_kernel(mode);
    });
}

int dlb_net_enable_graphs(dlb_net* net, int on) { return guarded([&] { DLB_REQUIRE(net, "null net"); net->enable_graphs(on != 0); }); }

int dlb_net_train_steps(dlb_net* net, int64_t batch, int64_t n_steps, double lr, int cost, double lambda, int64_t* items_out) {
    return guarded([&] {
        DLB_REQUIRE(net, "null net");
        DLB_REQUIRE(cost >= DLB_COST_SQUARED_ERROR && cost <= DLB_COST_LINEAR, "Unknown cost function");
        const int64_t items = net->train_steps(batch, n_steps, lr, cost, lambda);
        if (items_out) *items_out = items;
    });
}

int dlb_rank_slice(int64_t count, int rank, int world, int64_t* first, int64_t* mine) {
    return guarded([&] {
        DLB_REQUIRE(first && mine && count >= 0 && world >= 1 && rank >= 0 && rank < world, "Invalid rank / world size");
        const int64_t per = (count + world - 1) / world;
        *first = std::min<int64_t>(count, (int64_t)rank * per);
        *mine = std::min<int64_t>(count, *first + per) - *first;
    });
}

#ifndef DLB_CUDA_EMU
int dlb_event_create(void** out) { return guarded([&] { DLB_REQUIRE(out, "null argument"); cudaEvent_t e; DLB_CUDA(cudaEventCreate(&e)); *out = e; }); }
int dlb_event_destroy(void* ev) { return guarded([&] { if (ev) DLB_CUDA(cudaEventDestroy((cudaEvent_t)ev)); }); }
int dlb_event_record(dlb_context* ctx, void* ev) { return guarded([&] { DLB_REQUIRE(ctx && ev, "null argument"); DLB_CUDA(cudaEventRecord((cudaEvent_t)ev, ctx->stream)); }); }
int dlb_event_elapsed_ms(void* start, void* stop, double* ms) {
    return guarded([&] {
        DLB_REQUIRE(start && stop && ms, "null argument");
        DLB_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
        float f = 0; DLB_CUDA(cudaEventElapsedTime(&f, (cudaEvent_t)start, (cudaEvent_t)stop)); *ms = f;
    });
}
int dlb_timing_enable(int on) { return guarded([&] { g_timing_on = on != 0; }); }
// Writes "key\tcount\ttotal_ms\n" lines (NUL-terminated, truncated to cap) and clears the records.
int dlb_timing_report(char* buf, int64_t cap) {
    return guarded([&] {
        DLB_REQUIRE(buf && cap > 0, "null argument");
        DLB_CUDA(cudaDeviceSynchronize());
        std::vector<std::string> keys; std::vector<long long> counts; std::vector<double> totals;
        for (auto& t : g_timed) {
            float ms = 0; DLB_CUDA(cudaEventElapsedTime(&ms, t.e0, t.e1));
            size_t i = 0; for (; i < keys.size(); ++i) if (keys[i] == t.key) break;
            if (i == keys.size()) { keys.push_back(t.key); counts.push_back(0); totals.push_back(0.0); }
            counts[i]++; totals[i] += ms;
            g_event_pool.push_back(t.e0); g_event_pool.push_back(t.e1);
        }
        g_timed.clear();
        std::string out;
        for (size_t i = 0; i < keys.size(); ++i) out += keys[i] + "\t" + std::to_string(counts[i]) + "\t" + std::to_string(totals[i]) + "\n";
        std::snprintf(buf, (size_t)cap, "%s", out.c_str());
    });
}
#else
int dlb_event_create(void** out) { if (out) *out = (void*)0x1; return 0; }
int dlb_event_destroy(void*) { return 0; }
int dlb_event_record(dlb_context*, void*) { return 0; }
int dlb_event_elapsed_ms(void*, void*, double* ms) { if (ms) *ms = 1.0; return 0; }
int dlb_timing_enable(int) { return 0; }
int dlb_timing_report(char* buf, int64_t cap) { if (buf && cap > 0) buf[0] = 0; return 0; }
#endif

int dlb_net_evaluate_host(dlb_net* net, const void* x, const void* y, int64_t n, int cost, double l2_reg_factor, double* cost_out, double* correct_out) {
    return guarded([&] {
        DLB_REQUIRE(net, "null net");
        DLB_REQUIRE(cost >= DLB_COST_SQUARED_ERROR && cost <= DLB_COST_LINEAR, "Unknown cost function");
        net->evaluate_host(x, y, n, cost, l2_reg_factor, cost_out, correct_out);
    });
}

int dlb_nccl_get_unique_id(void* id_out) {
    return guarded([&] {
#ifdef DLB_CUDA_EMU
        (void)id_out; throw std::runtime_error("NCCL is not available");
#else
        DLB_REQUIRE(id_out, "null argument");
        auto& api = NcclApi::get();
        if (!api.ok) throw std::runtime_error(api.why);
        NcclApi::unique_id id;
        api.check(api.GetUniqueId(&id), "ncclGetUniqueId");
        std::memcpy(id_out, &id, sizeof id);
#endif
    });
}

int dlb_net_set_distributed(dlb_net* net, int rank, int world, const void* unique_id) {
    return guarded([&] {
        DLB_REQUIRE(net && world >= 1 && rank >= 0 && rank < world, "Invalid rank / world size");
#ifdef DLB_CUDA_EMU
        (void)unique_id; DLB_REQUIRE(world == 1, "NCCL is not available");
#else
        if (world > 1) {
            DLB_REQUIRE(unique_id, "null ncclUniqueId");
            auto& api = NcclApi::get();
            if (!api.ok) throw std::runtime_error(api.why);
            use_device(net->ctx);
            NcclApi::unique_id id;
            std::memcpy(&id, unique_id, sizeof id);
            if (net->comm) { api.CommDestroy(net->comm); net->comm = nullptr; }
            api.check(api.CommInitRank(&net->comm, world, id, rank), "ncclCommInitRank");
        }
#endif
        net->rank = rank; net->world = world;
    });
}

int dlb_net_p2p_export(dlb_net* net, void* handle_out_64_bytes) {
    return guarded([&] { DLB_REQUIRE(net && handle_out_64_bytes, "null argument"); net->p2p_export(handle_out_64_bytes); });
}
int dlb_net_set_distributed_p2p(dlb_net* net, int rank, int world, const void* handles) {
    return guarded([&] { DLB_REQUIRE(net, "null net"); net->p2p_connect(rank, world, handles); });
}

int dlb_shuffle_sequence(unsigned seed, int64_t n, int64_t epochs, int64_t* out) {
    return guarded([&] {
        DLB_REQUIRE(out && n >= 0 && epochs >= 0, "Invalid input.");
        std::mt19937 gen(seed);
        std::vector<std::size_t> idx((size_t)n);
        std::iota(idx.begin(), idx.end(), 0);
        for (int64_t e = 0; e < epochs; ++e) {
            std::ranges::shuffle(idx, gen);
            for (int64_t i = 0; i < n; ++i) out[e * n + i] = (int64_t)idx[(size_t)i];
        }
    });
}

int dlb_dropout_masks(unsigned seed, const int64_t* in_sizes, const double* keep_rates, int n_layers, int64_t n_batches, int real_bytes, void* out) {
    return guarded([&] {
        DLB_REQUIRE(in_sizes && keep_rates && out && n_layers > 0 && n_batches >= 0, "Invalid input.");
        DLB_REQUIRE(real_bytes == 4 || real_bytes == 8, "real_bytes must be 4 or 8");
        std::mt19937 gen(seed);
        auto run = [&](auto* dst) {
            using R = std::remove_pointer_t<decltype(dst)>;
            std::vector<R> vals; std::vector<int> idx;
            for (int64_t b = 0; b < n_batches; ++b)
                for (int l = 0; l < n_layers; ++l) {
                    const R keep = static_cast<R>(keep_rates[l]);
                    if (!(keep < R(1))) continue;
                    draw_keep_mask<R>(gen, (size_t)in_sizes[l], keep, dst, vals, idx);
                    dst += in_sizes[l];
                }
        };
        if (real_bytes == 4) run(static_cast<float*>(out)); else run(static_cast<double*>(out));
    });
}

}  // extern "C"

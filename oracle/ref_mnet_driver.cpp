// oracle/ref_mnet_driver.cpp — TEST INFRASTRUCTURE, never linked into or called from the product.
//
// A thin extern "C" driver around the UNMODIFIED reference recurrent stack (/root/reference/DeepLearning:
// MNet<CpuDC> with RMLayer<CpuDC>, NeuralNet/MNet.{h,inl}, NeuralNet/RMLayer.{h,inl}); compiled together with the
// reference's own sources (where they lie, never copied) by oracle/Makefile into oracle/_ref/libdlrefm_{f32,f64}.so.
// No arithmetic of its own: it builds the net, moves flat arrays in and out, and calls MNet::act / calc_gradient_sum /
// learn.  RMLayer keeps its parameters private and offers no accessor; this translation unit reaches them through the
// explicit-instantiation idiom (an explicit template instantiation may name private members) so that tests can start both
// implementations from the same weights - no reference file is edited.
//
// Flat parameter layout (the one include/dlb200.h documents for dlb_mnet): per layer  bias [H] | in-weights [H][I] |
// recurrent weights [H][H], row-major - the order of RMLayer's gradient containers (BIAS_GRAD_ID, IN_W_GRAD_ID,
// REC_W_GRAD_ID, RMLayer.h:120-122).
#include <NeuralNet/RMLayer.h>
#include <NeuralNet/MNet.h>
#include <Math/CostFunction.h>

#include <cstdint>
#include <cstring>
#include <string>

using namespace DeepLearning;

namespace {
thread_local std::string g_err;

// private-member access without touching the reference: Rob<Tag, &Class::member> is explicitly instantiated below
template <class Tag, typename Tag::type M> struct Rob { friend typename Tag::type get(Tag) { return M; } };
struct BiasesTag { using type = CpuDC::vector_t RMLayer<CpuDC>::*; friend type get(BiasesTag); };
struct WeightsTag { using type = std::vector<CpuDC::matrix_t> RMLayer<CpuDC>::*; friend type get(WeightsTag); };
struct LayersTag { using type = std::vector<MLayerHandle<CpuDC>> MNet<CpuDC>::*; friend type get(LayersTag); };
template struct Rob<BiasesTag, &RMLayer<CpuDC>::_biases>;
template struct Rob<WeightsTag, &RMLayer<CpuDC>::_weights>;
template struct Rob<LayersTag, &MNet<CpuDC>::_layers>;
CpuDC::vector_t& biases_of(RMLayer<CpuDC>& l) { return l.*get(BiasesTag{}); }
std::vector<CpuDC::matrix_t>& weights_of(RMLayer<CpuDC>& l) { return l.*get(WeightsTag{}); }
std::vector<MLayerHandle<CpuDC>>& layers_of(MNet<CpuDC>& n) { return n.*get(LayersTag{}); }

struct RefMNet {
    MNet<CpuDC> net;
    std::vector<RMLayer<CpuDC>*> layers;   // the net's own layer objects
    int depth = 0;
};

template <class F>
int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::runtime_error& e) { g_err = e.what(); return 1; }
    catch (...) { g_err = "unknown exception"; return 2; }
}

// [n][T][sz] flat -> n sequences of T tensors
std::vector<std::vector<Tensor>> to_sequences(const Real* data, int64_t n, int depth, const Index3d& size) {
    std::vector<std::vector<Tensor>> out((std::size_t)n);
    const auto sz = (std::size_t)size.coord_prod();
    for (int64_t i = 0; i < n; ++i)
        for (int t = 0; t < depth; ++t) {
            Tensor x(size, false);
            std::memcpy(x.begin(), data + ((std::size_t)i * depth + t) * sz, sz * sizeof(Real));
            out[(std::size_t)i].push_back(std::move(x));
        }
    return out;
}

LazyVector<std::span<const Tensor>> spans_of(const std::vector<std::vector<Tensor>>& seqs) {
    LazyVector<std::span<const Tensor>> out(seqs.size());
    for (std::size_t i = 0; i < seqs.size(); ++i) out[i] = std::span<const Tensor>(seqs[i].data(), seqs[i].size());
    return out;
}
}  // namespace

extern "C" {

int dlrefm_real_bytes() { return (int)sizeof(Real); }
const char* dlrefm_last_error() { return g_err.c_str(); }

// n_layers RMLayers: layer l maps dims[l] -> dims[l + 1] per time step; activation / clip threshold per layer.
void* dlrefm_create(int depth, int n_layers, const int* dims, const int* activations, const double* clip_thresholds, unsigned seed) {
    RefMNet* h = nullptr;
    const int rc = guarded([&] {
        MNet<CpuDC>::reset_random_generator(seed);
        h = new RefMNet();
        h->depth = depth;
        for (int l = 0; l < n_layers; ++l)
            h->net.append_layer<RMLayer>(depth, dims[l], Index3d{ 1, 1, dims[l + 1] }, FillRandomNormal, static_cast<ActivationFunctionId>(activations[l]),
                                         static_cast<Real>(clip_thresholds[l]));
        for (int l = 0; l < n_layers; ++l) {
            auto* p = dynamic_cast<RMLayer<CpuDC>*>(&layers_of(h->net)[(std::size_t)l].layer());
            if (!p) throw std::runtime_error("not an RMLayer");
            h->layers.push_back(p);
        }
    });
    if (rc != 0) { delete h; return nullptr; }
    return h;
}
void dlrefm_destroy(void* h) { delete static_cast<RefMNet*>(h); }

int64_t dlrefm_param_count(void* hv) {
    int64_t n = 0;
    for (auto* L : static_cast<RefMNet*>(hv)->layers) n += (int64_t)(biases_of(*L).size() + weights_of(*L)[0].size() + weights_of(*L)[1].size());
    return n;
}
int dlrefm_get_params(void* hv, Real* out) {
    return guarded([&] {
        for (auto* L : static_cast<RefMNet*>(hv)->layers) {
            std::memcpy(out, biases_of(*L).begin(), biases_of(*L).size() * sizeof(Real)); out += biases_of(*L).size();
            for (int w = 0; w < 2; ++w) { std::memcpy(out, weights_of(*L)[w].begin(), weights_of(*L)[w].size() * sizeof(Real)); out += weights_of(*L)[w].size(); }
        }
    });
}
int dlrefm_set_params(void* hv, const Real* in) {
    return guarded([&] {
        for (auto* L : static_cast<RefMNet*>(hv)->layers) {
            std::memcpy(biases_of(*L).begin(), in, biases_of(*L).size() * sizeof(Real)); in += biases_of(*L).size();
            for (int w = 0; w < 2; ++w) { std::memcpy(weights_of(*L)[w].begin(), in, weights_of(*L)[w].size() * sizeof(Real)); in += weights_of(*L)[w].size(); }
        }
    });
}

// MNet::act on n sequences: x [n][T][in] -> out [n][T][out]
int dlrefm_act(void* hv, const Real* x, int64_t n, Real* out) {
    return guarded([&] {
        auto* h = static_cast<RefMNet*>(hv);
        const auto in3 = h->net.in_size().xyz, out3 = h->net.out_size().xyz;
        const auto seqs = to_sequences(x, n, h->depth, in3);
        const auto osz = (std::size_t)out3.coord_prod();
        for (int64_t i = 0; i < n; ++i) {
            LazyVector<Tensor> input((std::size_t)h->depth);
            for (int t = 0; t < h->depth; ++t) input[(std::size_t)t] = seqs[(std::size_t)i][(std::size_t)t];
            const auto y = h->net.act(input);
            for (int t = 0; t < h->depth; ++t) std::memcpy(out + ((std::size_t)i * h->depth + t) * osz, y[(std::size_t)t].begin(), osz * sizeof(Real));
        }
    });
}

// MNet::calc_gradient_sum over the batch (MNet.inl:106-121): grads in the flat parameter layout
int dlrefm_gradient_sum(void* hv, const Real* x, const Real* y, int64_t n, int cost_id, Real* grads) {
    return guarded([&] {
        auto* h = static_cast<RefMNet*>(hv);
        const auto xs = to_sequences(x, n, h->depth, h->net.in_size().xyz), ys = to_sequences(y, n, h->depth, h->net.out_size().xyz);
        const CostFunction<Tensor> cost(static_cast<CostFunctionId>(cost_id));
        const auto g = h->net.calc_gradient_sum(spans_of(xs), spans_of(ys), cost);
        for (const auto& lg : g)
            for (const auto& t : lg[0].data) { std::memcpy(grads, t.begin(), t.size() * sizeof(Real)); grads += t.size(); }
    });
}

// MNet::learn on one mini-batch (MNet.inl:171-180): gradient sum, then every layer's update with scale -lr / n, gradient-norm
// clipping (RMLayer.inl:236-260) and the regularisation factor
int dlrefm_learn(void* hv, const Real* x, const Real* y, int64_t n, int cost_id, double learning_rate, double reg_factor) {
    return guarded([&] {
        auto* h = static_cast<RefMNet*>(hv);
        const auto xs = to_sequences(x, n, h->depth, h->net.in_size().xyz), ys = to_sequences(y, n, h->depth, h->net.out_size().xyz);
        const CostFunction<Tensor> cost(static_cast<CostFunctionId>(cost_id));
        h->net.learn(spans_of(xs), spans_of(ys), cost, static_cast<Real>(learning_rate), static_cast<Real>(reg_factor));
    });
}

}  // extern "C"

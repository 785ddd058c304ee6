// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, not part of the product.
//
// A thin extern "C" driver around the UNMODIFIED reference CPU implementation
// (/root/reference/DeepLearning, Net<CpuDC> with CLayer / PLayer / NLayer).  It is compiled together with
// the reference's own sources (where they lie, never copied) by oracle/Makefile into
// oracle/_ref/libdlref_f32.so and oracle/_ref/libdlref_f64.so.  Only tests/, __graft_entry__.smoke()
// and bench.py's CPU-baseline / reference arm load those libraries.
//
// The driver contains no arithmetic of its own: every number it returns was produced by the reference's
// code (Net.inl, CLayer.inl, PLayer.inl, NLayer.inl, Tensor.cpp, Matrix.cpp, ...).
//
// Flat parameter / gradient layout (the reference's container order):
//   for each layer in order:  NLayer -> [biases(out), weights(out x in) row-major]   (NLayer.inl:180-193)
//                             CLayer -> [biases(F x OH x OW), filter_0 .. filter_{F-1} each (C x kh x kw)] (CLayer.inl:199-213)
//                             PLayer -> nothing                                      (PLayer.inl:170-186)

#include <NeuralNet/Net.h>
#include <NeuralNet/CLayer.h>
#include <NeuralNet/PLayer.h>
#include <NeuralNet/NLayer.h>
#include <Math/CostFunction.h>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>

#include "../include/dlb200_spec.h"

using namespace DeepLearning;

namespace {
thread_local std::string g_err;

struct RefNet {
    Net<CpuDC> net;
    std::vector<dlb_layer_spec> specs;
};

template <class F>
int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
    catch (...) { g_err = "unknown exception"; return 1; }
}

void copy_out(const BasicCollection& c, Real*& dst) {
    std::memcpy(dst, c.begin(), c.size() * sizeof(Real));
    dst += c.size();
}

void copy_in(const BasicCollection& c_const, const Real*& src) {
    auto& c = const_cast<BasicCollection&>(c_const);  // scaffolding: the objects themselves are not const
    std::memcpy(c.begin(), src, c.size() * sizeof(Real));
    src += c.size();
}

template <class Fn>
void for_each_param_container(const Net<CpuDC>& net, Fn&& fn) {
    for (std::size_t l = 0; l < net.layers_count(); ++l) {
        const ALayer<CpuDC>* layer = &net[l];
        if (auto n = dynamic_cast<const NLayer<CpuDC>*>(layer)) {
            fn(n->biases());
            fn(n->weights());
        } else if (auto c = dynamic_cast<const CLayer<CpuDC>*>(layer)) {
            fn(c->biases());
            for (const auto& f : c->filters()) fn(f);
        }
    }
}

void flatten_gradient(const std::vector<LayerGradient<CpuDC>>& g, Real* dst) {
    for (const auto& lg : g)
        for (const auto& t : lg.data) copy_out(t, dst);
}
}  // namespace

extern "C" {

int dlref_real_bytes() { return static_cast<int>(sizeof(Real)); }

const char* dlref_last_error() { return g_err.c_str(); }

// Seeds Net's and ALayer's thread-local generators of the calling thread (Net.inl:599-603).
void dlref_reset_random_generator(unsigned seed) { Net<CpuDC>::reset_random_generator(seed); }

void* dlref_net_create_dropout(const dlb_layer_spec* specs, int n, unsigned seed, const double* keep_rates);

void* dlref_net_create(const dlb_layer_spec* specs, int n, unsigned seed) { return dlref_net_create_dropout(specs, n, seed, nullptr); }

// keep_rates (nullable): the layers' keep_rate constructor argument (CLayer.h:145-147, PLayer.h:104-105, NLayer.h:130-132);
// ALayer::set_keep_rate is protected, so the constructors are the only way in.
void* dlref_net_create_dropout(const dlb_layer_spec* specs, int n, unsigned seed, const double* keep_rates) {
    RefNet* rn = nullptr;
    const int rc = guarded([&] {
        Net<CpuDC>::reset_random_generator(seed);
        rn = new RefNet();
        for (int i = 0; i < n; ++i) {
            const auto& s = specs[i];
            const Index3d in{ s.in_c, s.in_h, s.in_w };
            const auto act = static_cast<ActivationFunctionId>(s.activation);
            const Real keep = keep_rates ? static_cast<Real>(keep_rates[i]) : Real(1);
            switch (s.type) {
            case DLB_LAYER_CONVOLUTION:
                rn->net.append_layer<CLayer>(in, Index2d{ s.win_h, s.win_w }, static_cast<std::size_t>(s.out_c), act,
                    Index3d{ 0, s.pad_h, s.pad_w }, Index3d{ 1, s.stride_h, s.stride_w }, keep);
                break;
            case DLB_LAYER_PULL:
                rn->net.append_layer<PLayer>(in, Index2d{ s.win_h, s.win_w }, static_cast<PoolTypeId>(s.pool_type), keep);
                break;
            case DLB_LAYER_FULL:
                rn->net.append_layer<NLayer>(static_cast<std::size_t>(in.coord_prod()), static_cast<std::size_t>(s.out_c), act,
                    Real(-1), Real(1), true, keep);
                break;
            default: throw std::runtime_error("dlref_net_create: unknown layer type");
            }
            rn->specs.push_back(s);
        }
    });
    if (rc) { delete rn; return nullptr; }
    return rn;
}

void dlref_net_destroy(void* h) { delete static_cast<RefNet*>(h); }

int64_t dlref_net_param_count(void* h) {
    int64_t total = 0;
    for_each_param_container(static_cast<RefNet*>(h)->net, [&](const BasicCollection& c) { total += static_cast<int64_t>(c.size()); });
    return total;
}

int dlref_net_get_params(void* h, Real* out) {
    return guarded([&] { for_each_param_container(static_cast<RefNet*>(h)->net, [&](const BasicCollection& c) { copy_out(c, out); }); });
}

int dlref_net_set_params(void* h, const Real* in) {
    return guarded([&] { for_each_param_container(static_cast<RefNet*>(h)->net, [&](const BasicCollection& c) { copy_in(c, in); }); });
}

// Output element count of layer l (ALayer::out_size().coord_prod()).
int64_t dlref_net_layer_out_size(void* h, int l) { return static_cast<RefNet*>(h)->net[l].out_size().coord_prod(); }
int64_t dlref_net_layer_in_size(void* h, int l) { return static_cast<RefNet*>(h)->net[l].in_size().coord_prod(); }

// One sample forward + cost derivative + backward, layer by layer through the public ALayer API,
// exactly the sequence of Net.inl:117-132 (act_bpg) and Net.inl:263-276 (learn's inner loop), with
// gradient_scale_factor = `scale` applied on top of the incoming content of `grads`.
//   acts      [sum_l out_size(l)]   outputs of every layer (post-activation)
//   derivs    [sum_l out_size(l)]   Trace.Derivatives of every layer (PLayer: left untouched, size 0 in the reference)
//   index_map [sum_l out_size(l)]   Trace.IndexMapping (PLayer MIN/MAX only; other layers untouched)
//   grads     [P]                   in: previous content (used when scale != 0); out: scale*prev + gradient
//   in_grads  [sum_l in_size(l)]    input gradient of every layer l >= 1 (layer 0: untouched, Net.inl:272)
//   cost_grad [out_size(L-1)]       the cost derivative fed to the last layer
int dlref_net_trace(void* h, const Real* input, const Real* target, int cost_id, double scale,
    Real* acts, Real* derivs, int32_t* index_map, Real* grads, Real* in_grads, Real* cost_grad) {
    return guarded([&] {
        const auto& net = static_cast<RefNet*>(h)->net;
        const auto L = net.layers_count();
        std::vector<LayerData<CpuDC>> cache(L);
        const auto in3 = net[0].in_size();
        cache[0].Input = Tensor(in3, false);
        std::memcpy(cache[0].Input.begin(), input, cache[0].Input.size() * sizeof(Real));
        Tensor out;
        for (std::size_t l = 0; l < L; ++l) {
            auto& dst = l + 1 < L ? cache[l + 1].Input : out;
            net[l].act(cache[l].Input, dst, &cache[l].Trace);
            const auto n = static_cast<std::size_t>(net[l].out_size().coord_prod());
            if (acts) { std::memcpy(acts, dst.begin(), n * sizeof(Real)); acts += n; }
            if (derivs) { std::memcpy(derivs, cache[l].Trace.Derivatives.begin(), cache[l].Trace.Derivatives.size() * sizeof(Real)); derivs += n; }
            if (index_map) { std::memcpy(index_map, cache[l].Trace.IndexMapping.data(), cache[l].Trace.IndexMapping.size() * sizeof(int)); index_map += n; }
        }
        Tensor ref(out.size_3d(), false);
        std::memcpy(ref.begin(), target, ref.size() * sizeof(Real));
        const auto cost = CostFunction<Tensor>(static_cast<CostFunctionId>(cost_id));
        cost.deriv_in_place(out, ref);
        if (cost_grad) std::memcpy(cost_grad, out.begin(), out.size() * sizeof(Real));

        std::vector<LayerGradient<CpuDC>> lg;
        net.allocate(lg, false);
        if (grads) {
            const Real* src = grads;
            for (auto& g : lg) for (auto& t : g.data) { std::memcpy(t.begin(), src, t.size() * sizeof(Real)); src += t.size(); }
        }
        std::vector<std::size_t> in_off(L, 0);
        for (std::size_t l = 1; l < L; ++l) in_off[l] = in_off[l - 1] + static_cast<std::size_t>(net[l - 1].in_size().coord_prod());
        Tensor deltas = out, next;
        for (long long l = static_cast<long long>(L) - 1; l >= 0; --l) {
            net[l].backpropagate(deltas, cache[l], next, lg[l], l != 0, static_cast<Real>(scale));
            if (l != 0 && in_grads) std::memcpy(in_grads + in_off[l], next.begin(), next.size() * sizeof(Real));
            std::swap(deltas, next);
        }
        if (grads) flatten_gradient(lg, grads);
    });
}

// Net<CpuDC>::learn (Net.inl:193-324) on n items; returns wall seconds spent inside learn() through *seconds.
// shuffle_seed >= 0 re-seeds the calling thread's generators first, so that the epoch permutations are
// std::ranges::shuffle(iota, mt19937(shuffle_seed)) applied successively (Net.inl:158-161,223).
int dlref_net_learn(void* h, const Real* inputs, const Real* refs, int64_t n, int64_t batch, int64_t epochs,
    double lr, int cost_id, double lambda, int parallelism, int64_t shuffle_seed, double* seconds) {
    return guarded([&] {
        auto& net = static_cast<RefNet*>(h)->net;
        const auto in3 = net.in_size();
        const auto out3 = net.out_size();
        const auto in_n = static_cast<std::size_t>(in3.coord_prod());
        const auto out_n = static_cast<std::size_t>(out3.coord_prod());
        std::vector<Tensor> items, labels;
        items.reserve(n); labels.reserve(n);
        for (int64_t i = 0; i < n; ++i) {
            items.emplace_back(in3, false);
            std::memcpy(items.back().begin(), inputs + i * in_n, in_n * sizeof(Real));
            labels.emplace_back(out3, false);
            std::memcpy(labels.back().begin(), refs + i * out_n, out_n * sizeof(Real));
        }
        if (shuffle_seed >= 0) Net<CpuDC>::reset_random_generator(static_cast<unsigned>(shuffle_seed));
        const auto t0 = std::chrono::steady_clock::now();
        net.learn(items, labels, static_cast<std::size_t>(batch), static_cast<std::size_t>(epochs), static_cast<Real>(lr),
            static_cast<CostFunctionId>(cost_id), static_cast<Real>(lambda), [](auto, auto) {}, parallelism);
        const auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    });
}

// The keep masks Net<CpuDC>::learn draws after reset_random_generator(seed) (which seeds ALayer's generator with the same
// seed, Net.inl:620-624): per mini-batch, per layer in order, SetUpDropoutMask (ALayer.inl:26-35) = the reference's own
// Vector::fill_with_random_selection_map(size * keep_rate) on that generator.  sizes[l] = input size of layer l,
// keep[l] >= 1 = the layer draws nothing.  out: n_batches blocks of sum(sizes[l] for keep[l] < 1) values.
int dlref_dropout_masks(unsigned seed, const int64_t* sizes, const double* keep, int n_layers, int64_t n_batches, Real* out) {
    return guarded([&] {
        std::mt19937 g(seed);
        std::vector<int> aux;
        for (int64_t b = 0; b < n_batches; ++b)
            for (int l = 0; l < n_layers; ++l) {
                const Real k = static_cast<Real>(keep[l]);
                if (!(k < Real(1))) continue;
                const auto n = static_cast<std::size_t>(sizes[l]);
                const auto selected = static_cast<std::size_t>(n * k);     // ALayer.inl:31
                Vector v(n);
                v.fill_with_random_selection_map(selected, aux, &g);
                std::memcpy(out, v.begin(), n * sizeof(Real));
                out += n;
            }
    });
}

// Inference (Net.inl:134-155) + max_element_id (Net.inl:432-433): predicted class per item; optional outputs.
int dlref_net_eval(void* h, const Real* inputs, int64_t n, int32_t* classes, Real* outputs) {
    return guarded([&] {
        const auto& net = static_cast<RefNet*>(h)->net;
        const auto in3 = net.in_size();
        const auto in_n = static_cast<std::size_t>(in3.coord_prod());
        const auto out_n = static_cast<std::size_t>(net.out_size().coord_prod());
        Tensor item(in3, false);
        for (int64_t i = 0; i < n; ++i) {
            std::memcpy(item.begin(), inputs + i * in_n, in_n * sizeof(Real));
            const auto out = net.act(item);
            if (classes) classes[i] = static_cast<int32_t>(out.max_element_id());
            if (outputs) std::memcpy(outputs + i * out_n, out.begin(), out_n * sizeof(Real));
        }
    });
}

// Cost value of one output/reference pair (CostFunction.inl -> CostFunctionHelper::evaluate_cost).
double dlref_cost(const Real* output, const Real* reference, int64_t n, int cost_id) {
    Tensor o(1, 1, static_cast<std::size_t>(n), false), r(1, 1, static_cast<std::size_t>(n), false);
    std::memcpy(o.begin(), output, n * sizeof(Real));
    std::memcpy(r.begin(), reference, n * sizeof(Real));
    return static_cast<double>(CostFunction<Tensor>(static_cast<CostFunctionId>(cost_id))(o, r));
}

// Synthetic data exactly as SURVEY.md §8(d): Tensor(size, 0, 1, &mt19937(seed)) per item, drawn sequentially
// from ONE generator (BasicCollection.cpp uniform_random_fill -> std::uniform_real_distribution<Real>).
int dlref_uniform_items(Real* dst, int64_t n, int c, int hgt, int w, double lo, double hi, unsigned seed) {
    return guarded([&] {
        std::mt19937 g(seed);
        const Index3d sz{ c, hgt, w };
        const auto m = static_cast<std::size_t>(sz.coord_prod());
        for (int64_t i = 0; i < n; ++i) {
            Tensor t(sz, static_cast<Real>(lo), static_cast<Real>(hi), &g);
            std::memcpy(dst + i * m, t.begin(), m * sizeof(Real));
        }
    });
}

// The epoch permutations Net<CpuDC>::learn draws after reset_random_generator(seed): get_indices (Net.inl:164-171)
// followed by one std::ranges::shuffle per epoch on the same mt19937 (Net.inl:158-161, 223).  libstdc++-specific,
// which is why the C oracle and the GPU path take the permutations from here / from the same std:: calls.
int dlref_shuffle_sequence(unsigned seed, int64_t n, int64_t epochs, int64_t* out) {
    return guarded([&] {
        std::mt19937 g(seed);
        std::vector<std::size_t> idx(static_cast<std::size_t>(n));
        std::iota(idx.begin(), idx.end(), 0);
        for (int64_t e = 0; e < epochs; ++e) {
            std::ranges::shuffle(idx, g);
            for (int64_t i = 0; i < n; ++i) out[e * n + i] = static_cast<int64_t>(idx[i]);
        }
    });
}

}  // extern "C"

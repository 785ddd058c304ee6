// tests/cpp/ref_adapter_driver.cpp — the reference's own Net template driving libdlb200 through the binding of
// deeplearning_b200/host/gpu_dc.hpp (TEST INFRASTRUCTURE; built by tests/cpp/Makefile against /root/reference, with the
// shim headers of oracle/shim/; the resulting tests/cpp/_build/libref_adapter_{f32,f64}.so travels to the GPU box).
//
// rad_run: builds TWO identical reference nets (Net<CpuDC>, same seed -> same random weights from the reference's own
// initialisers), trains one with the reference's own Net<CpuDC>::learn on the host and the other through the GpuDC binding
// (data set in GpuDC tensors on the device, DeepLearning::learn -> dlb_net_learn), and hands both parameter vectors and both
// nets' predicted classes (computed by the reference's own act on the two trained nets) back to the test.
#include <gpu_dc.hpp>

#include <cstdint>
#include <cstdlib>
#include <string>
#include <csignal>
#include <execinfo.h>
#include <unistd.h>

using namespace DeepLearning;

namespace {
thread_local std::string g_err;

// RAD_DEBUG_SEGV=1: print a native backtrace when a crash happens inside the driver (development aid)
void segv_backtrace(int sig) {
    void* frames[64];
    const int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}
void maybe_install_segv_handler() {
    if (std::getenv("RAD_DEBUG_SEGV")) { std::signal(SIGSEGV, segv_backtrace); std::signal(SIGABRT, segv_backtrace); }
}

void build(Net<CpuDC>& net, const dlb_layer_spec* specs, int n) {
    for (int i = 0; i < n; ++i) {
        const auto& s = specs[i];
        const Index3d in{ s.in_c, s.in_h, s.in_w };
        const auto act = static_cast<ActivationFunctionId>(s.activation);
        switch (s.type) {
        case DLB_LAYER_CONVOLUTION:
            net.append_layer<CLayer>(in, Index2d{ s.win_h, s.win_w }, static_cast<std::size_t>(s.out_c), act, Index3d{ 0, s.pad_h, s.pad_w },
                                     Index3d{ 1, s.stride_h, s.stride_w });
            break;
        case DLB_LAYER_PULL: net.append_layer<PLayer>(in, Index2d{ s.win_h, s.win_w }, static_cast<PoolTypeId>(s.pool_type)); break;
        case DLB_LAYER_FULL:
            net.append_layer<NLayer>(static_cast<std::size_t>(in.coord_prod()), static_cast<std::size_t>(s.out_c), act, Real(-1), Real(1), true);
            break;
        default: throw std::runtime_error("unknown layer type");
        }
    }
}

std::vector<Tensor> to_tensors(const Real* data, int64_t n, const Index3d& size) {
    std::vector<Tensor> out;
    const auto sz = static_cast<std::size_t>(size.coord_prod());
    for (int64_t i = 0; i < n; ++i) { Tensor t(size, false); std::memcpy(t.begin(), data + i * sz, sz * sizeof(Real)); out.push_back(std::move(t)); }
    return out;
}
}  // namespace

extern "C" {

int rad_real_bytes() { return static_cast<int>(sizeof(Real)); }
const char* rad_last_error() { return g_err.c_str(); }

// Round trip of the binding's converters: specs recovered from the reference layers must equal the specs the net was built from, and
// flatten -> scatter -> flatten must be the identity.  Returns 0 when everything matches.
int rad_check_converters(const dlb_layer_spec* specs, int n_layers, unsigned seed) {
    try {
        Net<CpuDC>::reset_random_generator(seed);
        Net<CpuDC> net;
        build(net, specs, n_layers);
        for (int l = 0; l < n_layers; ++l) {
            const dlb_layer_spec got = dlb200_binding::to_spec(net[l]);
            const dlb_layer_spec& want = specs[l];
            // an NLayer only knows its flat input dimension
            const bool same_in = want.type == DLB_LAYER_FULL ? got.in_c * got.in_h * got.in_w == want.in_c * want.in_h * want.in_w
                                                            : got.in_c == want.in_c && got.in_h == want.in_h && got.in_w == want.in_w;
            const bool same = got.type == want.type && same_in &&
                (want.type == DLB_LAYER_PULL || got.activation == want.activation) && got.out_c == (want.type == DLB_LAYER_PULL ? 0 : want.out_c) &&
                (want.type == DLB_LAYER_FULL || (got.win_h == want.win_h && got.win_w == want.win_w)) &&
                (want.type != DLB_LAYER_CONVOLUTION || (got.pad_h == want.pad_h && got.pad_w == want.pad_w && got.stride_h == want.stride_h && got.stride_w == want.stride_w)) &&
                (want.type != DLB_LAYER_PULL || got.pool_type == want.pool_type);
            if (!same) { g_err = "to_spec mismatch at layer " + std::to_string(l); return 1; }
        }
        auto flat = dlb200_binding::flatten_params(net);
        for (auto& v : flat) v = v * Real(0.5) + Real(0.25);
        dlb200_binding::scatter_params(net, flat);
        if (dlb200_binding::flatten_params(net) != flat) { g_err = "scatter_params / flatten_params are not inverse"; return 1; }
        return 0;
    } catch (const std::runtime_error& e) { g_err = e.what(); return 2; } catch (...) { g_err = "unknown exception"; return 3; }
}

// Net::to_script of the reference net built from `specs` (Net.inl:521-530)
int rad_net_script(const dlb_layer_spec* specs, int n_layers, char* buf, int64_t cap) {
    try {
        Net<CpuDC> net;
        build(net, specs, n_layers);
        const std::string s = net.to_script();
        if ((int64_t)s.size() + 1 > cap) { g_err = "buffer too small"; return 1; }
        std::memcpy(buf, s.c_str(), s.size() + 1);
        return 0;
    } catch (const std::runtime_error& e) { g_err = e.what(); return 2; } catch (...) { g_err = "unknown exception"; return 3; }
}

// The reference's own script reader on `script` (Net::try_load_from_script, Net.inl:56-93), written back by Net::to_script
int rad_script_roundtrip(const char* script, char* buf, int64_t cap) {
    try {
        const Net<CpuDC> net{std::string(script)};
        const std::string s = net.to_script();
        if ((int64_t)s.size() + 1 > cap) { g_err = "buffer too small"; return 1; }
        std::memcpy(buf, s.c_str(), s.size() + 1);
        return 0;
    } catch (const std::runtime_error& e) { g_err = e.what(); return 2; } catch (...) { g_err = "unknown exception"; return 3; }
}

// The binding's containers against the reference's own: the gradient SUM of a batch in LayerGradient containers (vs the sum of
// Net::calc_gradient_and_value over the items), and the per-layer LayerData traces of the first item (vs the reference layers' own act
// with trace data).  max_rel_err: largest |got - want| / (|want| + scale) seen; index maps must be equal.  Returns 0 when the shapes agree.
int rad_check_containers(const dlb_layer_spec* specs, int n_layers, unsigned seed, const Real* x, const Real* y, int64_t n_items, int cost_id,
                         double* max_rel_err_gradient, double* max_rel_err_traces, int64_t* index_map_mismatches) {
    try {
        Net<CpuDC>::reset_random_generator(seed);
        Net<CpuDC> net;
        build(net, specs, n_layers);
        const Index3d in{ specs[0].in_c, specs[0].in_h, specs[0].in_w };
        const Index3d out = net[net.layers_count() - 1].out_size();
        const auto items = to_tensors(x, n_items, in), refs = to_tensors(y, n_items, out);
        const auto cost = static_cast<CostFunctionId>(cost_id);
        auto rel = [](const BasicCollection& got, const BasicCollection& want, double& worst) {
            if (got.size() != want.size()) throw std::runtime_error("container sizes differ");
            double scale = 0; for (std::size_t i = 0; i < want.size(); ++i) scale = std::max(scale, std::abs((double)want.begin()[i]));
            for (std::size_t i = 0; i < want.size(); ++i)
                worst = std::max(worst, std::abs((double)got.begin()[i] - (double)want.begin()[i]) / (std::abs((double)want.begin()[i]) + scale + 1e-30));
        };
        // gradients
        std::vector<LayerGradient<CpuDC>> want;
        net.allocate(want, true);
        for (int64_t i = 0; i < n_items; ++i) {
            auto g = std::get<0>(net.calc_gradient_and_value(items[(std::size_t)i], refs[(std::size_t)i], cost));
            for (std::size_t l = 0; l < want.size(); ++l) want[l] += g[l];
        }
        const auto got = calc_gradient_sum(net, items, refs, cost);
        if (got.size() != want.size()) throw std::runtime_error("layer counts differ");
        *max_rel_err_gradient = 0;
        for (std::size_t l = 0; l < want.size(); ++l) {
            if (got[l].data.size() != want[l].data.size()) throw std::runtime_error("LayerGradient::data sizes differ");
            for (std::size_t k = 0; k < want[l].data.size(); ++k) {
                if (got[l].data[k].size_3d() != want[l].data[k].size_3d()) throw std::runtime_error("gradient tensor shapes differ");
                rel(got[l].data[k], want[l].data[k], *max_rel_err_gradient);
            }
        }
        // traces of the first item: the reference's layers, one after the other, with trace data
        Tensor value;
        const auto traces = act_bpg_traces(net, items[0], refs[0], cost, value);
        *max_rel_err_traces = 0; *index_map_mismatches = 0;
        Tensor cur = items[0];
        for (std::size_t l = 0; l < net.layers_count(); ++l) {
            LayerTraceData<CpuDC> tr;
            Tensor nxt;
            net[l].act(cur, nxt, &tr);
            if (traces[l].Input.size_3d() != net[l].in_size()) throw std::runtime_error("LayerData::Input shape differs");
            Tensor cur_shaped = cur; cur_shaped.reshape(net[l].in_size());
            rel(traces[l].Input, cur_shaped, *max_rel_err_traces);
            if (traces[l].Trace.Derivatives.size() != tr.Derivatives.size()) throw std::runtime_error("Trace.Derivatives sizes differ");
            if (tr.Derivatives.size()) rel(traces[l].Trace.Derivatives, tr.Derivatives, *max_rel_err_traces);
            if (traces[l].Trace.IndexMapping.size() != tr.IndexMapping.size()) throw std::runtime_error("Trace.IndexMapping sizes differ");
            for (std::size_t i = 0; i < tr.IndexMapping.size(); ++i) *index_map_mismatches += traces[l].Trace.IndexMapping[i] != tr.IndexMapping[i];
            cur = std::move(nxt);
        }
        rel(value, cur, *max_rel_err_traces);
        return 0;
    } catch (const std::runtime_error& e) { g_err = e.what(); return 2; } catch (...) { g_err = "unknown exception"; return 3; }
}

// SURVEY.md section 8 f4 - the per-sample layer API of the binding (GpuLayer::act / backpropagate / update on GpuDC containers, one batched
// operator launch with batch = 1 each) against the reference's own layers on the host, layer by layer, the way the reference's GPU tests
// compare the two data contexts (DeepLearningCudaTest/NetLayersCudaTest.cpp:77-172): outputs, Trace.Derivatives, Trace.IndexMapping, input
// gradients, layer gradients (with the scaled accumulation: two items, the second with gradient_scale_factor = 0.5), parameters after
// ALayer::update.  max_rel_err: largest |got - want| / (|want| + scale); index maps must be equal.
int rad_check_per_sample(const dlb_layer_spec* specs, int n_layers, unsigned seed, const Real* x, const Real* y, int cost_id, double* max_rel_err,
                         int64_t* index_map_mismatches) {
    try {
        Net<CpuDC>::reset_random_generator(seed);
        Net<CpuDC> net;
        build(net, specs, n_layers);
        const auto L = net.layers_count();
        const Index3d in{ specs[0].in_c, specs[0].in_h, specs[0].in_w };
        const Index3d out = net[L - 1].out_size();
        const auto items = to_tensors(x, 2, in), refs = to_tensors(y, 2, out);
        const CostFunction<Tensor> cost_cpu(static_cast<CostFunctionId>(cost_id));
        *max_rel_err = 0; *index_map_mismatches = 0;
        auto rel = [&](const Real* got, const Real* want, std::size_t n) {
            double scale = 0; for (std::size_t i = 0; i < n; ++i) scale = std::max(scale, std::abs((double)want[i]));
            for (std::size_t i = 0; i < n; ++i) *max_rel_err = std::max(*max_rel_err, std::abs((double)got[i] - (double)want[i]) / (std::abs((double)want[i]) + scale + 1e-30));
        };
        auto rel_t = [&](const GpuDC::tensor_t& got, const BasicCollection& want) {
            if (got.size() != want.size()) throw std::runtime_error("per-sample API: tensor sizes differ");
            const Tensor h = got.to_host(); rel(h.begin(), want.begin(), want.size());
        };
        std::vector<dlb200_binding::GpuLayer> dev;
        for (std::size_t l = 0; l < L; ++l) dev.emplace_back(net[l]);
        std::vector<LayerGradient<CpuDC>> g_cpu(L);
        std::vector<LayerGradient<GpuDC>> g_gpu(L);
        for (std::size_t l = 0; l < L; ++l) net[l].allocate(g_cpu[l], true);
        for (int item = 0; item < 2; ++item) {
            const Real scale = item == 0 ? Real(0) : Real(0.5);
            // forward with traces, both sides
            std::vector<LayerData<CpuDC>> d_cpu(L); std::vector<LayerData<GpuDC>> d_gpu(L);
            Tensor cur = items[(std::size_t)item]; GpuDC::tensor_t cur_g = GpuDC::from_host(cur);
            for (std::size_t l = 0; l < L; ++l) {
                cur.reshape(net[l].in_size());
                d_cpu[l].Input = cur;
                d_gpu[l].Input = cur_g;
                Tensor nxt; net[l].act(cur, nxt, &d_cpu[l].Trace);
                GpuDC::tensor_t nxt_g(net[l].out_size(), false);
                dev[l].act(cur_g, nxt_g, &d_gpu[l].Trace);
                rel_t(nxt_g, nxt);
                if (d_cpu[l].Trace.Derivatives.size()) rel_t(d_gpu[l].Trace.Derivatives, d_cpu[l].Trace.Derivatives);
                if (d_cpu[l].Trace.IndexMapping.size() != d_gpu[l].Trace.IndexMapping.size()) throw std::runtime_error("per-sample API: index map sizes differ");
                const auto map_g = d_gpu[l].Trace.IndexMapping.to_stdvector();
                for (std::size_t i = 0; i < map_g.size(); ++i) *index_map_mismatches += map_g[i] != d_cpu[l].Trace.IndexMapping[i];
                cur = std::move(nxt); cur_g = std::move(nxt_g);
            }
            // cost derivative, then backward through every layer
            cost_cpu.deriv_in_place(cur, refs[(std::size_t)item]);
            const GpuDC::tensor_t ref_g = GpuDC::from_host(refs[(std::size_t)item]);
            dlb200_binding::cost_deriv_in_place(cur_g, ref_g, static_cast<CostFunctionId>(cost_id));
            rel_t(cur_g, cur);
            for (long long l = (long long)L - 1; l >= 0; --l) {
                Tensor in_grad; GpuDC::tensor_t in_grad_g;
                cur.reshape(net[(std::size_t)l].out_size());   // (an NLayer hands its input gradient back flat)
                net[(std::size_t)l].backpropagate(cur, d_cpu[(std::size_t)l], in_grad, g_cpu[(std::size_t)l], l != 0, scale);
                dev[(std::size_t)l].backpropagate(cur_g, d_gpu[(std::size_t)l], in_grad_g, g_gpu[(std::size_t)l], l != 0, scale);
                if (l != 0) { rel_t(in_grad_g, in_grad); cur = std::move(in_grad); cur_g = std::move(in_grad_g); }
            }
        }
        // gradients (after the scaled accumulation of the second item) and the update
        for (std::size_t l = 0; l < L; ++l) {
            std::vector<Real> flat;
            for (const auto& t : g_cpu[l].data) flat.insert(flat.end(), t.begin(), t.end());
            if (flat.empty()) continue;
            if (g_gpu[l].data.size() != 2 || g_gpu[l].data[0].size() + g_gpu[l].data[1].size() != flat.size()) throw std::runtime_error("per-sample API: gradient sizes differ");
            const Tensor gb = g_gpu[l].data[0].to_host(), gw = g_gpu[l].data[1].to_host();
            rel(gb.begin(), flat.data(), gb.size()); rel(gw.begin(), flat.data() + gb.size(), gw.size());
            net[l].update(g_cpu[l], Real(-0.05), Real(0.01));
            dev[l].update(g_gpu[l], Real(-0.05), Real(0.01));
        }
        const auto want = dlb200_binding::flatten_params(net);
        std::vector<Real> got;
        for (std::size_t l = 0; l < L; ++l) { const auto p = dev[l].params_to_host(); got.insert(got.end(), p.begin(), p.end()); }
        if (got.size() != want.size()) throw std::runtime_error("per-sample API: parameter counts differ");
        rel(got.data(), want.data(), want.size());
        return 0;
    } catch (const std::runtime_error& e) { g_err = e.what(); return 2; } catch (...) { g_err = "unknown exception"; return 3; }
}

int rad_run(const dlb_layer_spec* specs, int n_layers, unsigned init_seed, unsigned learn_seed, const Real* x, const Real* y, int64_t n_items,
            int64_t batch, int64_t epochs, double lr, int cost_id, double lambda, int cpu_threads, Real* params_cpu, Real* params_gpu,
            int32_t* classes_cpu, int32_t* classes_gpu, int32_t* epochs_seen_gpu, double* eval_cost_gpu, double* eval_correct_gpu) {
    maybe_install_segv_handler();
    try {
        const Index3d in{ specs[0].in_c, specs[0].in_h, specs[0].in_w };
        Net<CpuDC> net_cpu, net_gpu;
        Net<CpuDC>::reset_random_generator(init_seed); build(net_cpu, specs, n_layers);
        Net<CpuDC>::reset_random_generator(init_seed); build(net_gpu, specs, n_layers);
        const Index3d out = net_cpu[net_cpu.layers_count() - 1].out_size();
        const auto items = to_tensors(x, n_items, in), refs = to_tensors(y, n_items, out);
        const auto cost = static_cast<CostFunctionId>(cost_id);

        // (1) the reference, on the host
        Net<CpuDC>::reset_random_generator(learn_seed);
        net_cpu.learn(items, refs, static_cast<std::size_t>(batch), static_cast<std::size_t>(epochs), static_cast<Real>(lr), cost,
                      static_cast<Real>(lambda), [](const auto, const auto) {}, cpu_threads);

        // (2) the same net through the GpuDC binding: the data set goes to the device as GpuDC tensors (GpuDC::from_host), the model stays
        // in the reference's Net object and comes back trained
        std::vector<GpuDC::tensor_t> d_items, d_refs;
        for (const auto& t : items) d_items.push_back(GpuDC::from_host(t));
        for (const auto& t : refs) d_refs.push_back(GpuDC::from_host(t));
        int32_t seen = 0;
        DeepLearning::learn(net_gpu, d_items, d_refs, static_cast<std::size_t>(batch), static_cast<std::size_t>(epochs), static_cast<Real>(lr), cost,
                            static_cast<Real>(lambda), [&](const std::size_t, const Real) { ++seen; }, learn_seed);
        *epochs_seen_gpu = seen;
        // the device tensors convert back unchanged (GpuDC::to_host)
        const Tensor back = GpuDC::to_host(d_items[0]);
        if (!(back == items[0])) throw std::runtime_error("GpuDC::to_host(GpuDC::from_host(t)) != t");

        const auto pc = dlb200_binding::flatten_params(net_cpu), pg = dlb200_binding::flatten_params(net_gpu);
        std::memcpy(params_cpu, pc.data(), pc.size() * sizeof(Real));
        std::memcpy(params_gpu, pg.data(), pg.size() * sizeof(Real));
        for (int64_t i = 0; i < n_items; ++i) {   // predictions of both trained nets by the reference's own inference
            classes_cpu[i] = static_cast<int32_t>(net_cpu.act(items[i]).max_element_id());
            classes_gpu[i] = static_cast<int32_t>(net_gpu.act(items[i]).max_element_id());
        }
        const auto ev = DeepLearning::evaluate_cost_function_and_correct_answers(net_gpu, items, refs, cost, Real(0));
        *eval_cost_gpu = ev.Cost; *eval_correct_gpu = ev.CorrectAnswers;
        return 0;
    } catch (const std::runtime_error& e) { g_err = e.what(); return 2; } catch (...) { g_err = "unknown exception"; return 3; }
}

int64_t rad_param_count(const dlb_layer_spec* specs, int n_layers) {
    try {
        Net<CpuDC> net; build(net, specs, n_layers);
        return static_cast<int64_t>(dlb200_binding::flatten_params(net).size());
    } catch (const std::runtime_error& e) { g_err = e.what(); return -1; } catch (...) { g_err = "unknown exception"; return -1; }
}

}  // extern "C"

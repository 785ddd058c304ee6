// dlb200.hpp — C++ host-side mirror of the reference's Net / CLayer / PLayer / NLayer interface for the GPU data
// context, header-only, on top of the C ABI (include/dlb200.h).  No CUDA headers are needed to compile against it.
//
// The reference's client code looks like (DeepLearning/NeuralNet/Net.h:281-286, Net.inl:193-324):
//     Net<GpuDC> net;  Index3d sz{1, 28, 28};
//     sz = net.append_layer<CLayer>(sz, Index2d{5, 5}, 20, ActivationFunctionId::RELU);
//     sz = net.append_layer<PLayer>(sz, Index2d{2, 2}, PoolTypeId::MAX);
//     sz = net.append_layer<NLayer>(sz.coord_prod(), 100, ActivationFunctionId::RELU);
//     net.learn(items, labels, 256, epochs, 0.01, CostFunctionId::CROSS_ENTROPY, 0.1, callback);
// and reads the same here with dlb200::Net<Real> (Real = float | double is the reference's compile-time `Real`,
// defs.h:25-29).  Same argument order and meaning, same ids, and failures throw std::runtime_error carrying the
// message (the reference throws std::exception("...") at the same places).
#pragma once

#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dlb200.h"

namespace dlb200 {

enum class LayerTypeId : unsigned { UNKNOWN = 0, FULL = 1, CONVOLUTION = 2, PULL = 3 };                       // LayerTypeId.h:27-32
enum class ActivationFunctionId : unsigned { UNKNOWN = 0, SIGMOID = 1, TANH = 2, RELU = 3, SOFTMAX = 4, LINEAR = 5 };  // ActivationFunctionId.h:26-33
enum class CostFunctionId : unsigned { UNKNOWN = 0, SQUARED_ERROR = 1, CROSS_ENTROPY = 2, LINEAR = 3 };       // CostFunctionId.h:26-32
enum class PoolTypeId : unsigned { UNKNOWN = 0, MAX = 1, MIN = 2, AVERAGE = 3 };                              // PoolOperator.h:28-33

struct Index2d { long long x{}, y{}; };
struct Index3d {   // {layers, rows, columns}, LinAlg3d.h
    long long x{}, y{}, z{};
    long long coord_prod() const { return x * y * z; }
    bool operator==(const Index3d& o) const { return x == o.x && y == o.y && z == o.z; }
};

inline void check(int rc) {
    if (rc != 0) throw std::runtime_error(dlb_last_error());
}

// One device + one stream.  Replaces the reference's thread_local CudaSetup / cudaStreamPerThread (CudaSetup.cpp:29-76).
class Context {
    dlb_context* h_ = nullptr;
public:
    explicit Context(int device = 0, void* stream = nullptr) { check(dlb_context_create(device, stream, &h_)); }
    ~Context() { if (h_) dlb_context_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    dlb_context* handle() const { return h_; }
    void synchronize() const { check(dlb_context_synchronize(h_)); }
};

// Layer tags for append_layer<L>(...), with the constructor arguments of the reference's layers.
struct CLayer {   // CLayer(in_size, filter_window_size, filters_count, func_id, paddings = {0}, strides = {1}) — CLayer.h:145-148
    static dlb_layer_spec make(const Index3d& in, const Index2d& window, std::size_t filters, ActivationFunctionId act,
                               const Index3d& paddings = {0, 0, 0}, const Index3d& strides = {1, 1, 1}) {
        if (paddings.x != 0) throw std::runtime_error("Unexpected channel size");   // CLayer.inl:38-40
        dlb_layer_spec s{};
        s.type = DLB_LAYER_CONVOLUTION; s.activation = (int)act;
        s.in_c = (int)in.x; s.in_h = (int)in.y; s.in_w = (int)in.z;
        s.win_h = (int)window.x; s.win_w = (int)window.y; s.out_c = (int)filters;
        s.pad_h = (int)paddings.y; s.pad_w = (int)paddings.z; s.stride_h = (int)strides.y; s.stride_w = (int)strides.z;
        return s;
    }
};
struct PLayer {   // PLayer(in_size, pool_window_size, pool_operator_id) — PLayer.h:104-105
    static dlb_layer_spec make(const Index3d& in, const Index2d& window, PoolTypeId pool) {
        dlb_layer_spec s{};
        s.type = DLB_LAYER_PULL; s.pool_type = (int)pool;
        s.in_c = (int)in.x; s.in_h = (int)in.y; s.in_w = (int)in.z;
        s.win_h = (int)window.x; s.win_w = (int)window.y; s.stride_h = s.win_h; s.stride_w = s.win_w;
        return s;
    }
};
struct NLayer {   // NLayer(in_dim, out_dim, func_id) — NLayer.h:130-132
    static dlb_layer_spec make(std::size_t in_dim, std::size_t out_dim, ActivationFunctionId act = ActivationFunctionId::SIGMOID) {
        dlb_layer_spec s{};
        s.type = DLB_LAYER_FULL; s.activation = (int)act;
        s.in_c = 1; s.in_h = 1; s.in_w = (int)in_dim; s.out_c = (int)out_dim; s.stride_h = s.stride_w = 1;
        return s;
    }
};

inline long long conv_out(long long in, long long k, long long pad, long long stride) {   // ConvolutionUtils.cpp:29-37
    const long long t = in + 2 * pad - k;
    return t < 0 ? 0 : t / stride + 1;
}

inline Index3d out_size(const dlb_layer_spec& s) {
    if (s.type == DLB_LAYER_FULL) return {1, 1, s.out_c};
    if (s.type == DLB_LAYER_CONVOLUTION)
        return {s.out_c, conv_out(s.in_h, s.win_h, s.pad_h, s.stride_h), conv_out(s.in_w, s.win_w, s.pad_w, s.stride_w)};
    return {s.in_c, conv_out(s.in_h, s.win_h, 0, s.win_h), conv_out(s.in_w, s.win_w, 0, s.win_w)};
}

// Net<GpuDC> with Real fixed at compile time like the reference.
template <class Real>
class Net {
    static_assert(sizeof(Real) == 4 || sizeof(Real) == 8, "Real is float or double");
    Context* ctx_;
    std::vector<dlb_layer_spec> specs_;
    dlb_net* h_ = nullptr;

    dlb_net* handle() {
        if (!h_) check(dlb_net_create(ctx_->handle(), specs_.data(), (int)specs_.size(), (int)sizeof(Real), &h_));
        return h_;
    }
    struct Trampoline { const std::function<void(std::size_t, Real)>* fn; };
    static void on_epoch(int64_t epoch, double reg, void* user) {
        (*static_cast<Trampoline*>(user)->fn)((std::size_t)epoch, (Real)reg);
    }

public:
    explicit Net(Context& ctx) : ctx_(&ctx) {}
    ~Net() { if (h_) dlb_net_destroy(h_); }
    Net(const Net&) = delete;
    Net& operator=(const Net&) = delete;

    // Net::append_layer<L>(args...) — returns the output size of the appended layer (Net.h:281-286).
    template <class L, class... Args>
    Index3d append_layer(Args&&... args) {
        if (h_) throw std::runtime_error("layers cannot be appended after the net has been used");
        specs_.push_back(L::make(std::forward<Args>(args)...));
        return dlb200::out_size(specs_.back());
    }

    std::size_t layers_count() const { return specs_.size(); }
    Index3d in_size() const { return {specs_.front().in_c, specs_.front().in_h, specs_.front().in_w}; }
    Index3d out_size() const { return dlb200::out_size(specs_.back()); }
    std::size_t param_count() { return (std::size_t)dlb_net_param_count(handle()); }
    std::string describe_plan() { char buf[512]; check(dlb_net_describe_plan(handle(), buf, sizeof buf)); return buf; }

    // The keep_rate constructor argument of the reference's layers (CLayer.h:147, PLayer.h:105, NLayer.h:132; ALayer.inl:25-55):
    // one minus the dropout rate of the layer's input.  Masks are drawn per mini-batch from ALayer's generator, which
    // reset_random_generator(seed) seeds together with the permutation generator (Net.inl:620-624).
    void set_keep_rate(std::size_t layer, Real keep_rate) { check(dlb_net_set_keep_rate(handle(), (int)layer, (double)keep_rate)); }

    // Parameters in the reference's container order (GpuDC::from_host / to_host on every container).
    void set_params(const std::vector<Real>& flat) {
        if (flat.size() != param_count()) throw std::runtime_error("Collections must be of the same size");
        check(dlb_net_set_params(handle(), flat.data()));
    }
    std::vector<Real> get_params() {
        std::vector<Real> p(param_count());
        check(dlb_net_get_params(handle(), p.data()));
        return p;
    }

    void reset_random_generator(unsigned seed) { check(dlb_net_reset_random_generator(handle(), seed)); }   // Net.inl:599-603

    // Net::learn(training_items, reference_items, batch_size, epochs_count, learning_rate, cost_func_id, lambda,
    //            epoch_callback) — Net.inl:193-324.  Items are rows of `training_items` ([n][in_size]).
    void learn(const std::vector<Real>& training_items, const std::vector<Real>& reference_items, std::size_t batch_size,
               std::size_t epochs_count, Real learning_rate, CostFunctionId cost_func_id, Real lambda = Real(0),
               const std::function<void(std::size_t, Real)>& epoch_callback = [](std::size_t, Real) {}) {
        const std::size_t in_n = (std::size_t)in_size().coord_prod(), out_n = (std::size_t)out_size().coord_prod();
        const std::size_t n = training_items.size() / in_n;
        if (training_items.size() != n * in_n || reference_items.size() != n * out_n)
            throw std::runtime_error("Incompatible collection of training and reference items.");
        if (n == 0 || batch_size == 0) return;
        check(dlb_net_set_dataset(handle(), training_items.data(), reference_items.data(), (int64_t)n, 0));
        Trampoline t{&epoch_callback};
        check(dlb_net_learn(handle(), (int64_t)batch_size, (int64_t)epochs_count, (double)learning_rate, (int)cost_func_id,
                            (double)lambda, &Net::on_epoch, &t));
    }

    // Net::act for a batch of items; returns [n][out_size].
    std::vector<Real> act(const std::vector<Real>& items) {
        const std::size_t in_n = (std::size_t)in_size().coord_prod(), out_n = (std::size_t)out_size().coord_prod();
        const std::size_t n = items.size() / in_n;
        std::vector<Real> out(n * out_n);
        check(dlb_net_eval_host(handle(), items.data(), (int64_t)n, nullptr, out.data()));
        return out;
    }

    // Net::evaluate_cost_function_and_correct_answers (Net.inl:411-439).
    struct CostAndCorrectAnswers { Real Cost{}; Real CorrectAnswers{}; };
    CostAndCorrectAnswers evaluate_cost_function_and_correct_answers(const std::vector<Real>& test_input,
            const std::vector<Real>& reference_output, CostFunctionId cost_func_id, Real l2_reg_factor = Real(0)) {
        const std::size_t in_n = (std::size_t)in_size().coord_prod(), out_n = (std::size_t)out_size().coord_prod();
        const std::size_t n = test_input.size() / in_n;
        if (n == 0 || reference_output.size() != n * out_n) throw std::runtime_error("Invalid input.");
        double cost = 0, correct = 0;
        check(dlb_net_evaluate_host(handle(), test_input.data(), reference_output.data(), (int64_t)n, (int)cost_func_id,
                                    (double)l2_reg_factor, &cost, &correct));
        return {(Real)cost, (Real)correct};
    }

    // Net::count_correct_answers (Net.inl:441-472): arg-max of the output against arg-max of the label.
    std::size_t count_correct_answers(const std::vector<Real>& test_input, const std::vector<Real>& labels) {
        const std::size_t in_n = (std::size_t)in_size().coord_prod(), out_n = (std::size_t)out_size().coord_prod();
        const std::size_t n = test_input.size() / in_n;
        if (labels.size() != n * out_n) throw std::runtime_error("Invalid input.");
        std::vector<int32_t> cls(n);
        check(dlb_net_eval_host(handle(), test_input.data(), (int64_t)n, cls.data(), nullptr));
        std::size_t correct = 0;
        for (std::size_t i = 0; i < n; ++i) {
            std::size_t best = 0;
            for (std::size_t k = 1; k < out_n; ++k) if (labels[i * out_n + best] < labels[i * out_n + k]) best = k;
            correct += (std::size_t)cls[i] == best;
        }
        return correct;
    }
};

}  // namespace dlb200

// gpu_dc.hpp — the reference-side binding of libdlb200: what a maintainer of dragunovdenis/DeepLearning adds next to
// DeepLearningCuda/NeuralNet/DataContextCuda.{h,cpp} to put the B200 library behind the GPU data context.
//
// Compiled against the reference's OWN headers (-I <reference>/DeepLearning) and include/dlb200.h; no CUDA headers, no
// nvcc.  tests/cpp/ref_adapter_driver.cpp + tests/test_reference_adapter_gpu.py build it with g++ and run the reference's
// Net template through it on the GPU.
//
//   GpuDC                      the data-context policy (DataContextCuda.h:28-77): vector_t / matrix_t / tensor_t /
//                              index_array_t<I> as RAII device buffers over dlb_malloc / dlb_free with the reference's
//                              ownership conventions (grow-only resize, copy = device-to-device, move = pointer steal leaving
//                              {nullptr, 0}; BasicCudaCollection.cu:53-62, StandardTestUtils.h:257,282), to_host / from_host,
//                              supports_fast_convolution() == false.
//   to_spec / flatten_params / scatter_params / pack
//                              ALayer<CpuDC> -> dlb_layer_spec through the public layer API (get_type_id, in_size, out_size,
//                              weight_tensor_size, to_script); the reference's container order (NLayer.inl:180-193,
//                              CLayer.inl:199-213) IS the flat parameter layout of libdlb200.
//   learn(Net<CpuDC>&, ...)    the batched replacement of Net<D>::learn (Net.inl:193-324) for a net whose model lives in the
//                              reference's own Net object and whose data set lives on the device in GpuDC tensors: same arguments,
//                              same epoch permutations (the trainer's generator is seeded like Net::reset_random_generator,
//                              Net.inl:620-624), same tail-batch rule, same update; the trained parameters are written back
//                              into the reference net, which is then used as before (act, count_correct_answers, save ...).
//   count_correct_answers / evaluate_cost_function_and_correct_answers     Net.inl:411-472 on the device.
//
// What is NOT here: the per-sample member functions of CudaTensor / CudaMatrix (convolve, mul_add, ...).
// Net<GpuDC> itself therefore cannot be instantiated with this GpuDC - constructing an NLayer<GpuDC> instantiates its virtual
// act / backpropagate, which call those members - and the binding goes through Net<CpuDC> for the model.
#pragma once

#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include <nlohmann/json.hpp>

#include <NeuralNet/Net.h>
#include <NeuralNet/CLayer.h>
#include <NeuralNet/NLayer.h>
#include <NeuralNet/PLayer.h>

#include "../../include/dlb200.h"

namespace DeepLearning {
namespace dlb200_binding {

inline void check(int rc) {
    if (rc != 0) throw std::runtime_error(dlb_last_error());   // the reference throws std::exception(message) at the same places (CudaUtils.cpp:24-30)
}

// One context per thread, like the reference's cudaStreamPerThread convention (CudaUtils.h:43-108).
inline dlb_context* context() {
    struct Holder {
        dlb_context* c = nullptr;
        Holder() { check(dlb_context_create(0, nullptr, &c)); }
        ~Holder() { if (c) dlb_context_destroy(c); }
    };
    static thread_local Holder h;
    return h.c;
}

// Device buffer with the reference's collection conventions.
template <class T>
class DeviceBuffer {
    T* _data = nullptr;
    std::size_t _size = 0, _capacity = 0;
public:
    DeviceBuffer() = default;
    explicit DeviceBuffer(std::size_t n) { resize(n); }
    DeviceBuffer(const DeviceBuffer& o) { assign(o); }
    DeviceBuffer(DeviceBuffer&& o) noexcept : _data(o._data), _size(o._size), _capacity(o._capacity) { o._data = nullptr; o._size = o._capacity = 0; }
    DeviceBuffer& operator=(const DeviceBuffer& o) { if (this != &o) assign(o); return *this; }
    DeviceBuffer& operator=(DeviceBuffer&& o) noexcept {
        if (this != &o) { release(); _data = o._data; _size = o._size; _capacity = o._capacity; o._data = nullptr; o._size = o._capacity = 0; }
        return *this;
    }
    ~DeviceBuffer() { release(); }
    void release() { if (_data) { dlb_free(context(), _data); _data = nullptr; } _size = _capacity = 0; }
    void resize(std::size_t n) {   // grow-only: a same-size (or smaller) request never reallocates
        if (n > _capacity) {
            void* p = nullptr;
            check(dlb_malloc(context(), n * sizeof(T), &p));
            if (_data) dlb_free(context(), _data);
            _data = static_cast<T*>(p); _capacity = n;
        }
        _size = n;
    }
    void assign(const DeviceBuffer& o) { resize(o._size); if (_size) check(dlb_memcpy_d2d(context(), _data, o._data, _size * sizeof(T))); }
    void upload(const T* host, std::size_t n) { resize(n); if (n) { check(dlb_memcpy_h2d(context(), _data, host, n * sizeof(T))); check(dlb_context_synchronize(context())); } }
    void download(T* host) const { if (_size) check(dlb_memcpy_d2h(context(), host, _data, _size * sizeof(T))); }
    T* begin() { return _data; }
    const T* begin() const { return _data; }
    T* end() { return _data + _size; }
    const T* end() const { return _data + _size; }
    std::size_t size() const { return _size; }
    bool empty() const { return _size == 0; }
    std::vector<T> to_stdvector() const { std::vector<T> v(_size); download(v.data()); return v; }
};

class GpuVector {
    DeviceBuffer<Real> _buf; std::size_t _dim = 0;
public:
    GpuVector() = default;
    explicit GpuVector(const Vector& host) : _dim(host.dim()) { _buf.upload(host.begin(), host.size()); }
    std::size_t dim() const { return _dim; }
    Index3d size_3d() const { return { 1, 1, static_cast<long long>(_dim) }; }
    std::size_t size() const { return _buf.size(); }
    Real* begin() { return _buf.begin(); }
    const Real* begin() const { return _buf.begin(); }
    Vector to_host() const { Vector v(_dim, false); _buf.download(v.begin()); return v; }
};

class GpuMatrix {
    DeviceBuffer<Real> _buf; std::size_t _rows = 0, _cols = 0;
public:
    GpuMatrix() = default;
    explicit GpuMatrix(const Matrix& host) : _rows(host.row_dim()), _cols(host.col_dim()) { _buf.upload(host.begin(), host.size()); }
    std::size_t row_dim() const { return _rows; }
    std::size_t col_dim() const { return _cols; }
    Index3d size_3d() const { return { 1, static_cast<long long>(_rows), static_cast<long long>(_cols) }; }
    std::size_t size() const { return _buf.size(); }
    Real* begin() { return _buf.begin(); }
    const Real* begin() const { return _buf.begin(); }
    Matrix to_host() const { Matrix m(_rows, _cols, false); _buf.download(m.begin()); return m; }
};

class GpuTensor {
    DeviceBuffer<Real> _buf; Index3d _size{ 0, 0, 0 };
public:
    GpuTensor() = default;
    GpuTensor(const Index3d& size, bool fill_zero = true) { resize(size); if (fill_zero && _buf.size()) check(dlb_memset(context(), _buf.begin(), 0, _buf.size() * sizeof(Real))); }
    explicit GpuTensor(const Tensor& host) : _size(host.size_3d()) { _buf.upload(host.begin(), host.size()); }
    void resize(const Index3d& size) { _buf.resize(static_cast<std::size_t>(size.coord_prod())); _size = size; }
    Index3d size_3d() const { return _size; }
    std::size_t size() const { return _buf.size(); }
    bool empty() const { return _buf.empty(); }
    Real* begin() { return _buf.begin(); }
    const Real* begin() const { return _buf.begin(); }
    Tensor to_host() const { Tensor t(_size, false); _buf.download(t.begin()); return t; }
};

template <class I> using GpuArray = DeviceBuffer<I>;

}  // namespace dlb200_binding

/// The GPU data context backed by libdlb200 (DataContextCuda.h:28-77).
class GpuDC {
public:
    using vector_t = dlb200_binding::GpuVector;
    using matrix_t = dlb200_binding::GpuMatrix;
    using tensor_t = dlb200_binding::GpuTensor;
    template <typename I> using index_array_t = dlb200_binding::GpuArray<I>;
    static CpuDC::vector_t to_host(const vector_t& s) { return s.to_host(); }
    static CpuDC::matrix_t to_host(const matrix_t& s) { return s.to_host(); }
    static CpuDC::tensor_t to_host(const tensor_t& s) { return s.to_host(); }
    static vector_t from_host(const CpuDC::vector_t& s) { return vector_t(s); }
    static matrix_t from_host(const CpuDC::matrix_t& s) { return matrix_t(s); }
    static tensor_t from_host(const CpuDC::tensor_t& s) { return tensor_t(s); }
    static bool constexpr supports_fast_convolution() { return false; }
};

namespace dlb200_binding {

inline Index3d parse_index3d(const std::string& s) {   // "{a, b, c}" as Index3d::to_string writes it
    long long v[3] = { 0, 0, 0 }; int k = 0; bool in_num = false, neg = false;
    for (char ch : s) {
        if (ch == '-') neg = true;
        else if (ch >= '0' && ch <= '9') { if (!in_num) { in_num = true; if (k < 3) v[k] = 0; } if (k < 3) v[k] = v[k] * 10 + (ch - '0'); }
        else { if (in_num) { if (k < 3 && neg) v[k] = -v[k]; ++k; in_num = false; neg = false; } }
    }
    if (in_num) { if (k < 3 && neg) v[k] = -v[k]; ++k; }
    if (k != 3) throw std::runtime_error("Can't parse Index3d: " + s);   // (the reference writes "{a, b, c}", LinAlg3d.cpp:32-35)
    return { v[0], v[1], v[2] };
}

/// ALayer -> dlb_layer_spec through the public layer API only; LayerTypeId.h:27-32, ActivationFunctionId.h:26-33 and
/// PoolOperator.h:28-33 ids map 1:1 onto the DLB_* constants.
inline dlb_layer_spec to_spec(const ALayer<CpuDC>& layer) {
    dlb_layer_spec s{};
    const auto in = layer.in_size(), out = layer.out_size(), w = layer.weight_tensor_size();
    const auto json = nlohmann::json::parse(layer.to_script());
    s.in_c = static_cast<int32_t>(in.x); s.in_h = static_cast<int32_t>(in.y); s.in_w = static_cast<int32_t>(in.z);
    s.stride_h = s.stride_w = 1;
    switch (layer.get_type_id()) {
    case LayerTypeId::FULL:
        s.type = DLB_LAYER_FULL; s.out_c = static_cast<int32_t>(out.coord_prod());
        s.activation = static_cast<int32_t>(parse_activation_type(json.at("Activation").get<std::string>()));
        break;
    case LayerTypeId::CONVOLUTION: {
        s.type = DLB_LAYER_CONVOLUTION; s.out_c = static_cast<int32_t>(out.x);
        s.win_h = static_cast<int32_t>(w.y); s.win_w = static_cast<int32_t>(w.z);
        s.activation = static_cast<int32_t>(parse_activation_type(json.at("Activation").get<std::string>()));
        Index3d pad{}, stride{};
        try { pad = parse_index3d(json.at("Paddings").get<std::string>()); stride = parse_index3d(json.at("Strides").get<std::string>()); }
        catch (const std::runtime_error& e) { throw std::runtime_error(std::string(e.what()) + " in layer script " + layer.to_script()); }
        s.pad_h = static_cast<int32_t>(pad.y); s.pad_w = static_cast<int32_t>(pad.z);
        s.stride_h = static_cast<int32_t>(stride.y); s.stride_w = static_cast<int32_t>(stride.z);
        break;
    }
    case LayerTypeId::PULL: {
        s.type = DLB_LAYER_PULL; s.win_h = static_cast<int32_t>(w.y); s.win_w = static_cast<int32_t>(w.z);
        s.stride_h = s.win_h; s.stride_w = s.win_w;
        const auto op = json.at("PoolOperator").get<std::string>();
        s.pool_type = op == DeepLearning::to_string(PoolTypeId::MAX) ? DLB_POOL_MAX : op == DeepLearning::to_string(PoolTypeId::MIN) ? DLB_POOL_MIN : DLB_POOL_AVERAGE;
        break;
    }
    default: throw std::runtime_error("to_spec: unsupported layer type");
    }
    return s;
}

/// The trainable containers of a net in the reference's order: per layer [biases, weights] / [biases, filter_0 .. filter_{F-1}].
template <class F>
void for_each_param_container(const Net<CpuDC>& net, F&& fn) {
    for (std::size_t l = 0; l < net.layers_count(); ++l) {
        const ALayer<CpuDC>* layer = &net[l];
        if (auto n = dynamic_cast<const NLayer<CpuDC>*>(layer)) { fn(n->biases()); fn(n->weights()); }
        else if (auto c = dynamic_cast<const CLayer<CpuDC>*>(layer)) { fn(c->biases()); for (const auto& f : c->filters()) fn(f); }
    }
}

inline std::vector<Real> flatten_params(const Net<CpuDC>& net) {
    std::vector<Real> flat;
    for_each_param_container(net, [&](const BasicCollection& c) { flat.insert(flat.end(), c.begin(), c.end()); });
    return flat;
}

/// Writes a flat parameter vector back into the net's own containers (the layers expose them read-only; their storage is the net's).
inline void scatter_params(Net<CpuDC>& net, const std::vector<Real>& flat) {
    std::size_t off = 0;
    for_each_param_container(net, [&](const BasicCollection& c) {
        if (off + c.size() > flat.size()) throw std::runtime_error("scatter_params: parameter count mismatch");
        std::memcpy(const_cast<Real*>(c.begin()), flat.data() + off, c.size() * sizeof(Real));
        off += c.size();
    });
    if (off != flat.size()) throw std::runtime_error("scatter_params: parameter count mismatch");
}

/// [n][size] device array of n equally sized device tensors (n device-to-device copies).
inline DeviceBuffer<Real> pack(const std::vector<GpuDC::tensor_t>& items) {
    DeviceBuffer<Real> out;
    if (items.empty()) return out;
    const std::size_t sz = items[0].size();
    out.resize(items.size() * sz);
    for (std::size_t i = 0; i < items.size(); ++i) {
        if (items[i].size() != sz) throw std::runtime_error("Incompatible collection of training and reference items.");
        check(dlb_memcpy_d2d(context(), out.begin() + i * sz, items[i].begin(), sz * sizeof(Real)));
    }
    return out;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Per-sample layer API on GpuDC containers: what ALayer<GpuDC>::act / backpropagate / update do for ONE item
// (ALayer.h:205-257; GPU originals: CudaTensor::convolve / convolution_gradient / min_max_pool / average_pool ..., CudaMatrix::mul_add /
// transpose_mul, CudaTensor.cu:395-761, CudaMatrix.cu:227-498).  Every call is one launch of the library's batched operator with
// batch = 1; the trace / gradient arguments are the reference's own LayerTraceData<GpuDC> / LayerData<GpuDC> / LayerGradient<GpuDC>.
// The batched trainer (learn above) is the fast path; this is the compatibility surface for code written against single items.
// ---------------------------------------------------------------------------------------------------------------------------------
namespace ops {   // Real -> the _f32 / _f64 entry points
inline int conv_forward(dlb_context* c, const dlb_layer_spec* s, const float* x, const float* b, const float* w, float* y, float* d) { return dlb_conv_forward_f32(c, s, 1, x, nullptr, b, w, y, d); }
inline int conv_forward(dlb_context* c, const dlb_layer_spec* s, const double* x, const double* b, const double* w, double* y, double* d) { return dlb_conv_forward_f64(c, s, 1, x, nullptr, b, w, y, d); }
inline int conv_backward(dlb_context* c, const dlb_layer_spec* s, const float* x, const float* dl, const float* dv, const float* w, float* g, float* db, float* dw, float* dx, double sc) { return dlb_conv_backward_f32(c, s, 1, x, nullptr, dl, dv, w, g, db, dw, dx, sc); }
inline int conv_backward(dlb_context* c, const dlb_layer_spec* s, const double* x, const double* dl, const double* dv, const double* w, double* g, double* db, double* dw, double* dx, double sc) { return dlb_conv_backward_f64(c, s, 1, x, nullptr, dl, dv, w, g, db, dw, dx, sc); }
inline int pool_forward(dlb_context* c, const dlb_layer_spec* s, const float* x, float* y, int32_t* m) { return dlb_pool_forward_f32(c, s, 1, x, y, m); }
inline int pool_forward(dlb_context* c, const dlb_layer_spec* s, const double* x, double* y, int32_t* m) { return dlb_pool_forward_f64(c, s, 1, x, y, m); }
inline int pool_backward(dlb_context* c, const dlb_layer_spec* s, const float* dl, const int32_t* m, float* dx) { return dlb_pool_backward_f32(c, s, 1, dl, m, dx); }
inline int pool_backward(dlb_context* c, const dlb_layer_spec* s, const double* dl, const int32_t* m, double* dx) { return dlb_pool_backward_f64(c, s, 1, dl, m, dx); }
inline int full_forward(dlb_context* c, const dlb_layer_spec* s, const float* x, const float* b, const float* w, float* y, float* d) { return dlb_full_forward_f32(c, s, 1, x, nullptr, b, w, y, d); }
inline int full_forward(dlb_context* c, const dlb_layer_spec* s, const double* x, const double* b, const double* w, double* y, double* d) { return dlb_full_forward_f64(c, s, 1, x, nullptr, b, w, y, d); }
inline int full_backward(dlb_context* c, const dlb_layer_spec* s, const float* x, const float* dl, const float* dv, const float* w, float* g, float* db, float* dw, float* dx, double sc) { return dlb_full_backward_f32(c, s, 1, x, nullptr, dl, dv, w, g, db, dw, dx, sc); }
inline int full_backward(dlb_context* c, const dlb_layer_spec* s, const double* x, const double* dl, const double* dv, const double* w, double* g, double* db, double* dw, double* dx, double sc) { return dlb_full_backward_f64(c, s, 1, x, nullptr, dl, dv, w, g, db, dw, dx, sc); }
inline int sgd_update(dlb_context* c, float* p, const float* g, int64_t n, int reg, double lr, double rf, double gs) { return dlb_sgd_update_f32(c, p, g, n, reg, lr, rf, gs); }
inline int sgd_update(dlb_context* c, double* p, const double* g, int64_t n, int reg, double lr, double rf, double gs) { return dlb_sgd_update_f64(c, p, g, n, reg, lr, rf, gs); }
inline int cost_gradient(dlb_context* c, int id, int n, const float* o, const float* r, float* g) { return dlb_cost_gradient_f32(c, id, 1, n, o, r, nullptr, g); }
inline int cost_gradient(dlb_context* c, int id, int n, const double* o, const double* r, double* g) { return dlb_cost_gradient_f64(c, id, 1, n, o, r, nullptr, g); }
}  // namespace ops

/// One reference layer mirrored on the device: its descriptor and its parameters [biases | weights or filters] in one buffer.
class GpuLayer {
    dlb_layer_spec _spec{};
    DeviceBuffer<Real> _params;
    std::size_t _bias_n = 0, _weight_n = 0;
    mutable DeviceBuffer<Real> _g;   // dC/dz scratch of backpropagate
    static std::vector<const BasicCollection*> containers(const ALayer<CpuDC>& layer) {
        std::vector<const BasicCollection*> out;
        if (auto n = dynamic_cast<const NLayer<CpuDC>*>(&layer)) { out.push_back(&n->biases()); out.push_back(&n->weights()); }
        else if (auto c = dynamic_cast<const CLayer<CpuDC>*>(&layer)) { out.push_back(&c->biases()); for (const auto& f : c->filters()) out.push_back(&f); }
        return out;
    }
public:
    explicit GpuLayer(const ALayer<CpuDC>& layer) : _spec(to_spec(layer)) {
        std::vector<Real> flat;
        const auto cs = containers(layer);
        for (std::size_t i = 0; i < cs.size(); ++i) { if (i == 0) _bias_n = cs[i]->size(); else _weight_n += cs[i]->size(); flat.insert(flat.end(), cs[i]->begin(), cs[i]->end()); }
        if (!flat.empty()) _params.upload(flat.data(), flat.size());
    }
    const dlb_layer_spec& spec() const { return _spec; }
    Index3d in_size() const { return { _spec.in_c, _spec.in_h, _spec.in_w }; }
    Index3d out_size(const ALayer<CpuDC>& like) const { return like.out_size(); }
    const Real* biases() const { return _params.begin(); }
    const Real* weights() const { return _params.begin() + _bias_n; }

    /// ALayer<D>::act for one item (CLayer.inl:108-137, PLayer.inl:115-138, NLayer.inl:131-142); `output` must have the layer's output size
    void act(const GpuTensor& input, GpuTensor& output, LayerTraceData<GpuDC>* const trace) const {
        Real* deriv = nullptr; int32_t* map = nullptr;
        if (trace && _spec.type != DLB_LAYER_PULL) { trace->Derivatives.resize(output.size_3d()); deriv = trace->Derivatives.begin(); }
        if (trace && _spec.type == DLB_LAYER_PULL && _spec.pool_type != DLB_POOL_AVERAGE) { trace->IndexMapping.resize(output.size()); map = trace->IndexMapping.begin(); }
        switch (_spec.type) {
        case DLB_LAYER_CONVOLUTION: check(ops::conv_forward(context(), &_spec, input.begin(), biases(), weights(), output.begin(), deriv)); break;
        case DLB_LAYER_PULL: check(ops::pool_forward(context(), &_spec, input.begin(), output.begin(), map)); break;
        default: check(ops::full_forward(context(), &_spec, input.begin(), biases(), weights(), output.begin(), deriv)); break;
        }
    }
    /// ALayer<D>::backpropagate for one item (CLayer.inl:139-197, PLayer.inl:140-168, NLayer.inl:144-171): layer_grad.data = {biases, weights}
    /// as ONE device tensor each (filters concatenated), scaled accumulation as in the reference (gradient_scale_factor)
    void backpropagate(const GpuTensor& deltas, const LayerData<GpuDC>& processing_data, GpuTensor& input_grad, LayerGradient<GpuDC>& layer_grad,
                       const bool evaluate_input_gradient = true, const Real gradient_scale_factor = static_cast<Real>(0)) const {
        Real* dx = nullptr;
        if (evaluate_input_gradient) { input_grad.resize(in_size()); dx = input_grad.begin(); }
        if (_spec.type == DLB_LAYER_PULL) {
            if (dx) check(ops::pool_backward(context(), &_spec, deltas.begin(), processing_data.Trace.IndexMapping.begin(), dx));
            return;
        }
        if (layer_grad.data.size() != 2) {
            layer_grad.data.resize(2);
            layer_grad.data[0] = GpuTensor(Index3d{ 1, 1, static_cast<long long>(_bias_n) }, true);
            layer_grad.data[1] = GpuTensor(Index3d{ 1, 1, static_cast<long long>(_weight_n) }, true);
        }
        _g.resize(deltas.size());
        if (_spec.type == DLB_LAYER_CONVOLUTION)
            check(ops::conv_backward(context(), &_spec, processing_data.Input.begin(), deltas.begin(), processing_data.Trace.Derivatives.begin(), weights(), _g.begin(),
                                     layer_grad.data[0].begin(), layer_grad.data[1].begin(), dx, gradient_scale_factor));
        else
            check(ops::full_backward(context(), &_spec, processing_data.Input.begin(), deltas.begin(), processing_data.Trace.Derivatives.begin(), weights(), _g.begin(),
                                     layer_grad.data[0].begin(), layer_grad.data[1].begin(), dx, gradient_scale_factor));
    }
    /// ALayer<D>::update (ALayer.inl:97-122): biases unregularised, weights with the regularisation factor; l_rate is signed as in the reference
    void update(const LayerGradient<GpuDC>& gradient, const Real& l_rate, const Real& reg_factor, const Real& grad_scale_factor = static_cast<Real>(1)) {
        if (_spec.type == DLB_LAYER_PULL) return;
        check(ops::sgd_update(context(), _params.begin(), gradient.data[0].begin(), static_cast<int64_t>(_bias_n), 0, l_rate, reg_factor, grad_scale_factor));
        check(ops::sgd_update(context(), _params.begin() + _bias_n, gradient.data[1].begin(), static_cast<int64_t>(_weight_n), 1, l_rate, reg_factor, grad_scale_factor));
    }
    /// the parameters as one host vector [biases | weights] (the layer's containers in the reference's order)
    std::vector<Real> params_to_host() const { return _params.to_stdvector(); }
};

/// CostFunction<T>::deriv_in_place for one item on the device (CostFunction.inl:47-51)
inline void cost_deriv_in_place(GpuTensor& in_out, const GpuTensor& reference, const CostFunctionId& id) {
    check(ops::cost_gradient(context(), static_cast<int>(id), static_cast<int>(in_out.size()), in_out.begin(), reference.begin(), in_out.begin()));
}

/// A libdlb200 net mirroring a reference net (layers, parameters, keep rates).
class DeviceNet {
    dlb_net* _net = nullptr;
public:
    explicit DeviceNet(const Net<CpuDC>& net) {
        std::vector<dlb_layer_spec> specs;
        for (std::size_t l = 0; l < net.layers_count(); ++l) specs.push_back(to_spec(net[l]));
        check(dlb_net_create(context(), specs.data(), static_cast<int>(specs.size()), static_cast<int>(sizeof(Real)), &_net));
        const auto flat = flatten_params(net);
        if (static_cast<int64_t>(flat.size()) != dlb_net_param_count(_net)) throw std::runtime_error("DeviceNet: parameter count mismatch");
        check(dlb_net_set_params(_net, flat.data()));
        for (std::size_t l = 0; l < net.layers_count(); ++l)   // dropout: ALayer::get_keep_rate, ALayer.inl:169-172
            if (net[l].get_keep_rate() < Real(1)) check(dlb_net_set_keep_rate(_net, static_cast<int>(l), net[l].get_keep_rate()));
    }
    ~DeviceNet() { if (_net) dlb_net_destroy(_net); }
    DeviceNet(const DeviceNet&) = delete;
    DeviceNet& operator=(const DeviceNet&) = delete;
    dlb_net* handle() const { return _net; }
    std::vector<Real> params() const { std::vector<Real> flat(static_cast<std::size_t>(dlb_net_param_count(_net))); check(dlb_net_get_params(_net, flat.data())); return flat; }
};

/// The reference's gradient containers - one LayerGradient per layer, data = {biases, weights | filters...} (LayerGradient.h:29-87,
/// Net::allocate Net.h:291) - filled from a flat gradient in the library's parameter order.
inline std::vector<LayerGradient<CpuDC>> to_layer_gradients(const Net<CpuDC>& net, const std::vector<Real>& flat) {
    std::vector<LayerGradient<CpuDC>> out;
    net.allocate(out, /*fill zeros*/ false);
    std::size_t off = 0;
    for (auto& lg : out)
        for (auto& t : lg.data) {
            if (off + t.size() > flat.size()) throw std::runtime_error("to_layer_gradients: parameter count mismatch");
            std::memcpy(t.begin(), flat.data() + off, t.size() * sizeof(Real));
            off += t.size();
        }
    if (off != flat.size()) throw std::runtime_error("to_layer_gradients: parameter count mismatch");
    return out;
}

/// [n][size] host array of n equally sized host tensors
inline std::vector<Real> pack_host(const std::vector<Tensor>& items) {
    std::vector<Real> out;
    if (items.empty()) return out;
    const std::size_t sz = items[0].size();
    out.resize(items.size() * sz);
    for (std::size_t i = 0; i < items.size(); ++i) {
        if (items[i].size() != sz) throw std::runtime_error("Incompatible collection of training and reference items.");
        std::memcpy(out.data() + i * sz, items[i].begin(), sz * sizeof(Real));
    }
    return out;
}

}  // namespace dlb200_binding

/// The sum over the items of Net<D>::calc_gradient_and_value's gradient (Net.inl:339-375; what learn accumulates per mini-batch,
/// Net.inl:263-276), in the reference's containers.
inline std::vector<LayerGradient<CpuDC>> calc_gradient_sum(const Net<CpuDC>& net, const std::vector<Tensor>& items, const std::vector<Tensor>& target_values,
                                                           const CostFunctionId& cost_func_id) {
    using namespace dlb200_binding;
    if (items.size() != target_values.size() || items.empty()) throw std::runtime_error("Invalid input.");
    DeviceNet dev(net);
    const auto x = pack_host(items), y = pack_host(target_values);
    check(dlb_net_forward_backward_host(dev.handle(), x.data(), y.data(), static_cast<int64_t>(items.size()), static_cast<int>(cost_func_id), /*scale*/ 0.0));
    std::vector<Real> flat(static_cast<std::size_t>(dlb_net_param_count(dev.handle())));
    check(dlb_net_read(dev.handle(), DLB_BUF_GRADIENT, 0, flat.data(), static_cast<int64_t>(flat.size() * sizeof(Real))));
    return to_layer_gradients(net, flat);
}

/// What Net<D>::act_bpg leaves in its context for ONE item (Net.inl:117-132): per layer LayerData{Input, Trace{Derivatives, IndexMapping}}
/// (LayerData.h:28-74), read back from the device traces; `out_value` receives the output of the last layer.
inline std::vector<LayerData<CpuDC>> act_bpg_traces(const Net<CpuDC>& net, const Tensor& item, const Tensor& target_value, const CostFunctionId& cost_func_id,
                                                    Tensor& out_value) {
    using namespace dlb200_binding;
    DeviceNet dev(net);
    check(dlb_net_forward_backward_host(dev.handle(), item.begin(), target_value.begin(), 1, static_cast<int>(cost_func_id), 0.0));
    const auto L = net.layers_count();
    std::vector<LayerData<CpuDC>> out(L);
    for (std::size_t l = 0; l < L; ++l) {
        const auto in3 = net[l].in_size(), out3 = net[l].out_size();
        if (l == 0) out[l].Input = item;
        else {   // the input of layer l is the output of layer l - 1, in layer l's own shape
            out[l].Input = Tensor(in3, false);
            check(dlb_net_read(dev.handle(), DLB_BUF_OUTPUT, static_cast<int>(l) - 1, out[l].Input.begin(), static_cast<int64_t>(out[l].Input.size() * sizeof(Real))));
        }
        const auto n_out = static_cast<std::size_t>(out3.coord_prod());
        if (dynamic_cast<const PLayer<CpuDC>*>(&net[l])) {   // a min / max PLayer keeps the arg-max map, an average one nothing (PLayer.inl:115-138)
            if (to_spec(net[l]).pool_type != DLB_POOL_AVERAGE) {
                std::vector<int32_t> map(n_out);
                check(dlb_net_read(dev.handle(), DLB_BUF_INDEX_MAP, static_cast<int>(l), map.data(), static_cast<int64_t>(n_out * sizeof(int32_t))));
                out[l].Trace.IndexMapping.assign(map.begin(), map.end());
            }
        } else {
            out[l].Trace.Derivatives = Tensor(out3, false);
            check(dlb_net_read(dev.handle(), DLB_BUF_DERIVATIVES, static_cast<int>(l), out[l].Trace.Derivatives.begin(), static_cast<int64_t>(n_out * sizeof(Real))));
        }
    }
    out_value = Tensor(net[L - 1].out_size(), false);
    check(dlb_net_read(dev.handle(), DLB_BUF_OUTPUT, static_cast<int>(L) - 1, out_value.begin(), static_cast<int64_t>(out_value.size() * sizeof(Real))));
    return out;
}

/// Net<D>::learn (Net.inl:193-324) for a model held in the reference's Net object and a data set held in GpuDC tensors.
/// `seed`: what Net::reset_random_generator(seed) would have been called with before the reference's own learn - the epoch
/// permutations (and dropout masks) are then the ones the reference draws.
inline void learn(Net<CpuDC>& net, const std::vector<GpuDC::tensor_t>& training_items, const std::vector<GpuDC::tensor_t>& reference_items,
                  const std::size_t batch_size, const std::size_t epochs_count, const Real learning_rate, const CostFunctionId& cost_func_id,
                  const Real& lambda, const std::function<void(const std::size_t, const Real)>& epoch_callback, const unsigned seed) {
    using namespace dlb200_binding;
    if (training_items.size() != reference_items.size())
        throw std::runtime_error("Incompatible collection of training and reference items.");      // Net.inl:197-198
    if (training_items.empty() || batch_size == 0) return;                                          // Net.inl:200-201
    DeviceNet dev(net);
    const auto items = pack(training_items), refs = pack(reference_items);
    check(dlb_net_set_dataset(dev.handle(), items.begin(), refs.begin(), static_cast<int64_t>(training_items.size()), /*on_device*/ 1));
    check(dlb_net_reset_random_generator(dev.handle(), seed));
    struct Ctx { const std::function<void(const std::size_t, const Real)>* cb; } user{ &epoch_callback };
    check(dlb_net_learn(dev.handle(), static_cast<int64_t>(batch_size), static_cast<int64_t>(epochs_count), learning_rate,
                        static_cast<int>(cost_func_id), lambda,
                        [](int64_t e, double reg, void* u) { const auto* cb = static_cast<Ctx*>(u)->cb; if (*cb) (*cb)(static_cast<std::size_t>(e), static_cast<Real>(reg)); },
                        &user));
    scatter_params(net, dev.params());
}

/// Net<D>::evaluate_cost_function_and_correct_answers (Net.inl:411-439) on the device; host items in, as the callers have them.
inline CostAndCorrectAnswers evaluate_cost_function_and_correct_answers(const Net<CpuDC>& net, const std::vector<Tensor>& test_input,
        const std::vector<Tensor>& reference_output, const CostFunctionId& cost_func_id, const Real l2_reg_factor) {
    using namespace dlb200_binding;
    if (test_input.size() != reference_output.size() || test_input.empty()) throw std::runtime_error("Invalid input.");
    DeviceNet dev(net);
    const std::size_t in = test_input[0].size(), out = reference_output[0].size();
    std::vector<Real> x(test_input.size() * in), y(test_input.size() * out);
    for (std::size_t i = 0; i < test_input.size(); ++i) {
        std::memcpy(x.data() + i * in, test_input[i].begin(), in * sizeof(Real));
        std::memcpy(y.data() + i * out, reference_output[i].begin(), out * sizeof(Real));
    }
    double cost = 0, correct = 0;
    check(dlb_net_evaluate_host(dev.handle(), x.data(), y.data(), static_cast<int64_t>(test_input.size()), static_cast<int>(cost_func_id), l2_reg_factor, &cost, &correct));
    CostAndCorrectAnswers result{};
    result.Cost = static_cast<Real>(cost); result.CorrectAnswers = static_cast<Real>(correct);
    return result;
}

}  // namespace DeepLearning

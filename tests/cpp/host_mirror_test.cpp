// tests/cpp/host_mirror_test.cpp — the C++ host mirror (deeplearning_b200/host/dlb200.hpp) used the way reference
// client code uses Net<D>, checked against the CPU oracle (oracle/dl_oracle.c, linked here as the checker).
// Restates the shape of the reference's layer/net tests (DeepLearningCudaTest/NetLayersCudaTest.cpp:77-172: GPU vs CPU
// on the same inputs): build LeNet-mini, copy identical parameters to both sides, learn, compare weights + classes.
// Built and run by tests/test_cpp_host_gpu.py (g++ only; no CUDA headers needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../deeplearning_b200/host/dlb200.hpp"
#include "../../oracle/dl_oracle.h"

using namespace dlb200;
using Real = orc_real;   // float with -DORACLE_SINGLE_PRECISION, else double: the same `Real` on both sides

int main() {
    try {
        Context ctx(0);
        Net<Real> net(ctx);
        Index3d sz{1, 12, 12};
        sz = net.append_layer<CLayer>(sz, Index2d{3, 3}, 4, ActivationFunctionId::RELU);
        sz = net.append_layer<PLayer>(sz, Index2d{2, 2}, PoolTypeId::MAX);
        sz = net.append_layer<NLayer>((std::size_t)sz.coord_prod(), 16, ActivationFunctionId::RELU);
        sz = net.append_layer<NLayer>((std::size_t)sz.coord_prod(), 10, ActivationFunctionId::SOFTMAX);
        if (!(sz == Index3d{1, 1, 10})) { std::printf("FAIL: unexpected output size\n"); return 1; }

        // the same net in the oracle
        std::vector<dlb_layer_spec> specs = {
            CLayer::make({1, 12, 12}, {3, 3}, 4, ActivationFunctionId::RELU), PLayer::make({4, 10, 10}, {2, 2}, PoolTypeId::MAX),
            NLayer::make(100, 16, ActivationFunctionId::RELU), NLayer::make(16, 10, ActivationFunctionId::SOFTMAX)};
        specs[2].in_c = 4; specs[2].in_h = 5; specs[2].in_w = 5;
        orc_net* orc = orc_net_create(specs.data(), (int)specs.size());
        if (!orc) { std::printf("FAIL: oracle: %s\n", orc_last_error()); return 1; }
        const std::size_t P = net.param_count();
        if ((int64_t)P != orc_net_param_count(orc)) { std::printf("FAIL: parameter count\n"); return 1; }

        std::mt19937 gen(5);
        std::uniform_real_distribution<double> u(-0.5, 0.5), u01(0.0, 1.0);
        std::vector<Real> params(P);
        for (auto& p : params) p = (Real)u(gen);
        net.set_params(params);
        orc_net_set_params(orc, params.data());

        const std::size_t n = 40, batch = 8, epochs = 2;
        std::vector<Real> x(n * 144), y(n * 10, Real(0));
        for (auto& v : x) v = (Real)u01(gen);
        for (std::size_t i = 0; i < n; ++i) y[i * 10 + i % 10] = Real(1);

        const unsigned seed = 7;
        std::vector<int64_t> perms(epochs * n);
        check(dlb_shuffle_sequence(seed, (int64_t)n, (int64_t)epochs, perms.data()));
        orc_net_learn(orc, x.data(), y.data(), (int64_t)n, (int64_t)batch, (int64_t)epochs, 0.05, DLB_COST_CROSS_ENTROPY, 0.1, 1, perms.data());

        std::size_t callbacks = 0;
        net.reset_random_generator(seed);
        net.learn(x, y, batch, epochs, Real(0.05), CostFunctionId::CROSS_ENTROPY, Real(0.1), [&](std::size_t, Real) { ++callbacks; });
        if (callbacks != epochs) { std::printf("FAIL: epoch callback ran %zu times\n", callbacks); return 1; }

        std::vector<Real> want(P);
        orc_net_get_params(orc, want.data());
        const std::vector<Real> got = net.get_params();
        double max_ref = 0, max_err = 0;
        for (std::size_t i = 0; i < P; ++i) { max_ref = std::fmax(max_ref, std::fabs((double)want[i])); max_err = std::fmax(max_err, std::fabs((double)got[i] - (double)want[i])); }
        const double rtol = sizeof(Real) == 4 ? 2e-4 : 2e-11;
        std::printf("parameters after learn: max |err| %.3e (max |ref| %.3f)\n", max_err, max_ref);
        if (!(max_err <= rtol * max_ref)) { std::printf("FAIL: parameters differ\n"); return 1; }

        std::vector<int32_t> cls(n);
        orc_net_eval(orc, x.data(), (int64_t)n, cls.data(), nullptr);
        std::vector<Real> labels(n * 10, Real(0));
        for (std::size_t i = 0; i < n; ++i) labels[i * 10 + (std::size_t)cls[i]] = Real(1);
        if (net.count_correct_answers(x, labels) != n) { std::printf("FAIL: predicted classes differ from the oracle's\n"); return 1; }

        const auto stats = net.evaluate_cost_function_and_correct_answers(x, labels, CostFunctionId::CROSS_ENTROPY);
        if (stats.CorrectAnswers != Real(1) || !(stats.Cost >= Real(0))) { std::printf("FAIL: evaluate_cost_function_and_correct_answers\n"); return 1; }

        // error behaviour: inconsistent layers throw with the reference's message
        bool threw = false;
        try { Net<Real> bad(ctx); bad.append_layer<NLayer>(10, 5); bad.append_layer<NLayer>(7, 3); bad.param_count(); }
        catch (const std::runtime_error& e) { threw = std::string(e.what()).find("Inconsistent input/output dimensions") != std::string::npos; }
        if (!threw) { std::printf("FAIL: inconsistent layers did not throw\n"); return 1; }

        orc_net_destroy(orc);
        std::printf("OK (%s)\n", net.describe_plan().c_str());
        return 0;
    } catch (const std::exception& e) {
        std::printf("FAIL: exception: %s\n", e.what());
        return 1;
    }
}

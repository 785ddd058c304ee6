"""The reference-side binding (deeplearning_b200/host/gpu_dc.hpp) compiled against the reference's OWN headers: the reference's
Net template drives libdlb200.

tests/cpp/ref_adapter_driver.cpp builds two identical Net<CpuDC> objects with the reference's own initialisers, trains one with
the reference's Net<CpuDC>::learn on the host and the other through GpuDC (data set as GpuDC tensors on the device, the batched
learn of the binding), and returns both.  The CPU part (converters) runs without a GPU; the training comparison is a gpu test.
The driver library is prebuilt in the build container (tests/cpp/Makefile, needs /root/reference) and travels with the snapshot."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import binding as ob
from tests import golden_cases as gc
from tests import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path(real):
    return os.path.join(HERE, "cpp", "_build", f"libref_adapter_{real}.so")


pytestmark = pytest.mark.skipif(not all(os.path.exists(lib_path(r)) for r in ("f32", "f64")),
                                reason="tests/cpp/_build/libref_adapter_*.so not built (needs /root/reference: make -C tests/cpp adapter)")

_libs = {}


def load(real):
    if real not in _libs:
        lib = C.CDLL(lib_path(real))
        lib.rad_last_error.restype = C.c_char_p
        lib.rad_param_count.restype = C.c_int64
        lib.rad_param_count.argtypes = [C.POINTER(ob.LayerSpec), C.c_int]
        lib.rad_check_converters.argtypes = [C.POINTER(ob.LayerSpec), C.c_int, C.c_uint]
        vp = C.c_void_p
        lib.rad_run.argtypes = [C.POINTER(ob.LayerSpec), C.c_int, C.c_uint, C.c_uint, vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int,
                                C.c_double, C.c_int, vp, vp, vp, vp, vp, vp, vp]
        assert lib.rad_real_bytes() == np.dtype(ob.dtype_of(real)).itemsize
        _libs[real] = lib
    return _libs[real]


NETS = {"mini_lenet": gc.MINI_LENET, "fixture": gc.FIXTURE_NET, "avg": gc.AVG_NET, "mini_mlp": gc.MINI_MLP, "lenet": ob.lenet_specs(),
        "cifar": ob.cifar_specs()}


# The GPU cases run in a FRESH process each: the float and the double build of the reference are two shared libraries with the same
# C++ symbols, and a process that has trained through one of them corrupts its heap when it then trains through the other (seen on the
# B200 as "munmap_chunk(): invalid pointer" inside std::regex of the second library, although both are built -fno-gnu-unique
# -Wl,-Bsymbolic and loaded RTLD_LOCAL); either one alone is clean under MALLOC_CHECK_=3.
def in_fresh_process(fn, *args):
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(1) as pool:
        return pool.apply(fn, args)


def _train_case(real, name, x, y, batch, epochs, lr, cost, lam, n_params):
    lib = load(real)
    specs = NETS[name]
    arr = (ob.LayerSpec * len(specs))(*specs)
    dt = ob.dtype_of(real)
    n_items = x.shape[0]
    p_cpu = np.zeros(n_params, dt); p_gpu = np.zeros(n_params, dt)
    c_cpu = np.zeros(n_items, np.int32); c_gpu = np.zeros(n_items, np.int32)
    seen = C.c_int32(0); ev_cost = C.c_double(0); ev_correct = C.c_double(0)
    rc = lib.rad_run(arr, len(specs), 7, 11, x.ctypes.data, y.ctypes.data, n_items, batch, epochs, lr, cost, lam, 1,
                     p_cpu.ctypes.data, p_gpu.ctypes.data, c_cpu.ctypes.data, c_gpu.ctypes.data, C.addressof(seen), C.addressof(ev_cost), C.addressof(ev_correct))
    return dict(rc=rc, error=lib.rad_last_error().decode(), p_cpu=p_cpu, p_gpu=p_gpu, c_cpu=c_cpu, c_gpu=c_gpu, seen=seen.value,
                ev_cost=ev_cost.value, ev_correct=ev_correct.value)


def _container_case(real, name, x, y, cost):
    lib = load(real)
    vp = C.c_void_p
    lib.rad_check_containers.argtypes = [C.POINTER(ob.LayerSpec), C.c_int, C.c_uint, vp, vp, C.c_int64, C.c_int, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    specs = NETS[name]
    arr = (ob.LayerSpec * len(specs))(*specs)
    eg, et, mism = C.c_double(-1), C.c_double(-1), C.c_int64(-1)
    rc = lib.rad_check_containers(arr, len(specs), 7, x.ctypes.data, y.ctypes.data, x.shape[0], cost, C.byref(eg), C.byref(et), C.byref(mism))
    return dict(rc=rc, error=lib.rad_last_error().decode(), eg=eg.value, et=et.value, mism=mism.value)


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("name", list(NETS))
def test_converters_round_trip(name, real):
    """to_spec recovers the layer descriptors from the reference's layers through their public API; flatten_params / scatter_params
    are inverse and use the reference's container order (no GPU involved)."""
    lib = load(real)
    specs = NETS[name]
    arr = (ob.LayerSpec * len(specs))(*specs)
    assert lib.rad_check_converters(arr, len(specs), 3) == 0, lib.rad_last_error().decode()
    orc = ob.OracleNet(specs, real)
    assert lib.rad_param_count(arr, len(specs)) == orc.param_count


@pytest.mark.gpu
@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("name,n_items,batch,epochs,lam", [("mini_lenet", 100, 16, 2, 0.1), ("fixture", 23, 5, 2, 1.5), ("lenet", 768, 256, 1, 0.3)])
def test_reference_net_trained_through_the_gpu_binding(name, n_items, batch, epochs, lam, real):
    """Net<CpuDC> trained by the reference on the host vs the same net trained through the GpuDC binding on the B200: parameters within
    the bar, the reference's own inference on both trained nets predicts the same classes."""
    lib = load(real)
    specs = NETS[name]
    arr = (ob.LayerSpec * len(specs))(*specs)
    dt = ob.dtype_of(real)
    n_params = lib.rad_param_count(arr, len(specs))
    in_n = specs[0].in_c * specs[0].in_h * specs[0].in_w
    out_shape = ob.out_shape(specs[-1])
    out_n = out_shape[0] * out_shape[1] * out_shape[2]
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, (n_items, in_n)).astype(dt)
    cost = ob.SQUARED_ERROR if specs[-1].activation in (ob.LINEAR, ob.TANH) else ob.CROSS_ENTROPY
    if cost == ob.CROSS_ENTROPY:
        y = np.zeros((n_items, out_n), dt); y[np.arange(n_items), rng.integers(0, out_n, n_items)] = 1
    else:
        y = rng.uniform(0.05, 0.95, (n_items, out_n)).astype(dt)
    out = in_fresh_process(_train_case, real, name, x, y, batch, epochs, 0.05, cost, lam, int(n_params))
    assert out["rc"] == 0, out["error"]
    p_cpu, p_gpu, c_cpu, c_gpu = out["p_cpu"], out["p_gpu"], out["c_cpu"], out["c_gpu"]
    seen, ev_cost, ev_correct = C.c_int32(out["seen"]), C.c_double(out["ev_cost"]), C.c_double(out["ev_correct"])
    assert seen.value == epochs, "the epoch callback must fire once per epoch"
    assert np.abs(p_cpu).max() > 0 and not np.array_equal(p_cpu, p_gpu * 0)
    H.assert_close("parameters: GpuDC binding vs the reference's own learn", p_gpu, p_cpu, real, rtol_mult=20)
    # the two nets' parameters agree within the bar, not bit for bit: an item whose two best outputs are within that bar of each other may
    # flip (identical classes on IDENTICAL parameters are asserted in tests/test_net_gpu.py and tests/test_reference_gpu.py)
    assert np.mean(c_cpu == c_gpu) >= 0.97, f"the reference's inference on the two trained nets disagrees on {int(np.sum(c_cpu != c_gpu))} of {n_items} items"
    assert 0.0 <= ev_correct.value <= 1.0 and np.isfinite(ev_cost.value)
    # (the device's own inference inside evaluate vs the reference's inference on the same parameters: a near-tie may flip one item)
    assert abs(ev_correct.value - float(np.mean(c_gpu == np.argmax(y, axis=1)))) <= 2.0 / n_items + 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("name,n_items", [("mini_lenet", 9), ("fixture", 5), ("avg", 6), ("lenet", 12)])
def test_reference_containers_from_device_buffers(name, n_items, real):
    """the binding hands back the reference's OWN container types - std::vector<LayerGradient<CpuDC>> (the
    gradient sum of a mini-batch) and std::vector<LayerData<CpuDC>> (Input / Trace.Derivatives / Trace.IndexMapping of an item) - built
    from the device buffers; compared, container by container, with what the reference's Net computes on the host; index maps exactly."""
    specs = NETS[name]
    dt = ob.dtype_of(real)
    in_n = specs[0].in_c * specs[0].in_h * specs[0].in_w
    out_shape = ob.out_shape(specs[-1])
    out_n = out_shape[0] * out_shape[1] * out_shape[2]
    rng = np.random.default_rng(9)
    x = rng.uniform(0, 1, (n_items, in_n)).astype(dt)
    cost = ob.SQUARED_ERROR if specs[-1].activation in (ob.LINEAR, ob.TANH) else ob.CROSS_ENTROPY
    if cost == ob.CROSS_ENTROPY:
        y = np.zeros((n_items, out_n), dt); y[np.arange(n_items), rng.integers(0, out_n, n_items)] = 1
    else:
        y = rng.uniform(0.05, 0.95, (n_items, out_n)).astype(dt)
    out = in_fresh_process(_container_case, real, name, x, y, cost)
    assert out["rc"] == 0, out["error"]
    eg, et, mism = C.c_double(out["eg"]), C.c_double(out["et"]), C.c_int64(out["mism"])
    bar = 2e-5 if real == "f32" else 1e-11
    assert 0 <= eg.value < bar, f"gradient containers: max relative error {eg.value:.3e}"
    assert 0 <= et.value < bar, f"trace containers: max relative error {et.value:.3e}"
    assert mism.value == 0, f"{mism.value} arg-max indices differ"


def _per_sample_case(real, name, x, y, cost):
    lib = load(real)
    vp = C.c_void_p
    lib.rad_check_per_sample.argtypes = [C.POINTER(ob.LayerSpec), C.c_int, C.c_uint, vp, vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    specs = NETS[name]
    arr = (ob.LayerSpec * len(specs))(*specs)
    err, mism = C.c_double(-1), C.c_int64(-1)
    rc = lib.rad_check_per_sample(arr, len(specs), 7, x.ctypes.data, y.ctypes.data, cost, C.byref(err), C.byref(mism))
    return dict(rc=rc, error=lib.rad_last_error().decode(), err=err.value, mism=mism.value)


@pytest.mark.gpu
@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("name", ["mini_lenet", "fixture", "avg", "mini_mlp"])
def test_per_sample_layer_api(name, real):
    """GpuLayer::act / backpropagate / update on GpuDC containers (one item per call, the library's batched operators
    with batch = 1) against the reference's own layers on the host, layer by layer - outputs, Trace.Derivatives, Trace.IndexMapping (exactly),
    input gradients, layer gradients with the scaled accumulation, parameters after update - as the reference's own GPU tests do
    (DeepLearningCudaTest/NetLayersCudaTest.cpp:77-172)."""
    specs = NETS[name]
    dt = ob.dtype_of(real)
    in_n = specs[0].in_c * specs[0].in_h * specs[0].in_w
    out_shape = ob.out_shape(specs[-1])
    out_n = out_shape[0] * out_shape[1] * out_shape[2]
    rng = np.random.default_rng(13)
    x = rng.uniform(0, 1, (2, in_n)).astype(dt)
    cost = ob.SQUARED_ERROR if specs[-1].activation in (ob.LINEAR, ob.TANH) else ob.CROSS_ENTROPY
    if cost == ob.CROSS_ENTROPY:
        y = np.zeros((2, out_n), dt); y[np.arange(2), rng.integers(0, out_n, 2)] = 1
    else:
        y = rng.uniform(0.05, 0.95, (2, out_n)).astype(dt)
    out = in_fresh_process(_per_sample_case, real, name, x, y, cost)
    assert out["rc"] == 0, out["error"]
    bar = 2e-5 if real == "f32" else 1e-11
    assert 0 <= out["err"] < bar, f"per-sample layer API: max relative error {out['err']:.3e}"
    assert out["mism"] == 0, f"{out['mism']} arg-max indices differ"

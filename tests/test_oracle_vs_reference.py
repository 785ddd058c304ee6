"""CPU: the C oracle against the UNMODIFIED reference compiled into oracle/_ref (live comparison).  Skipped where the
prebuilt reference libraries are absent; tests/test_golden_oracle.py covers the same ground from fixtures."""
import numpy as np
import pytest

from oracle import binding as ob
from tests import golden_cases as gc
from tests import helpers as H

pytestmark = pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built (needs /root/reference)")

NETS = {"mini_lenet": gc.MINI_LENET, "fixture": gc.FIXTURE_NET, "avg": gc.AVG_NET, "mini_mlp": gc.MINI_MLP,
        "lenet": ob.lenet_specs(), "cifar": ob.cifar_specs()}


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("name", list(NETS))
def test_single_sample_traces(name, real):
    """Layer outputs, derivatives, index maps, gradients with and without gradient_scale_factor, with the
    reference's OWN random initialisation (libstdc++ mt19937 + distributions, NLayer.inl:27-39, CLayer.inl:42-48)."""
    specs = NETS[name]
    ref = ob.RefNet(specs, real, seed=3)
    orc = ob.OracleNet(specs, real)
    orc.set_params(ref.get_params())
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, ref.in_sizes[0]).astype(ref.dtype)
    cost = ob.SQUARED_ERROR if specs[-1].activation in (ob.LINEAR, ob.TANH) else ob.CROSS_ENTROPY
    t = np.zeros(ref.out_sizes[-1], ref.dtype); t[1] = 1
    g0 = rng.normal(size=ref.param_count).astype(ref.dtype)
    for scale in (0.0, 0.37):
        a = ref.trace(x, t, cost, scale, g0)
        b = orc.trace(x, t, cost, scale, g0)
        for k in ("acts", "derivs", "grads", "in_grads", "cost_grad"):
            H.assert_close(f"{name} {k} (scale {scale})", b[k], a[k], real, rtol_mult=5)
        assert np.array_equal(a["index_map"], b["index_map"])


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_learn_with_thread_chunks(real, threads):
    """Net::learn with `threads` host threads: per-thread sequential accumulation + two-level tree sum
    (Net.inl:238-300) restated sequentially must reproduce the reference's parameters and predicted classes."""
    specs = gc.MINI_LENET
    n = 100
    ref = ob.RefNet(specs, real, seed=0)
    orc = ob.OracleNet(specs, real)
    orc.set_params(ref.get_params())
    x = ob.ref_uniform_items(real, n, (1, 12, 12))
    y = ob.one_hot_labels(real, n)
    ref.learn(x, y, 16, 2, 0.05, ob.CROSS_ENTROPY, 0.1, parallelism=threads, shuffle_seed=5)
    orc.learn(x, y, 16, 2, 0.05, ob.CROSS_ENTROPY, 0.1, threads, ob.ref_shuffle_sequence(5, n, 2))
    H.assert_close("parameters", orc.get_params(), ref.get_params(), real, rtol_mult=20)
    assert np.array_equal(orc.eval(x), ref.eval(x))


def test_cost_values_and_saturated_cross_entropy():
    """Cost values and derivatives incl. the nan_to_num branches (Functions.h:65-78): x = 0 / 1 against one-hot labels."""
    for real in ("f32", "f64"):
        ref = ob.RefNet([ob.full((1, 1, 3), 6, ob.LINEAR)], real, seed=1)
        orc = ob.OracleNet([ob.full((1, 1, 3), 6, ob.LINEAR)], real)
        dt = ref.dtype
        # a LINEAR layer with zero weights and the wanted outputs as biases produces exactly those outputs
        outs = np.array([0.0, 1.0, 0.3, 1.0, 0.0, 0.7], dt)
        refs = np.array([0.0, 1.0, 0.4, 0.0, 1.0, 1.0], dt)
        p = np.zeros(ref.param_count, dt); p[:6] = outs
        ref.set_params(p); orc.set_params(p)
        for cost in (ob.SQUARED_ERROR, ob.CROSS_ENTROPY, ob.COST_LINEAR):
            a = ref.trace(np.zeros(3, dt), refs, cost)
            b = orc.trace(np.zeros(3, dt), refs, cost)
            assert np.allclose(a["cost_grad"], b["cost_grad"], rtol=1e-6, atol=0), (cost, a["cost_grad"], b["cost_grad"])
            assert np.isclose(ref.cost(outs, refs, cost), orc.cost(outs, refs, cost), rtol=1e-6)
        a = ref.trace(np.zeros(3, dt), refs, ob.CROSS_ENTROPY)["cost_grad"]
        assert a[3] == 0 and a[4] == 0, "saturated cross-entropy branches must have derivative 0"


def test_shuffle_sequence_is_the_reference_permutation():
    """dlb_shuffle_sequence (product, host-only) == the permutations Net::learn draws (same libstdc++ calls)."""
    import deeplearning_b200 as dl
    assert np.array_equal(dl.shuffle_sequence(5, 100, 3), ob.ref_shuffle_sequence(5, 100, 3))


DROPOUT_KEEP = [0.8, 1.0, 0.5, 0.9]   # mini-LeNet: masks on the image, the first NLayer's input and the head's input


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("threads", [1, 3])
def test_learn_with_dropout(real, threads):
    """ALayer.inl:25-55 / Net.inl:126,228-230,274-275: per-mini-batch keep masks on the layer inputs and input gradients.
    The reference draws the masks itself (reset_random_generator(seed) seeds ALayer's generator too); the oracle is fed the
    same masks, produced by the reference's own Vector::fill_with_random_selection_map from that seed."""
    specs = gc.MINI_LENET
    n, batch, epochs, seed = 50, 8, 2, 11
    ref = ob.RefNet(specs, real, seed=3, keep_rates=DROPOUT_KEEP)
    orc = ob.OracleNet(specs, real, keep_rates=DROPOUT_KEEP)
    orc.set_params(ref.get_params())
    x = ob.ref_uniform_items(real, n, (1, 12, 12))
    y = ob.one_hot_labels(real, n)
    ref.learn(x, y, batch, epochs, 0.05, ob.CROSS_ENTROPY, 0.1, parallelism=threads, shuffle_seed=seed)
    masks = ob.ref_dropout_masks(real, seed, orc.in_sizes, DROPOUT_KEEP, ob.count_batches(n, batch) * epochs)
    assert masks.shape[1] == orc.mask_size == 144 + 100 + 16
    assert np.all((masks == 0) | (masks == 1)) and np.all(masks[:, :144].sum(axis=1) == int(np.float32(144) * np.float32(0.8)))
    orc.learn(x, y, batch, epochs, 0.05, ob.CROSS_ENTROPY, 0.1, threads, ob.ref_shuffle_sequence(seed, n, epochs), masks)
    H.assert_close("parameters after learn with dropout", orc.get_params(), ref.get_params(), real, rtol_mult=20)
    # inference scales the inputs of the dropping layers by keep_rate (Net.inl:143)
    ca, oa = ref.eval(x[:9], True)
    cb, ob_ = orc.eval(x[:9], True)
    H.assert_close("inference outputs with keep rates", ob_, oa, real, rtol_mult=5)
    assert np.array_equal(ca, cb)
    # and dropout matters: the same run without masks ends elsewhere
    plain = ob.OracleNet(specs, real)
    plain.set_params(ob.RefNet(specs, real, seed=3).get_params())
    plain.learn(x, y, batch, epochs, 0.05, ob.CROSS_ENTROPY, 0.1, threads, ob.ref_shuffle_sequence(seed, n, epochs))
    assert np.abs(plain.get_params() - orc.get_params()).max() > 1e-3


@pytest.mark.parametrize("real", ["f32", "f64"])
def test_dropout_masks_are_the_reference_masks(real):
    """dlb_dropout_masks (product, host-only) == the reference's mask stream, bit for bit (same libstdc++ calls)."""
    import deeplearning_b200 as dl
    sizes, keep = [144, 400, 100, 16], [0.8, 1.0, 0.5, 0.37]
    a = dl.dropout_masks(7, sizes, keep, 5, real)
    b = ob.ref_dropout_masks(real, 7, sizes, keep, 5)
    assert a.shape == b.shape == (5, 144 + 100 + 16) and np.array_equal(a, b)

"""Definitions of the golden fixtures (tests/golden/*.npz, written by tools/make_golden.py from the reference itself).

Inputs and initial parameters are regenerated from numpy seeds; the .npz files hold only what the reference computed.
"""
import numpy as np

from oracle import binding as ob

STRIDE = 61   # sampling stride for the vectors of the full-size LeNet fixture

MINI_LENET = [ob.conv((1, 12, 12), (3, 3), 4, ob.RELU), ob.pool((4, 10, 10), (2, 2), ob.POOL_MAX),
              ob.full((4, 5, 5), 16, ob.RELU), ob.full((1, 1, 16), 10, ob.SOFTMAX)]
# the reference's own layer-test fixtures (Shared/NetLayersTestUtils.h:34-65) chained into one net
FIXTURE_NET = [ob.conv((5, 13, 17), (3, 3), 7, ob.SIGMOID, pad=(3, 6), stride=(2, 2)), ob.pool((7, 9, 14), (3, 4), ob.POOL_MIN),
               ob.full((7, 3, 3), 23, ob.TANH), ob.full((1, 1, 23), 5, ob.LINEAR)]
AVG_NET = [ob.conv((2, 9, 8), (2, 3), 3, ob.TANH, pad=(1, 0), stride=(1, 2)), ob.pool((3, 10, 3), (3, 2), ob.POOL_AVERAGE),
           ob.conv((3, 3, 1), (2, 1), 5, ob.RELU, pad=(1, 1)), ob.full((5, 4, 3), 4, ob.SOFTMAX)]
MINI_MLP = [ob.full((1, 1, 20), 15, ob.SIGMOID), ob.full((1, 1, 15), 10, ob.SIGMOID)]

CASES = [
    dict(name="mini_lenet", specs=MINI_LENET, cost=ob.CROSS_ENTROPY, reals=("f32", "f64"), param_seed=3, data_seed=5, n_items=40,
         one_hot=True, n_trace=4, store_traces=True, learn=dict(batch=8, epochs=2, lr=0.05, lam=0.1, seed=5)),
    dict(name="fixture", specs=FIXTURE_NET, cost=ob.SQUARED_ERROR, reals=("f32", "f64"), param_seed=4, data_seed=6, n_items=23,
         one_hot=False, n_trace=3, store_traces=True, learn=dict(batch=5, epochs=2, lr=0.02, lam=1.5, seed=7)),
    dict(name="avg", specs=AVG_NET, cost=ob.CROSS_ENTROPY, reals=("f32", "f64"), param_seed=8, data_seed=9, n_items=24,
         one_hot=True, n_trace=3, store_traces=True, learn=dict(batch=10, epochs=2, lr=0.05, lam=0.0, seed=2)),
    dict(name="mini_mlp", specs=MINI_MLP, cost=ob.CROSS_ENTROPY, reals=("f32", "f64"), param_seed=10, data_seed=11, n_items=64,
         one_hot=True, n_trace=4, store_traces=True, learn=dict(batch=10, epochs=3, lr=0.5, lam=0.1, seed=1)),
    # BASELINE.json configs[1] at full size: LeNet, mini-batch 256, two steps
    dict(name="lenet", specs=ob.lenet_specs(), cost=ob.CROSS_ENTROPY, reals=("f32", "f64"), param_seed=21, data_seed=22, n_items=512,
         one_hot=True, n_trace=8, store_traces=False, learn=dict(batch=256, epochs=1, lr=0.05, lam=0.5, seed=3)),
]


def case_data(case, in_n, out_n, real):
    rng = np.random.default_rng(case["data_seed"])
    dtype = ob.dtype_of(real)
    n = case["n_items"]
    x = rng.uniform(0.0, 1.0, (n, in_n)).astype(dtype)
    if case["one_hot"]:
        y = np.zeros((n, out_n), dtype)
        y[np.arange(n), rng.integers(0, out_n, n)] = 1
    else:
        y = rng.uniform(0.05, 0.95, (n, out_n)).astype(dtype)
    return x, y


def load(case, real):
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"{case['name']}_{real}.npz")
    return np.load(path)

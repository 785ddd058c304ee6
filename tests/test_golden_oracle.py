"""CPU: the C oracle (oracle/dl_oracle.c) against the golden fixtures the reference itself produced
(tests/golden/*.npz, tools/make_golden.py).  This is what pins the oracle where /root/reference is absent."""
import numpy as np
import pytest

from oracle import binding as ob
from tests import golden_cases as gc
from tests import helpers as H


def _setup(case, real):
    orc = ob.OracleNet(case["specs"], real)
    orc.set_params(H.synthetic_params(orc.param_count, case["specs"], real, seed=case["param_seed"]))
    x, y = gc.case_data(case, orc.in_sizes[0], orc.out_sizes[-1], real)
    return orc, x, y, gc.load(case, real)


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("case", gc.CASES, ids=[c["name"] for c in gc.CASES])
def test_oracle_traces_match_reference_fixture(case, real):
    orc, x, y, gold = _setup(case, real)
    nt = case["n_trace"]
    got = H.oracle_batch_trace(orc, x[:nt], y[:nt], case["cost"])
    if case["store_traces"]:
        for k in ("acts", "derivs", "in_grads", "cost_grad", "grads"):
            H.assert_close(f"{case['name']} {k}", got[k], gold["trace_" + k], real, rtol_mult=5)
        assert np.array_equal(got["index_map"], gold["trace_index_map"])
    else:
        H.assert_close("gradient (strided)", got["grads"][::gc.STRIDE], gold["trace_grads_strided"], real, rtol_mult=5)
        assert np.array_equal(got["index_map"], gold["trace_index_map"])
        H.assert_close("net output", got["acts"][:, -orc.out_sizes[-1]:], gold["trace_out"], real)


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("case", gc.CASES, ids=[c["name"] for c in gc.CASES])
def test_oracle_learn_matches_reference_fixture(case, real):
    orc, x, y, gold = _setup(case, real)
    ln = case["learn"]
    orc.learn(x, y, ln["batch"], ln["epochs"], ln["lr"], case["cost"], ln["lam"], threads=1, perms=gold["learn_perms"])
    after = orc.get_params()
    want = gold["learn_params"]
    H.assert_close("parameters after learn", after if case["store_traces"] else after[::gc.STRIDE], want, real, rtol_mult=20)
    s = np.array([after.astype(np.float64).sum(), (after.astype(np.float64) ** 2).sum()])
    assert np.allclose(s, gold["learn_param_sum"], rtol=1e-5 if real == "f32" else 1e-11)
    cls, outs = orc.eval(x, want_outputs=True)
    assert np.array_equal(cls, gold["learn_classes"]), "predicted classes differ from the reference's"
    H.assert_close("outputs after learn", outs[:16], gold["learn_outputs"], real, rtol_mult=20)

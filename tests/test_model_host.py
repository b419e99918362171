"""Host builders of pi / transition / emission table against the dumps of the
unmodified reference in the golden vectors.  No GPU needed."""
import numpy as np
import pytest

from helpers import GOLDEN_NAMES, Golden
from smcpp_b200 import capi


def _f(x):
    return float(np.asarray(x).reshape(-1)[0])


def _build(g, keys=None, pol_err=None, theta=None):
    i = g.inp
    return capi.host_model_inputs(i["hidden_states"], i["model_a"], i["model_s"], _f(i["theta"]) if theta is None else theta,
                                  _f(i["rho"]), _f(i["alpha"]), _f(i["pol_err"]) if pol_err is None else pol_err, i["sfs"],
                                  i["n"], i["na"], g.ref["keys"] if keys is None else keys)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_pi_transition_emission_match_reference(name):
    g = Golden(name)
    out = _build(g)
    r = g.ref
    # the oracle build runs compute_expms in 113-bit arithmetic, as we do: the chain reproduces bit for bit;
    # against the reference's real 256-bit MPFR the difference is <= 3e-13 relative (SURVEY probe P5)
    assert np.abs(out["pi"] - r["pi"]).max() <= 1e-15
    assert np.abs(out["T"] / r["T"] - 1).max() <= 1e-13
    assert np.abs(out["E"] / r["E"] - 1).max() <= 1e-13
    ok = ~np.isnan(r["eta_avg_coal_times"])
    assert np.abs(out["avg_coal_times"][ok] / r["eta_avg_coal_times"][ok] - 1).max() <= 1e-14


def test_transition_rows_sum_like_the_reference():
    g = Golden("c2_1500")
    T = _build(g)["T"]
    assert np.abs(T.sum(1) - 1).max() == pytest.approx(1e-5 / 33, rel=1e-6)   # 1 - 1e-5/(M+1), not 1
    assert (T > 0).all()


def test_emission_special_keys():
    g = Golden("c1_2k")
    out = _build(g)
    keys = [tuple(k) for k in g.ref["keys"]]
    assert np.all(out["E"][keys.index((-1, 0, 0))] == 1.0)                      # all missing -> 1
    e0, e1 = out["E"][keys.index((0, 0, 0))], out["E"][keys.index((1, 0, 0))]
    assert np.allclose(e0 + e1, 1.0, atol=1e-15) and np.allclose(out["E"][keys.index((2, 0, 0))], e0)   # parity of a


def test_emission_errors_follow_the_reference():
    g = Golden("c1_2k")
    with pytest.raises(RuntimeError, match="theta <= 0"):
        _build(g, theta=0.0)
    # ("probability vector not in [0, 1]" is unreachable from finite inputs: incorporate_theta renormalises and floors)


def test_polarization_error_folds_keys():
    g = Golden("c1_2k")
    a = _build(g, pol_err=0.0)["E"]
    b = _build(g, pol_err=0.5)["E"]
    keys = [tuple(k) for k in g.ref["keys"]]
    k1, k2 = keys.index((0, 1, 4)), keys.index((2, 3, 4))                      # folded partners: a -> 2-a, b -> nb-b
    assert not np.allclose(a[k1], a[k2])
    assert np.allclose(b[k1], b[k2], rtol=1e-12)

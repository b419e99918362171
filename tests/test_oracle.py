"""The oracle port (oracle/hmm_port.c) against the golden vectors produced by the unmodified reference,
and -- where oracle/_ref is built -- against the reference run live."""
import numpy as np
import pytest

from helpers import GOLDEN_NAMES, Golden, relmax
from oracle import port, refrun
from smcpp_b200 import synth


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_port_matches_golden(name):
    g = Golden(name)
    for c, obs in enumerate(g.contigs):
        out = port.hmm_estep(obs, g.ref, want_alpha=True)
        ref_ll = g.ref["ll"][c]
        assert abs(out["ll"] - ref_ll) <= 1e-12 * abs(ref_ll)
        a_ref = g.ref[f"alpha_hat_{c}"]
        # the float forward pass is reproduced bit for bit (the port mirrors Eigen's summation orders,
        # including the alignment-dependent one of sum() when M is not a multiple of 4)
        assert np.array_equal(out["alpha_hat"], a_ref)
        assert out["ll"] == ref_ll
        for k in ("xisum", "gamma0", "gamma_sums"):
            assert relmax(out[k], g.ref[k][c]) < 1e-12, k
        assert np.array_equal(out["key_present"], g.ref["key_present"][c])


def test_span_table_matches_reference_formula():
    g = Golden("c2_1500")
    d = g.ref["eig_dscaled"][0]
    for span in (2, 7, 500, 50000):
        q = port.span_table(d, span)
        a, b = 3, 11
        d1, d2 = max(d[a], d[b], key=abs), min(d[a], d[b], key=abs)
        expect = np.exp(span * np.log(d1) + np.log1p(-(d2 / d1) ** span)) / (d1 - d2)
        assert q[a, b] == pytest.approx(expect, rel=1e-13)
        assert q[a, a] == pytest.approx(span * d[a] ** (span - 1), rel=1e-13)
        assert np.allclose(q, q.T, rtol=0, atol=0)


def test_port_raises_reference_span_error():
    g = Golden("c1_2k")
    ref = dict(g.ref)
    ref["eig_key_idx"] = ref["eig_key_idx"][:0]          # no eigensystems -> span > 1 rows hit hmm.cpp:132-133
    with pytest.raises(RuntimeError, match="span"):
        port.hmm_estep(g.contigs[0][:50], ref)


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref/ref_harness not built")
def test_port_matches_live_reference():
    w = synth.make_workload("live", 2, 900, 16, 5, seed0=4242)
    ref = refrun.run(w, dump_alpha=True)
    for c, obs in enumerate(w.contigs):
        out = port.hmm_estep(obs, ref, want_alpha=True)
        assert np.array_equal(out["alpha_hat"], ref[f"alpha_hat_{c}"])
        assert out["ll"] == ref["ll"][c]
        assert relmax(out["xisum"], ref["xisum"][c]) < 1e-12


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref/ref_harness not built")
def test_reference_transition_rows_sum_like_the_reference_says():
    # rows of T sum to 1 - 1e-5/(M+1), not 1
    w = synth.make_workload("rows", 1, 64, 16, 4)
    ref = refrun.run(w)
    assert np.abs(ref["T"].sum(1) - 1).max() == pytest.approx(1e-5 / 17, rel=1e-6)

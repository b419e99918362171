"""The drop-in, compiled and run: oracle/_ref/ref_harness_b200 is the reference's own InferenceManager (its unmodified
translation units) with InferenceManager::Estep taken from integration/_b200_estep.cpp, i.e. routed through the C ABI
into libsmcpp_b200.so.  The same bundle goes through it and through the unmodified ref_harness; everything the Python
surface of the reference can see afterwards must agree: loglik(), Q(), getXisums(), getGammaSums() (key sets and
values), gamma.col(0), and with saveGamma the full posterior.

Reference seam: src/inference_manager.cpp:108-150, include/hmm.h:11,33-37, smcpp/_smcpp.pyx:185-191.
"""
import numpy as np
import pytest

from helpers import LL_RTOL, STAT_RTOL, relmax
from oracle import refrun
from smcpp_b200 import synth

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (refrun.available() and refrun.available_b200()),
                                 reason="oracle/_ref/ref_harness{,_b200} did not travel to this box")]


def both(w, save_gamma=False):
    a = refrun.start(w, threads=len(w.contigs), save_gamma=save_gamma)
    b = refrun.run(w, threads=1, save_gamma=save_gamma, harness=refrun.HARNESS_B200, timeout=600)
    return a.result(timeout=900), b


@pytest.mark.parametrize("cfg,scale", [("C1", 1.0), ("C2", 0.03), ("C4", 0.03), ("C5-64", 0.01)])
def test_reference_inference_manager_with_the_library_estep(cfg, scale):
    w = synth.config(cfg, scale)
    ref, got = both(w)
    C = len(w.contigs)
    # loglik(): per contig (src/inference_manager.cpp:174-177)
    assert np.all(np.abs(got["ll"] - ref["ll"]) <= LL_RTOL * np.abs(ref["ll"]))
    # getXisums(), gamma.col(0), getGammaSums(): values and the per-contig key sets
    for k in ("xisum", "gamma0", "gamma_sums"):
        for c in range(C):
            assert relmax(got[k][c], ref[k][c]) <= STAT_RTOL, (k, c)
    assert np.array_equal(got["key_present"], ref["key_present"])
    assert np.array_equal(got["keys"], ref["keys"])
    # Q(): HMM::Q of the REFERENCE evaluated on the members the binding filled (src/hmm.cpp:155-193)
    assert np.allclose(got["Q"], ref["Q"], rtol=1e-9, atol=0), (got["Q"], ref["Q"])
    # the binding never runs TransitionBundle::update(T, true): no host eigensystems, no span tables
    assert got["eig_key_idx"].shape[0] == 0 and ref["eig_key_idx"].shape[0] > 0
    # identical inputs on both sides (do_dirty_work is the reference's own code in both binaries)
    for k in ("pi", "T", "E"):
        assert np.array_equal(got[k], ref[k]), k


def test_save_gamma_through_the_binding():
    w = synth.config("C2", 0.002)
    ref, got = both(w, save_gamma=True)
    g_ref, g_got = ref["gamma_full_0"], got["gamma_full_0"]
    assert g_got.shape == g_ref.shape == (w.contigs[0].shape[0] + 1, w.M)
    span = np.concatenate([[1], w.contigs[0][:, 0]])[:, None]
    assert relmax(g_got / span, g_ref / span) <= 1e-6          # columns as distributions (float alpha_hat noise of one column)
    assert np.allclose(g_got[1:].sum(1), w.contigs[0][:, 0], rtol=1e-9)
    assert abs(got["ll"][0] - ref["ll"][0]) <= LL_RTOL * abs(ref["ll"][0])


def test_malformed_data_do_not_throw_out_of_estep():
    """A library failure inside Estep must not escape as an exception (the Cython declaration has no `except +`): the
    harness finishes, the log-likelihoods are NaN."""
    w = synth.config("C1", 0.05)
    # the reference's constructor rejects span <= 0 itself; a failure the constructor cannot see: no usable device
    got = refrun.run(w, threads=1, harness=refrun.HARNESS_B200, timeout=300, extra_env={"SMCPP_B200_DEVICE": "99"})
    assert np.isnan(got["ll"]).all()

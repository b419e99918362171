"""Irregular spectra of diag(e_key) Td^T: complex eigenvalue pairs and negative real eigenvalues.

The reference keeps only the REAL PARTS of a complex eigensystem (include/transition_bundle.h:19-24), so P_r Pinv_r != I and
its per-block formulas (src/hmm.cpp:113-127, with element-wise abs() and log()) are no longer equivalent to the O(M^2)
displacement form; a negative eigenvalue makes its span tables NaN (std::log of a negative number,
src/transition_bundle.cpp:46-50).  The library detects both and follows the reference literally for such keys
(k_stats_literal, sequential chains, NaN semantics of the backward step) -- these tests pin that against the compiled
reference, which is driven with an overridden transition matrix / emission table (its own model never produces such
spectra; HMM::Estep accepts any input).
"""
import numpy as np
import pytest

from helpers import LL_RTOL, STAT_RTOL
from oracle import port, refrun
from smcpp_b200 import capi, synth


def irregular_workload(kind, M=8, L=500, seed=3):
    rng = np.random.default_rng(seed)
    w = synth.make_workload("irregular-" + kind, 2, L, M, 4, seed0=4000 + seed)
    keys = np.unique(np.concatenate([c[:, 1:] for c in w.contigs]), axis=0)
    K = keys.shape[0]
    if kind == "complex":
        # a cyclic drift (i -> i+1) makes the chain strongly non-reversible: complex pairs in the spectrum
        T = 0.15 * rng.random((M, M)) + 0.1 * np.eye(M) + 1.2 * np.roll(np.eye(M), 1, axis=1)
        T /= T.sum(1, keepdims=True)
        E = 0.3 + 0.7 * rng.random((K, M))
    else:
        # an alternating two-cycle: eigenvalue close to -0.7
        T = np.full((M, M), 0.02)
        for i in range(M):
            T[i, i ^ 1] = 1.0
        T += 0.15 * np.eye(M)
        T /= T.sum(1, keepdims=True)
        E = np.ones((K, M)) * (0.5 + 0.5 * rng.random((K, 1)))
    T = (1 - 1e-5) * T + 1e-5 / (M + 1)
    w.overrides = {"override_T": T, "override_E": E}
    return w


def close(a, b, rtol):
    a, b = np.asarray(a), np.asarray(b)
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return False
    ok = ~np.isnan(b)
    if not ok.any():
        return True
    scale = np.abs(b[ok]).max()
    return bool(np.abs(a[ok] - b[ok]).max() <= rtol * scale)


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref/ref_harness is not built")
@pytest.mark.parametrize("kind", ["complex", "negative"])
def test_port_follows_the_reference_on_irregular_spectra(kind):
    """CPU: the oracle port is literal, so it must agree with the compiled reference here too (pins the checker)."""
    w = irregular_workload(kind)
    ref = refrun.run(w)
    d = ref["eig_dscaled"]
    assert ref["eig_cplx"].any() if kind == "complex" else (d < -0.1).any()
    for c, obs in enumerate(w.contigs):
        o = port.hmm_estep(obs, ref)
        assert close(o["ll"], ref["ll"][c], 1e-12)
        assert close(o["xisum"], ref["xisum"][c], 1e-9)
        assert close(o["gamma_sums"], ref["gamma_sums"][c], 1e-9)


@pytest.mark.gpu
@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref/ref_harness did not travel to this box")
@pytest.mark.parametrize("kind", ["complex", "negative"])
@pytest.mark.parametrize("M", [8, 34])
def test_device_follows_the_reference_on_irregular_spectra(kind, M):
    w = irregular_workload(kind, M=M, L=400 if M > 8 else 500)
    ref = refrun.run(w)
    ctx = capi.Context(0)
    ctx.set_contigs(w.contigs, w.npop, ref["keys"])
    out = ctx.estep(ref["pi"], ref["T"], ref["E"], ref)               # the reference's own (real-part) eigensystems
    st = ctx.stats()
    assert st["literal_keys"] >= 1
    assert st["n_chunks"] == len(w.contigs), "irregular spectra run the literal, sequential chain"
    for c in range(len(w.contigs)):
        assert close(out["ll"][c], ref["ll"][c], LL_RTOL), (out["ll"][c], ref["ll"][c])
        for k in ("xisum", "gamma0", "gamma_sums"):
            assert close(out[k][c], ref[k][c], STAT_RTOL), (k, c)
    if kind == "negative":
        assert np.isnan(ref["xisum"]).any(), "the reference's span tables are NaN for a negative eigenvalue"
    # the regular path is back as soon as the spectrum is regular again
    g = ref
    T2 = 0.5 * (g["T"] + g["T"].T)
    T2 /= T2.sum(1, keepdims=True)
    out2 = ctx.estep(ref["pi"], T2, ref["E"], None)
    assert np.isfinite(out2["ll"]).all()
    ctx.close()


@pytest.mark.gpu
def test_library_eigensystems_flag_complex_spectra():
    w = irregular_workload("complex")
    keys = np.unique(np.concatenate([c[:, 1:] for c in w.contigs]), axis=0)
    ctx = capi.Context(0)
    ctx.set_contigs(w.contigs, w.npop, keys)
    pi = np.full(w.M, 1.0 / w.M)
    out = ctx.estep(pi, w.overrides["override_T"], w.overrides["override_E"], None)
    assert ctx.stats()["literal_keys"] >= 1
    eig = ctx.eigensystems(w.overrides["override_T"], w.overrides["override_E"])
    assert eig["eig_cplx"].any()
    # same formulas on the library's own real-part eigensystems: the port is the checker
    ref = {"pi": pi, "T": w.overrides["override_T"], "E": w.overrides["override_E"], "keys": keys, **eig}
    for c, obs in enumerate(w.contigs):
        o = port.hmm_estep(obs, ref)
        assert close(out["ll"][c], o["ll"], LL_RTOL)
        assert close(out["xisum"][c], o["xisum"], STAT_RTOL)
        assert close(out["gamma_sums"][c], o["gamma_sums"], STAT_RTOL)
    ctx.close()

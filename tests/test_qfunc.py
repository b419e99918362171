"""The M-step objective on the device (smcpp_b200_q) against the reference's own Q().

ref_harness evaluates the REFERENCE's InferenceManager::Q() (src/inference_manager.cpp:116-126 -> HMM::Q, src/hmm.cpp:155-193)
right after its E-step and dumps the four terms next to the statistics they were computed from.  Loading exactly those
statistics into the device (set_statistics) isolates Q: same inputs, same summation order -> agreement to rounding.
"""
import numpy as np
import pytest

from helpers import GOLDEN_NAMES, Golden
from smcpp_b200 import capi
from smcpp_b200.inference import InferenceManager

pytestmark = pytest.mark.gpu


def stats_of(ref):
    return ref["xisum"], ref["gamma0"], ref["gamma_sums"]


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_q_from_the_references_statistics(name):
    g = Golden(name)
    ref = g.ref
    ctx = capi.Context(0)
    ctx.set_contigs(g.contigs, g.npop, ref["keys"])
    ctx.set_statistics(*stats_of(ref))
    q = ctx.q(ref["pi"], ref["T"], ref["E"])
    assert np.allclose(q, ref["Q"], rtol=1e-12, atol=0), (q, ref["Q"])
    ctx.close()


@pytest.mark.parametrize("name", ["c2_1500", "c4_twopop_1200", "ragged"])
def test_q_after_the_device_estep(name):
    g = Golden(name)
    ref = g.ref
    ctx = capi.Context(0)
    ctx.set_contigs(g.contigs, g.npop, ref["keys"])
    ctx.estep(ref["pi"], ref["T"], ref["E"], ref)
    q = ctx.q(ref["pi"], ref["T"], ref["E"])
    assert np.allclose(q, ref["Q"], rtol=1e-8, atol=0), (q, ref["Q"])      # statistics agree to ~1e-9
    ctx.close()


def test_q_gradient_is_the_contraction_with_dlog():
    g = Golden("c2_1500")
    ref = g.ref
    M, K = ref["pi"].shape[0], ref["E"].shape[0]
    rng = np.random.default_rng(5)
    D = 7
    # directions proportional to the values (the transition matrix has entries of 3e-7: a unit step would leave the domain of log)
    dpi, dT, dE = rng.standard_normal((D, M)) * ref["pi"], rng.standard_normal((D, M, M)) * ref["T"], rng.standard_normal((D, K, M)) * ref["E"]
    ctx = capi.Context(0)
    ctx.set_contigs(g.contigs, g.npop, ref["keys"])
    ctx.set_statistics(*stats_of(ref))
    q, dq = ctx.q(ref["pi"], ref["T"], ref["E"], dpi, dT, dE)
    assert np.allclose(q, ref["Q"], rtol=1e-12, atol=0)
    g0, xi, gs = ref["gamma0"].sum(0), ref["xisum"].sum(0), ref["gamma_sums"].sum(0)
    nb = ref["keys"][:, 2::3].sum(1)
    want = np.stack([np.einsum("dm,m->d", dpi / ref["pi"], g0),
                     np.einsum("dkm,km->d", (dE / ref["E"])[:, nb == 0], gs[nb == 0]),
                     np.einsum("dkm,km->d", (dE / ref["E"])[:, nb > 0], gs[nb > 0]),
                     np.einsum("dij,ij->d", dT / ref["T"], xi)])
    assert np.allclose(dq, want, rtol=1e-10, atol=1e-9 * np.abs(want).max())
    # and it is the derivative of q along a direction: central difference on pi / T / E
    h = 1e-5
    p = 3
    up = ctx.q(ref["pi"] + h * dpi[p], ref["T"] + h * dT[p], ref["E"] + h * dE[p])
    dn = ctx.q(ref["pi"] - h * dpi[p], ref["T"] - h * dT[p], ref["E"] - h * dE[p])
    assert np.allclose((up - dn) / (2 * h), dq[:, p], rtol=1e-5, atol=1e-6 * np.abs(dq[:, p]).max())
    ctx.close()


def test_q_with_a_zero_emission_entry_is_minus_infinity_for_its_class():
    g = Golden("c1_2k")
    ref = g.ref
    E = ref["E"].copy()
    nb = ref["keys"][:, 2::3].sum(1)
    k = int(np.flatnonzero((nb > 0) & ref["key_present"][0].astype(bool))[0])
    E[k, 2] = 0.0
    ctx = capi.Context(0)
    ctx.set_contigs(g.contigs, g.npop, ref["keys"])
    ctx.set_statistics(*stats_of(ref))
    q = ctx.q(ref["pi"], ref["T"], E)
    assert q[2] == -np.inf and np.isfinite(q[[0, 3]]).all()
    with pytest.raises(RuntimeError, match="no statistics"):
        c2 = capi.Context(0)
        c2.set_contigs(g.contigs, g.npop, ref["keys"])
        c2.q(ref["pi"], ref["T"], ref["E"])
    ctx.close()


def test_inference_manager_q_before_and_after_the_first_estep():
    """Before the first E-step the reference's HMM constructor has pre-filled gamma_sums with span * pi (src/hmm.cpp:16-27)."""
    g = Golden("c2_1500")
    ref = g.ref
    im = InferenceManager(g.contigs, np.arange(ref["pi"].shape[0] + 1.0), g.npop, keys=ref["keys"])
    im.set_hmm_inputs(ref["pi"], ref["T"], ref["E"], ref)
    q0 = im.Q()
    span = g.contigs[0][:, 0].astype(np.float64)
    assert q0[0] == 0.0 and q0[3] == 0.0
    le = np.log(ref["E"])
    kid = {tuple(int(v) for v in k): i for i, k in enumerate(ref["keys"])}
    ids = np.array([kid[tuple(int(v) for v in r)] for r in g.contigs[0][:, 1:]])
    nb = ref["keys"][:, 2::3].sum(1)
    tot = sum(span[i] * float(le[ids[i]] @ ref["pi"]) for i in range(len(ids)))
    assert q0[1] + q0[2] == pytest.approx(tot, rel=1e-10)
    im.E_step()
    assert np.allclose(im.Q(), ref["Q"], rtol=1e-8, atol=0)
    im.close()

"""Host eigensolver of the library (smcpp_b200_host_eig / host_eigensystems): no GPU needed."""
import numpy as np
import pytest

from helpers import Golden
from smcpp_b200 import capi


@pytest.mark.parametrize("n", [1, 2, 3, 8, 17, 32, 64, 128])
def test_random_general_matrices(n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    P, Pi, dr, di = capi.host_eig(A)
    ev = np.linalg.eigvals(A)
    mine = dr + 1j * di
    # same spectrum (greedy matching)
    rem = list(ev)
    for z in mine:
        j = int(np.argmin([abs(z - r) for r in rem]))
        assert abs(z - rem[j]) < 1e-9 * max(1.0, abs(z))
        rem.pop(j)
    # complex pairs are conjugate, unit-norm eigenvector columns
    assert np.all(np.abs(np.sort(di) + np.sort(di)[::-1]) < 1e-12)


@pytest.mark.parametrize("name", ["c1_2k", "c2_1500", "m17_800", "m64_600", "ref_test_inference"])
def test_eigensystems_match_reference_through_the_operator(name):
    g = Golden(name)
    ref = g.ref
    mine = capi.host_eigensystems(ref["T"], ref["E"], ref["eig_key_idx"])
    M = g.M
    for e, k in enumerate(ref["eig_key_idx"]):
        A = np.diag(ref["E"][k]) @ ref["T"].T
        assert mine["eig_cplx"][e] == ref["eig_cplx"][e] == 0
        assert mine["eig_scale"][e] == pytest.approx(ref["eig_scale"][e], rel=1e-13)
        assert np.abs(np.sort(mine["eig_d"][e]) - np.sort(ref["eig_d"][e])).max() < 1e-13
        assert np.abs(mine["eig_P"][e] @ mine["eig_Pinv"][e] - np.eye(M)).max() < 1e-11
        for span in (1, 2, 50, 1000):
            op_m = mine["eig_P"][e] @ np.diag(mine["eig_dscaled"][e] ** span) @ mine["eig_Pinv"][e]
            op_r = ref["eig_P"][e] @ np.diag(ref["eig_dscaled"][e] ** span) @ ref["eig_Pinv"][e]
            assert np.abs(op_m - op_r).max() < 1e-11
        assert np.abs(mine["eig_P"][e] @ np.diag(mine["eig_d"][e]) @ mine["eig_Pinv"][e] - A).max() < 1e-12
        # unit 2-norm columns, like Eigen's EigenSolver::eigenvectors()
        assert np.allclose(np.linalg.norm(mine["eig_P"][e], axis=0), 1.0, atol=1e-12)


def test_complex_spectrum_keeps_real_parts_like_the_reference():
    # rotation-like matrix: complex pair; the reference keeps Re(P), Re(P^-1) (transition_bundle.h:21-24)
    A = np.array([[0.9, -0.3, 0.0], [0.3, 0.9, 0.0], [0.1, 0.0, 0.5]])
    P, Pi, dr, di = capi.host_eig(A)
    w, V = np.linalg.eig(A)
    assert (di != 0).sum() == 2
    # P is determined up to per-column phase; compare the invariant P_c diag(d) P_c^-1 via numpy instead
    assert np.allclose(np.sort(dr), np.sort(w.real), atol=1e-13)

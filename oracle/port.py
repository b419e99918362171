"""ctypes front-end of the CPU restatement in hmm_port.c (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    so = os.path.join(_HERE, "libsmcb_oracle.so")
    src = os.path.join(_HERE, "hmm_port.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.smcb_oracle_hmm_estep.restype = ctypes.c_int
        _LIB.smcb_oracle_span_table.restype = None
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def key_ids(obs: np.ndarray, keys: np.ndarray) -> np.ndarray:
    """Index of each row's key in `keys` (K x 3P, lexicographic = the reference's std::map order,
    reference include/block_key.h:51-60)."""
    lut = {tuple(int(x) for x in k): i for i, k in enumerate(keys)}
    uniq, inv = np.unique(obs[:, 1:], axis=0, return_inverse=True)
    m = np.array([lut[tuple(int(x) for x in u)] for u in uniq], np.int32)
    return m[inv.reshape(-1)].astype(np.int32)


def span_table(d_scaled: np.ndarray, span: int) -> np.ndarray:
    M = d_scaled.shape[0]
    q = np.empty((M, M))
    lib().smcb_oracle_span_table(ctypes.c_int(M), ctypes.c_int(int(span)), _p(np.ascontiguousarray(d_scaled), ctypes.c_double),
                                 _p(q, ctypes.c_double))
    return q.T.copy()


def hmm_estep(obs: np.ndarray, ref: dict, save_gamma=False, want_alpha=False) -> dict:
    """Run the port on one contig.  `ref` holds pi, T, keys, E, eig_* exactly as ref_harness dumps them
    (or as the product computes them)."""
    obs = np.ascontiguousarray(obs, np.int32)
    L = obs.shape[0]
    pi = np.ascontiguousarray(ref["pi"], np.float64)
    M = pi.shape[0]
    T = np.ascontiguousarray(ref["T"], np.float64)
    E = np.ascontiguousarray(ref["E"], np.float64)
    K = E.shape[0]
    kid = key_ids(obs, ref["keys"])
    span = np.ascontiguousarray(obs[:, 0], np.int32)
    eig_of_key = np.full(K, -1, np.int32)
    for e, k in enumerate(ref["eig_key_idx"]):
        eig_of_key[int(k)] = e
    P = np.ascontiguousarray(ref["eig_P"], np.float64)
    Pi = np.ascontiguousarray(ref["eig_Pinv"], np.float64)
    d = np.ascontiguousarray(ref["eig_d"], np.float64)
    ds = np.ascontiguousarray(ref["eig_dscaled"], np.float64)
    sc = np.ascontiguousarray(ref["eig_scale"], np.float64)
    ll = np.zeros(1)
    xisum = np.zeros((M, M))
    gamma0 = np.zeros(M)
    gsum = np.zeros((K, M))
    present = np.zeros(K, np.uint8)
    gfull = np.zeros((L + 1, M)) if save_gamma else None
    ah = np.zeros((L + 1, M), np.float32) if want_alpha else None
    lc = np.zeros(L + 1) if want_alpha else None
    D, F, I, U = ctypes.c_double, ctypes.c_float, ctypes.c_int32, ctypes.c_uint8
    rc = lib().smcb_oracle_hmm_estep(
        ctypes.c_int(M), ctypes.c_long(L), ctypes.c_int(K), _p(span, I), _p(kid, I), _p(pi, D), _p(T, D), _p(E, D),
        _p(eig_of_key, I), _p(P, D), _p(Pi, D), _p(d, D), _p(ds, D), _p(sc, D), _p(ll, D), _p(xisum, D), _p(gamma0, D),
        _p(gsum, D), _p(present, U), _p(gfull, D), _p(ah, F), _p(lc, D))
    if rc:
        raise RuntimeError("span")   # the reference's std::runtime_error("span"), src/hmm.cpp:133
    out = {"ll": float(ll[0]), "xisum": xisum, "gamma0": gamma0, "gamma_sums": gsum, "key_present": present}
    if save_gamma:
        out["gamma_full"] = gfull
    if want_alpha:
        out["alpha_hat"] = ah
        out["log_c"] = lc
    return out

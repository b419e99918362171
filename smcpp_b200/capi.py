"""ctypes binding of libsmcpp_b200.so (the C ABI declared in include/smcpp_b200.h).

The library is the product: if it cannot be loaded, or no CUDA device is usable, every entry point
raises -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMCPP_B200_LIB") or os.path.join(_HERE, "libsmcpp_b200.so")   # (override: profiling builds, tools/)
_lib = None

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_u8p = ctypes.POINTER(ctypes.c_uint8)


class Stats(ctypes.Structure):
    _fields_ = [("n_chunks", ctypes.c_int32), ("chunk_blocks", ctypes.c_int32), ("burn_in_blocks", ctypes.c_int32),
                ("fwd_sweeps", ctypes.c_int32), ("bwd_sweeps", ctypes.c_int32), ("fwd_redone", ctypes.c_int32),
                ("bwd_redone", ctypes.c_int32), ("kernel_launches", ctypes.c_int32),
                ("ms_setup", ctypes.c_float), ("ms_forward", ctypes.c_float), ("ms_backward", ctypes.c_float),
                ("ms_stats", ctypes.c_float), ("ms_finalize", ctypes.c_float), ("ms_total", ctypes.c_float),
                ("fwd_max_mismatch", ctypes.c_double), ("bwd_max_mismatch", ctypes.c_double),
                ("ms_forward_only", ctypes.c_float), ("mma_rounds", ctypes.c_int32), ("mma_steps", ctypes.c_int32),
                ("literal_keys", ctypes.c_int32), ("restarts", ctypes.c_int32), ("converged", ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/smcpp_b200.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = [
    "smcpp_b200_abi_version", "smcpp_b200_create", "smcpp_b200_destroy", "smcpp_b200_last_error",
    "smcpp_b200_set_option", "smcpp_b200_set_contigs", "smcpp_b200_num_keys", "smcpp_b200_get_keys",
    "smcpp_b200_num_eig_keys", "smcpp_b200_get_eig_keys", "smcpp_b200_get_key_present", "smcpp_b200_total_blocks",
    "smcpp_b200_eigensystems", "smcpp_b200_host_eig", "smcpp_b200_host_eigensystems", "smcpp_b200_estep",
    "smcpp_b200_host_initial_distribution", "smcpp_b200_host_average_coal_times", "smcpp_b200_host_transition",
    "smcpp_b200_host_emission",
    "smcpp_b200_reduced_device_ptr", "smcpp_b200_copy_reduced_to_device", "smcpp_b200_estep_device", "smcpp_b200_fetch", "smcpp_b200_get_stats",
    "smcpp_b200_fp64_peak", "smcpp_b200_stream", "smcpp_b200_debug_alpha_hat", "smcpp_b200_set_save_gamma",
    "smcpp_b200_fetch_gamma", "smcpp_b200_q", "smcpp_b200_set_statistics",
    "smcpp_b200_multi_create", "smcpp_b200_multi_destroy", "smcpp_b200_multi_last_error", "smcpp_b200_multi_num_devices",
    "smcpp_b200_multi_context", "smcpp_b200_multi_set_contigs", "smcpp_b200_multi_num_keys", "smcpp_b200_multi_get_keys",
    "smcpp_b200_multi_get_shard", "smcpp_b200_multi_estep",
    "smcpp_b200_obs_create", "smcpp_b200_obs_destroy", "smcpp_b200_obs_last_error", "smcpp_b200_obs_upload", "smcpp_b200_obs_thin",
    "smcpp_b200_obs_bin", "smcpp_b200_obs_recode_monomorphic", "smcpp_b200_obs_compress", "smcpp_b200_obs_rows", "smcpp_b200_obs_last_ms",
    "smcpp_b200_obs_download", "smcpp_b200_obs_recode_nonseg", "smcpp_b200_obs_break_long_spans", "smcpp_b200_obs_piece_offsets",
    "smcpp_b200_obs_select_piece",
]


def build(force: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc")
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src))
    newest = max(newest, os.path.getmtime(os.path.join(_HERE, "..", "include", "smcpp_b200.h")))
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.check_call(["make", "-s", "-j", "8", "-C", src], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(smcpp_b200 has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.smcpp_b200_last_error.restype = ctypes.c_char_p
        L.smcpp_b200_last_error.argtypes = [ctypes.c_void_p]
        L.smcpp_b200_total_blocks.restype = ctypes.c_int64
        L.smcpp_b200_total_blocks.argtypes = [ctypes.c_void_p]
        L.smcpp_b200_destroy.restype = None
        L.smcpp_b200_destroy.argtypes = [ctypes.c_void_p]
        for name in ("smcpp_b200_num_keys", "smcpp_b200_num_eig_keys"):
            getattr(L, name).argtypes = [ctypes.c_void_p]
        L.smcpp_b200_obs_last_error.restype = ctypes.c_char_p
        L.smcpp_b200_obs_last_error.argtypes = [ctypes.c_void_p]
        L.smcpp_b200_obs_rows.restype = ctypes.c_int64
        L.smcpp_b200_obs_rows.argtypes = [ctypes.c_void_p]
        L.smcpp_b200_obs_last_ms.restype = ctypes.c_float
        L.smcpp_b200_obs_last_ms.argtypes = [ctypes.c_void_p]
        L.smcpp_b200_obs_destroy.restype = None
        L.smcpp_b200_obs_destroy.argtypes = [ctypes.c_void_p]
        L.smcpp_b200_multi_last_error.restype = ctypes.c_char_p
        L.smcpp_b200_multi_last_error.argtypes = [ctypes.c_void_p]
        L.smcpp_b200_multi_destroy.restype = None
        L.smcpp_b200_multi_destroy.argtypes = [ctypes.c_void_p]
        for name in ("smcpp_b200_multi_num_keys", "smcpp_b200_multi_num_devices"):
            getattr(L, name).argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


def ptr(a, t):
    if a is None:
        return None
    return a.ctypes.data_as(t)


def host_eig(A: np.ndarray):
    """Eigen-decomposition of a general real matrix by the library's host routine -> (P, Pinv, d_re, d_im)."""
    A = np.ascontiguousarray(A, np.float64)
    n = A.shape[0]
    P = np.empty((n, n)); Pi = np.empty((n, n)); dr = np.empty(n); di = np.empty(n)
    rc = lib().smcpp_b200_host_eig(ctypes.c_int(n), ptr(A, c_f64p), ptr(P, c_f64p), ptr(Pi, c_f64p), ptr(dr, c_f64p),
                                   ptr(di, c_f64p))
    if rc:
        raise RuntimeError("smcpp_b200_host_eig failed")
    return P, Pi, dr, di


def host_eigensystems(T: np.ndarray, E: np.ndarray, eig_key_idx: np.ndarray) -> dict:
    T = np.ascontiguousarray(T, np.float64); E = np.ascontiguousarray(E, np.float64)
    idx = np.ascontiguousarray(eig_key_idx, np.int32)
    M, K, ne = T.shape[0], E.shape[0], idx.shape[0]
    out = {"eig_key_idx": idx, "eig_P": np.empty((ne, M, M)), "eig_Pinv": np.empty((ne, M, M)), "eig_d": np.empty((ne, M)),
           "eig_dscaled": np.empty((ne, M)), "eig_scale": np.empty(ne), "eig_cplx": np.zeros(ne, np.int32)}
    rc = lib().smcpp_b200_host_eigensystems(ctypes.c_int(M), ctypes.c_int(K), ctypes.c_int(ne), ptr(idx, c_i32p),
                                            ptr(T, c_f64p), ptr(E, c_f64p), ptr(out["eig_P"], c_f64p),
                                            ptr(out["eig_Pinv"], c_f64p), ptr(out["eig_d"], c_f64p),
                                            ptr(out["eig_dscaled"], c_f64p), ptr(out["eig_scale"], c_f64p),
                                            ptr(out["eig_cplx"], c_i32p))
    if rc:
        raise RuntimeError("smcpp_b200_host_eigensystems failed")
    return out


def host_model_inputs(hidden_states, model_a, model_s, theta, rho, alpha, pol_err, sfs, n, na, keys) -> dict:
    """pi, transition and emission table of one E-step, built by the library's host routines (rows a3-a5 of
    SURVEY 8a): the value parts of what the reference's do_dirty_work() produces."""
    hs = np.ascontiguousarray(hidden_states, np.float64)
    a = np.ascontiguousarray(model_a, np.float64); s = np.ascontiguousarray(model_s, np.float64)
    sfs = np.ascontiguousarray(sfs, np.float64); keys = np.ascontiguousarray(keys, np.int32)
    n = np.ascontiguousarray(np.atleast_1d(n), np.int32); na = np.ascontiguousarray(np.atleast_1d(na), np.int32)
    M, P, K, npc = hs.shape[0] - 1, n.shape[0], keys.shape[0], a.shape[0]
    out = {"pi": np.empty(M), "T": np.empty((M, M)), "E": np.empty((K, M)), "avg_coal_times": np.empty(M), "keys": keys}
    L = lib()
    D = ctypes.c_double
    rc = L.smcpp_b200_host_initial_distribution(ctypes.c_int(M), ptr(hs, c_f64p), ctypes.c_int(npc), ptr(a, c_f64p), ptr(s, c_f64p),
                                                ptr(out["pi"], c_f64p))
    rc |= L.smcpp_b200_host_average_coal_times(ctypes.c_int(M), ptr(hs, c_f64p), ctypes.c_int(npc), ptr(a, c_f64p), ptr(s, c_f64p),
                                               ptr(out["avg_coal_times"], c_f64p))
    rc |= L.smcpp_b200_host_transition(ctypes.c_int(M), ptr(hs, c_f64p), ctypes.c_int(npc), ptr(a, c_f64p), ptr(s, c_f64p),
                                       D(rho), ptr(out["T"], c_f64p))
    if rc:
        raise RuntimeError("smcpp_b200 host model builders failed")
    err = ctypes.create_string_buffer(256)
    rc = L.smcpp_b200_host_emission(ctypes.c_int(P), ptr(n, c_i32p), ptr(na, c_i32p), ctypes.c_int(M), ptr(hs, c_f64p),
                                    ctypes.c_int(npc), ptr(a, c_f64p), ptr(s, c_f64p), D(theta), D(alpha), D(pol_err),
                                    ptr(sfs, c_f64p), ctypes.c_int(K), ptr(keys, c_i32p), ptr(out["E"], c_f64p), err,
                                    ctypes.c_int(256))
    if rc:
        raise RuntimeError(err.value.decode() or "smcpp_b200_host_emission failed")
    return out


class Context:
    """One GPU's E-step engine (thin, typed view of the C ABI)."""

    def __init__(self, device: int = 0):
        self._h = ctypes.c_void_p()
        rc = lib().smcpp_b200_create(ctypes.byref(self._h), ctypes.c_int(device))
        if rc:
            raise RuntimeError("smcpp_b200_create: " + lib().smcpp_b200_last_error(None).decode())
        self.C = 0
        self.K = 0
        self.M = 0
        # tuning knobs for experiments: SMCPP_B200_OPTS="target_warps=2368,burn_in_blocks=768"
        for item in os.environ.get("SMCPP_B200_OPTS", "").split(","):
            if "=" in item:
                k, v = item.split("=", 1)
                self.set_option(k.strip(), float(v))

    def close(self):
        if self._h and not getattr(self, "_borrowed", False):
            lib().smcpp_b200_destroy(self._h)
        self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc:
            raise RuntimeError(f"{what}: " + lib().smcpp_b200_last_error(self._h).decode())

    def set_option(self, name: str, value: float):
        self._check(lib().smcpp_b200_set_option(self._h, name.encode(), ctypes.c_double(value)), "set_option")

    def set_contigs(self, contigs, npop: int, keys: np.ndarray | None = None):
        arrs = [np.ascontiguousarray(c, np.int32) for c in contigs]
        W = 1 + 3 * npop
        for a in arrs:
            if a.ndim != 2 or a.shape[1] != W:
                raise ValueError(f"observations must be [L, {W}] int32")
        C = len(arrs)
        ptrs = (c_i32p * C)(*[ptr(a, c_i32p) for a in arrs])
        lens = np.asarray([a.shape[0] for a in arrs], np.int32)
        kp, nk = None, 0
        if keys is not None:
            keys = np.ascontiguousarray(keys, np.int32)
            kp, nk = ptr(keys, c_i32p), keys.shape[0]
        self._check(lib().smcpp_b200_set_contigs(self._h, ctypes.c_int(C), ptrs, ptr(lens, c_i32p), ctypes.c_int(npop),
                                                 kp, ctypes.c_int(nk)), "set_contigs")
        self.C, self.npop = C, npop
        self.K = lib().smcpp_b200_num_keys(self._h)
        self.lengths = lens

    @property
    def keys(self) -> np.ndarray:
        k = np.empty((self.K, 3 * self.npop), np.int32)
        self._check(lib().smcpp_b200_get_keys(self._h, ptr(k, c_i32p)), "get_keys")
        return k

    @property
    def eig_keys(self) -> np.ndarray:
        n = lib().smcpp_b200_num_eig_keys(self._h)
        k = np.empty(n, np.int32)
        if n:
            self._check(lib().smcpp_b200_get_eig_keys(self._h, ptr(k, c_i32p)), "get_eig_keys")
        return k

    @property
    def key_present(self) -> np.ndarray:
        p = np.empty((self.C, self.K), np.uint8)
        self._check(lib().smcpp_b200_get_key_present(self._h, ptr(p, c_u8p)), "get_key_present")
        return p

    @property
    def total_blocks(self) -> int:
        return int(lib().smcpp_b200_total_blocks(self._h))

    def _eig_args(self, eig):
        if eig is None:
            return 0, None, None, None, None, None, []
        sel = slice(None)
        if "eig_key_idx" in eig:
            # eigensystems are matched by key index: a context may hold a subset of the eigen keys
            pos = {int(k): i for i, k in enumerate(np.asarray(eig["eig_key_idx"]))}
            sel = [pos[int(k)] for k in self.eig_keys]
        keep = [np.ascontiguousarray(np.asarray(eig[k], np.float64)[sel]) for k in ("eig_P", "eig_Pinv", "eig_d", "eig_dscaled", "eig_scale")]
        return (keep[0].shape[0], ptr(keep[0], c_f64p), ptr(keep[1], c_f64p), ptr(keep[2], c_f64p), ptr(keep[3], c_f64p),
                ptr(keep[4], c_f64p), keep)

    def estep(self, pi, T, E, eig: dict | None = None) -> dict:
        """One E-step.  `eig` = dict with eig_P, eig_Pinv, eig_d, eig_dscaled, eig_scale in the order of
        self.eig_keys (e.g. the reference's own eigensystems), or None to let the library compute them."""
        pi = np.ascontiguousarray(pi, np.float64); T = np.ascontiguousarray(T, np.float64)
        E = np.ascontiguousarray(E, np.float64)
        M, K, C = pi.shape[0], self.K, self.C
        if E.shape != (K, M) or T.shape != (M, M):
            raise ValueError("shape mismatch: T must be [M,M], E must be [K,M]")
        ne, pP, pPi, pd, pds, psc, keep = self._eig_args(eig)
        out = {"ll": np.empty(C), "xisum": np.empty((C, M, M)), "gamma0": np.empty((C, M)),
               "gamma_sums": np.empty((C, K, M)), "reduced": np.empty(1 + M + M * M + K * M)}
        self._check(lib().smcpp_b200_estep(self._h, ctypes.c_int(M), ptr(pi, c_f64p), ptr(T, c_f64p), ptr(E, c_f64p),
                                           ctypes.c_int(ne), pP, pPi, pd, pds, psc, ptr(out["ll"], c_f64p),
                                           ptr(out["xisum"], c_f64p), ptr(out["gamma0"], c_f64p),
                                           ptr(out["gamma_sums"], c_f64p), ptr(out["reduced"], c_f64p)), "estep")
        self.M = M
        out["key_present"] = self.key_present
        return out

    def estep_device(self, pi, T, E, eig: dict | None = None, upload: bool = True):
        pi = np.ascontiguousarray(pi, np.float64); T = np.ascontiguousarray(T, np.float64)
        E = np.ascontiguousarray(E, np.float64)
        M = pi.shape[0]
        ne, pP, pPi, pd, pds, psc, keep = self._eig_args(eig)
        self._check(lib().smcpp_b200_estep_device(self._h, ctypes.c_int(M), ptr(pi, c_f64p), ptr(T, c_f64p),
                                                  ptr(E, c_f64p), ctypes.c_int(ne), pP, pPi, pd, pds, psc,
                                                  ctypes.c_int(1 if upload else 0)), "estep_device")
        self.M = M

    def fetch(self) -> dict:
        M, K, C = self.M, self.K, self.C
        out = {"ll": np.empty(C), "xisum": np.empty((C, M, M)), "gamma0": np.empty((C, M)),
               "gamma_sums": np.empty((C, K, M)), "reduced": np.empty(1 + M + M * M + K * M)}
        self._check(lib().smcpp_b200_fetch(self._h, ptr(out["ll"], c_f64p), ptr(out["xisum"], c_f64p),
                                           ptr(out["gamma0"], c_f64p), ptr(out["gamma_sums"], c_f64p),
                                           ptr(out["reduced"], c_f64p)), "fetch")
        return out

    def set_statistics(self, xisum, gamma0, gamma_sums):
        """Load per-contig statistics ([C,M,M], [C,M], [C,K,M]) instead of computing them (constructor pre-fill, checkpoint)."""
        xisum = np.ascontiguousarray(xisum, np.float64); gamma0 = np.ascontiguousarray(gamma0, np.float64)
        gamma_sums = np.ascontiguousarray(gamma_sums, np.float64)
        M = gamma0.shape[1]
        if xisum.shape != (self.C, M, M) or gamma_sums.shape != (self.C, self.K, M):
            raise ValueError("set_statistics: expected xisum[C,M,M], gamma0[C,M], gamma_sums[C,K,M]")
        self._check(lib().smcpp_b200_set_statistics(self._h, ctypes.c_int(M), ptr(xisum, c_f64p), ptr(gamma0, c_f64p),
                                                    ptr(gamma_sums, c_f64p)), "set_statistics")
        self.M = M

    def q(self, pi, T, E, dpi=None, dT=None, dE=None):
        """M-step objective from the resident statistics: q[4], and dq[4, D] when the derivative arrays dpi[D,M], dT[D,M,M],
        dE[D,K,M] are given (reference InferenceManager::Q on autodiff scalars)."""
        pi = np.ascontiguousarray(pi, np.float64); T = np.ascontiguousarray(T, np.float64); E = np.ascontiguousarray(E, np.float64)
        M = pi.shape[0]
        if T.shape != (M, M) or E.shape != (self.K, M):
            raise ValueError("q: T must be [M,M], E must be [K,M]")
        q = np.empty(4)
        D = 0
        dq = None
        if dpi is not None:
            dpi = np.ascontiguousarray(dpi, np.float64); dT = np.ascontiguousarray(dT, np.float64); dE = np.ascontiguousarray(dE, np.float64)
            D = dpi.shape[0]
            if dpi.shape != (D, M) or dT.shape != (D, M, M) or dE.shape != (D, self.K, M):
                raise ValueError("q: derivative arrays must be dpi[D,M], dT[D,M,M], dE[D,K,M]")
            dq = np.empty((4, D))
        self._check(lib().smcpp_b200_q(self._h, ctypes.c_int(M), ptr(pi, c_f64p), ptr(T, c_f64p), ptr(E, c_f64p), ctypes.c_int(D),
                                       ptr(dpi, c_f64p), ptr(dT, c_f64p), ptr(dE, c_f64p), ptr(q, c_f64p), ptr(dq, c_f64p)), "q")
        return q if dq is None else (q, dq)

    def reduced_device_ptr(self):
        p = ctypes.c_void_p(); n = ctypes.c_int64()
        self._check(lib().smcpp_b200_reduced_device_ptr(self._h, ctypes.byref(p), ctypes.byref(n)), "reduced_device_ptr")
        return p.value, n.value

    def copy_reduced_to_device(self, dst_ptr: int, count: int):
        self._check(lib().smcpp_b200_copy_reduced_to_device(self._h, ctypes.c_void_p(dst_ptr), ctypes.c_int64(count)),
                    "copy_reduced_to_device")

    def eigensystems(self, T, E) -> dict:
        return host_eigensystems(T, E, self.eig_keys)

    def stats(self) -> dict:
        s = Stats()
        self._check(lib().smcpp_b200_get_stats(self._h, ctypes.byref(s)), "get_stats")
        return s.as_dict()

    def fp64_peak_tflops(self) -> float:
        v = ctypes.c_double()
        self._check(lib().smcpp_b200_fp64_peak(self._h, ctypes.byref(v)), "fp64_peak")
        return v.value

    def stream(self) -> int:
        p = ctypes.c_void_p()
        self._check(lib().smcpp_b200_stream(self._h, ctypes.byref(p)), "stream")
        return p.value or 0

    def set_save_gamma(self, on, normalise: bool = False):
        """on: keep every column of the posterior; normalise: columns divided by their sums on the device (`smc++ posterior`)."""
        self._check(lib().smcpp_b200_set_save_gamma(self._h, ctypes.c_int((2 if normalise else 1) if on else 0)), "set_save_gamma")

    def fetch_gamma(self, contig: int) -> np.ndarray:
        """Posterior of one contig as [L+1, M] (the reference holds M x (L+1))."""
        L = int(self.lengths[contig])
        out = np.empty((L + 1, self.M))
        self._check(lib().smcpp_b200_fetch_gamma(self._h, ctypes.c_int(contig), ptr(out, c_f64p)), "fetch_gamma")
        return out

    def debug_alpha_hat(self, contig: int) -> np.ndarray:
        L = int(self.lengths[contig])
        out = np.empty((L + 1, self.M), np.float32)
        self._check(lib().smcpp_b200_debug_alpha_hat(self._h, ctypes.c_int(contig), ptr(out, c_f32p)), "debug_alpha_hat")
        return out


class MultiContext:
    """Several GPUs in one process (include/smcpp_b200.h: smcpp_b200_multi_*): contigs sharded over the devices, one
    in-library ncclAllReduce of the packed statistics per E-step."""

    def __init__(self, devices):
        devs = np.ascontiguousarray(list(devices), np.int32)
        self._h = ctypes.c_void_p()
        if lib().smcpp_b200_multi_create(ctypes.byref(self._h), ptr(devs, c_i32p), ctypes.c_int(len(devs))):
            raise RuntimeError("smcpp_b200_multi_create: " + lib().smcpp_b200_multi_last_error(None).decode())
        self.devices = [int(d) for d in devs]
        self.C = self.K = 0

    def _check(self, rc, what):
        if rc:
            raise RuntimeError(f"{what}: " + lib().smcpp_b200_multi_last_error(self._h).decode())

    def set_contigs(self, contigs, npop: int):
        arrs = [np.ascontiguousarray(c, np.int32) for c in contigs]
        C = len(arrs)
        ptrs = (c_i32p * C)(*[ptr(a, c_i32p) for a in arrs])
        lens = np.asarray([a.shape[0] for a in arrs], np.int32)
        self._check(lib().smcpp_b200_multi_set_contigs(self._h, ctypes.c_int(C), ptrs, ptr(lens, c_i32p), ctypes.c_int(npop)), "multi_set_contigs")
        self.C, self.npop = C, npop
        self.K = lib().smcpp_b200_multi_num_keys(self._h)

    @property
    def keys(self):
        k = np.empty((self.K, 3 * self.npop), np.int32)
        self._check(lib().smcpp_b200_multi_get_keys(self._h, ptr(k, c_i32p)), "multi_get_keys")
        return k

    def shard(self, i: int):
        n = ctypes.c_int32()
        self._check(lib().smcpp_b200_multi_get_shard(self._h, ctypes.c_int(i), None, ctypes.byref(n)), "multi_get_shard")
        out = np.empty(n.value, np.int32)
        self._check(lib().smcpp_b200_multi_get_shard(self._h, ctypes.c_int(i), ptr(out, c_i32p), ctypes.byref(n)), "multi_get_shard")
        return out

    def context(self, i: int) -> "Context":
        """A non-owning view of the i-th device's context (options, stats, q)."""
        h = ctypes.c_void_p()
        self._check(lib().smcpp_b200_multi_context(self._h, ctypes.c_int(i), ctypes.byref(h)), "multi_context")
        c = Context.__new__(Context)
        c._h, c._borrowed = h, True
        c.K, c.npop, c.M = self.K, self.npop, 0
        c.C = len(self.shard(i))
        return c

    def estep(self, pi, T, E) -> dict:
        pi = np.ascontiguousarray(pi, np.float64); T = np.ascontiguousarray(T, np.float64); E = np.ascontiguousarray(E, np.float64)
        M, K, C = pi.shape[0], self.K, self.C
        if E.shape != (K, M) or T.shape != (M, M):
            raise ValueError("shape mismatch: T must be [M,M], E must be [K,M]")
        out = {"ll": np.empty(C), "xisum": np.empty((C, M, M)), "gamma0": np.empty((C, M)), "gamma_sums": np.empty((C, K, M)),
               "key_present": np.empty((C, K), np.uint8), "reduced": np.empty(1 + M + M * M + K * M)}
        self._check(lib().smcpp_b200_multi_estep(self._h, ctypes.c_int(M), ptr(pi, c_f64p), ptr(T, c_f64p), ptr(E, c_f64p),
                                                 ptr(out["ll"], c_f64p), ptr(out["xisum"], c_f64p), ptr(out["gamma0"], c_f64p),
                                                 ptr(out["gamma_sums"], c_f64p), ptr(out["key_present"], c_u8p),
                                                 ptr(out["reduced"], c_f64p)), "multi_estep")
        return out

    def close(self):
        if self._h:
            lib().smcpp_b200_multi_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ObsPipeline:
    """Device-resident observation pre-processing of ONE contig (include/smcpp_b200.h: smcpp_b200_obs_*): the reference's
    Thin -> BinObservations -> RecodeMonomorphic -> Compress chain (smcpp/analysis/analysis.py:60-63)."""

    def __init__(self, rows: np.ndarray, device: int = 0):
        rows = np.ascontiguousarray(rows, np.int32)
        if rows.ndim != 2 or rows.shape[1] not in (4, 7):
            raise ValueError("rows must be int32 [L, 1 + 3*npop]")
        self._h = ctypes.c_void_p()
        if lib().smcpp_b200_obs_create(ctypes.byref(self._h), ctypes.c_int(device)):
            raise RuntimeError("smcpp_b200_obs_create: " + lib().smcpp_b200_obs_last_error(None).decode())
        self.ms = {}
        self.upload(rows)

    def upload(self, rows: np.ndarray):
        """(Re)load one contig's rows; device buffers of earlier runs are reused."""
        rows = np.ascontiguousarray(rows, np.int32)
        self.npop = (rows.shape[1] - 1) // 3
        self._check(lib().smcpp_b200_obs_upload(self._h, ptr(rows, c_i32p), ctypes.c_int64(rows.shape[0]), ctypes.c_int(self.npop)), "upload")
        return self

    def _check(self, rc, what):
        if rc:
            raise RuntimeError(f"obs_{what}: " + lib().smcpp_b200_obs_last_error(self._h).decode())

    def _done(self, what):
        self.ms[what] = float(lib().smcpp_b200_obs_last_ms(self._h))
        return self

    def thin(self, thinning: int, offset: int = 0):
        self._check(lib().smcpp_b200_obs_thin(self._h, ctypes.c_int(thinning), ctypes.c_int(offset)), "thin")
        return self._done("thin")

    def bin(self, a, w: int):
        a = np.ascontiguousarray(a, np.int64)
        self._check(lib().smcpp_b200_obs_bin(self._h, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.c_int64(w)), "bin")
        return self._done("bin")

    def recode_monomorphic(self, a):
        a = np.ascontiguousarray(a, np.int64)
        self._check(lib().smcpp_b200_obs_recode_monomorphic(self._h, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))), "recode_monomorphic")
        return self._done("recode_monomorphic")

    def compress(self):
        self._check(lib().smcpp_b200_obs_compress(self._h), "compress")
        return self._done("compress")

    def recode_nonseg(self, cutoff: int):
        self._check(lib().smcpp_b200_obs_recode_nonseg(self._h, ctypes.c_int64(cutoff)), "recode_nonseg")
        return self._done("recode_nonseg")

    def break_long_spans(self, span_cutoff: int) -> np.ndarray:
        """Cuts the current rows at long missing spans; returns the piece offsets (n_pieces + 1) into the broken array, which
        stays on the device; select_piece(p) makes one piece the current rows."""
        npieces = ctypes.c_int64()
        self._check(lib().smcpp_b200_obs_break_long_spans(self._h, ctypes.c_int64(span_cutoff), ctypes.byref(npieces)), "break_long_spans")
        self._done("break_long_spans")
        self.piece_offsets = np.empty(npieces.value + 1, np.int64)
        self._check(lib().smcpp_b200_obs_piece_offsets(self._h, self.piece_offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))), "piece_offsets")
        return self.piece_offsets

    def select_piece(self, p: int):
        self._check(lib().smcpp_b200_obs_select_piece(self._h, ctypes.c_int64(p), ctypes.c_int64(int(self.piece_offsets[p])),
                                                      ctypes.c_int64(int(self.piece_offsets[p + 1]))), "select_piece")
        return self

    @property
    def n_rows(self) -> int:
        return int(lib().smcpp_b200_obs_rows(self._h))

    def rows(self) -> np.ndarray:
        out = np.empty((self.n_rows, 1 + 3 * self.npop), np.int32)
        self._check(lib().smcpp_b200_obs_download(self._h, ptr(out, c_i32p)), "download")
        return out

    def close(self):
        if self._h:
            lib().smcpp_b200_obs_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

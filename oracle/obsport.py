"""ctypes front-end of oracle/obs_port.c (TEST INFRASTRUCTURE, see oracle/__init__.py): the reference's observation
pre-processing chain thin_data -> bin_observations -> RecodeMonomorphic -> compress_repeated_obs, restated on the CPU."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
I32P = ctypes.POINTER(ctypes.c_int32)
LP = ctypes.POINTER(ctypes.c_long)


def build() -> str:
    so = os.path.join(_HERE, "libsmcb_obs_oracle.so")
    src = os.path.join(_HERE, "obs_port.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        for f in ("smcb_oracle_thin", "smcb_oracle_bin", "smcb_oracle_compress", "smcb_oracle_recode_nonseg", "smcb_oracle_break_long_spans"):
            getattr(_LIB, f).restype = ctypes.c_long
        _LIB.smcb_oracle_recode_monomorphic.restype = None
    return _LIB


def _rows(a):
    a = np.ascontiguousarray(a, np.int32)
    assert a.ndim == 2 and (a.shape[1] - 1) % 3 == 0
    return a, (a.shape[1] - 1) // 3


def thin_data(data, thinning: int, offset: int = 0) -> np.ndarray:
    data, npop = _rows(data)
    cap = int((2 * np.ceil(data[:, 0] / thinning)).sum()) + 2 * data.shape[0] + 2
    out = np.zeros((cap, data.shape[1]), np.int32)
    r = lib().smcb_oracle_thin(data.ctypes.data_as(I32P), ctypes.c_long(data.shape[0]), ctypes.c_int(npop), ctypes.c_int(thinning),
                               ctypes.c_int(offset), out.ctypes.data_as(I32P), ctypes.c_long(cap))
    assert r >= 0
    return out[:r].copy()


def bin_observations(data, a, w: int) -> np.ndarray:
    data, npop = _rows(data)
    data = data.copy()
    na = np.ascontiguousarray(a, np.int64)
    out = np.zeros((int(data[:, 0].astype(np.int64).sum()) // w + 1, data.shape[1]), np.int32)
    r = lib().smcb_oracle_bin(data.ctypes.data_as(I32P), ctypes.c_long(data.shape[0]), ctypes.c_int(npop), na.ctypes.data_as(LP),
                              ctypes.c_long(w), out.ctypes.data_as(I32P))
    return out[:r].copy()


def recode_monomorphic(data, a) -> np.ndarray:
    data, npop = _rows(data)
    data = data.copy()
    na = np.ascontiguousarray(a, np.int64)
    lib().smcb_oracle_recode_monomorphic(data.ctypes.data_as(I32P), ctypes.c_long(data.shape[0]), ctypes.c_int(npop), na.ctypes.data_as(LP))
    return data


def compress_repeated_obs(data) -> np.ndarray:
    data, _ = _rows(data)
    out = np.zeros_like(data)
    r = lib().smcb_oracle_compress(data.ctypes.data_as(I32P), ctypes.c_long(data.shape[0]), ctypes.c_int(data.shape[1]),
                                   out.ctypes.data_as(I32P))
    return out[:r].copy()


def recode_nonseg(data, cutoff: int) -> np.ndarray:
    data, npop = _rows(data)
    data = data.copy()
    lib().smcb_oracle_recode_nonseg(data.ctypes.data_as(I32P), ctypes.c_long(data.shape[0]), ctypes.c_int(npop), ctypes.c_long(cutoff))
    return data


def break_long_spans(data, cutoff: int) -> list:
    """List of pieces (each with its leading missing row), as the reference returns one Contig per piece."""
    data, npop = _rows(data)
    out = np.zeros((data.shape[0] + 1, data.shape[1]), np.int32)
    off = np.zeros(data.shape[0] + 2, np.int64)
    n = lib().smcb_oracle_break_long_spans(data.ctypes.data_as(I32P), ctypes.c_long(data.shape[0]), ctypes.c_int(npop), ctypes.c_long(cutoff),
                                           out.ctypes.data_as(I32P), off.ctypes.data_as(LP))
    return [out[off[i]:off[i + 1]].copy() for i in range(n)]

"""SMCB1 bundles: a flat list of named little-endian arrays.

A plain container used by pytest / bench.py to hand workloads to external checker binaries and read
their dumps back (a C++ reader/writer of the same format ships with the test infrastructure).  Format::

    SMCB1\\n
    <name> <dtype> <ndim> <d0> ... <dN-1>\\n      dtype in {i4,i8,f4,f8,u1}
    <raw bytes>\\n
"""
from __future__ import annotations

import numpy as np

_DT = {"i4": np.int32, "i8": np.int64, "f4": np.float32, "f8": np.float64, "u1": np.uint8}
_RDT = {np.dtype(v): k for k, v in _DT.items()}


def save(path, arrays: dict) -> None:
    with open(path, "wb") as f:
        f.write(b"SMCB1\n")
        for name, a in arrays.items():
            a = np.asarray(a)
            if a.dtype == np.bool_:
                a = a.astype(np.uint8)
            if a.dtype not in _RDT:
                if np.issubdtype(a.dtype, np.integer):
                    a = a.astype(np.int32)
                elif np.issubdtype(a.dtype, np.floating):
                    a = a.astype(np.float64)
                else:
                    raise TypeError(f"bundle: unsupported dtype {a.dtype} for {name}")
            a = np.ascontiguousarray(a)
            if a.ndim == 0:
                a = a.reshape(1)
            hdr = f"{name} {_RDT[a.dtype]} {a.ndim} " + " ".join(str(d) for d in a.shape) + "\n"
            f.write(hdr.encode())
            f.write(a.tobytes())
            f.write(b"\n")


def load(path) -> dict:
    out = {}
    with open(path, "rb") as f:
        if not f.readline().startswith(b"SMCB1"):
            raise ValueError(f"bundle: bad magic in {path}")
        while True:
            line = f.readline()
            if not line.strip():
                break
            parts = line.decode().split()
            name, dt, ndim = parts[0], parts[1], int(parts[2])
            shape = tuple(int(x) for x in parts[3:3 + ndim])
            count = int(np.prod(shape)) if shape else 1
            raw = f.read(count * np.dtype(_DT[dt]).itemsize)
            f.read(1)
            out[name] = np.frombuffer(raw, dtype=_DT[dt]).reshape(shape).copy()
    return out

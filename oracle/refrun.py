"""Runs the compiled reference (oracle/_ref/ref_harness).  TEST INFRASTRUCTURE, see oracle/__init__.py."""
from __future__ import annotations

import os
import subprocess
import tempfile

from smcpp_b200 import bundle

_HERE = os.path.dirname(os.path.abspath(__file__))
HARNESS = os.path.join(_HERE, "_ref", "ref_harness")


def available() -> bool:
    return os.path.exists(HARNESS) and os.access(HARNESS, os.X_OK)


def build() -> bool:
    """Compile the reference in place when /root/reference is here; the GPU box uses the prebuilt file."""
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "ref_build")])
    return available()


def run(workload, threads: int = 1, repeat: int = 1, save_gamma=False, dump_alpha=False, timeout=None) -> dict:
    if not available():
        raise RuntimeError("oracle/_ref/ref_harness is not built (make -C oracle/ref_build)")
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.smcb"), os.path.join(td, "out.smcb")
        bundle.save(fin, workload.to_bundle(save_gamma=save_gamma, dump_alpha=dump_alpha))
        env = dict(os.environ, OMP_PROC_BIND="spread", OMP_NUM_THREADS=str(threads))
        r = subprocess.run([HARNESS, fin, fout, str(threads), str(repeat)], capture_output=True, text=True, env=env,
                           timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError(f"ref_harness failed ({r.returncode}): {r.stderr[-2000:]}")
        return bundle.load(fout)

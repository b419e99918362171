"""Runs the compiled reference (oracle/_ref/ref_harness).  TEST INFRASTRUCTURE, see oracle/__init__.py."""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile

from smcpp_b200 import bundle

_HERE = os.path.dirname(os.path.abspath(__file__))
HARNESS = os.path.join(_HERE, "_ref", "ref_harness")
HARNESS_B200 = os.path.join(_HERE, "_ref", "ref_harness_b200")   # the same harness with the reference's Estep routed to libsmcpp_b200


def available() -> bool:
    return os.path.exists(HARNESS) and os.access(HARNESS, os.X_OK)


def available_b200() -> bool:
    return os.path.exists(HARNESS_B200) and os.access(HARNESS_B200, os.X_OK)


def build() -> bool:
    """Compile the reference in place when /root/reference is here; the GPU box uses the prebuilt file."""
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "ref_build")])
    return available()


class Pending:
    """A reference run in flight (several can run side by side: single-contig configs use one thread each)."""

    def __init__(self, proc, td, fout):
        self.proc, self.td, self.fout = proc, td, fout
        self._res = None

    def result(self, timeout=None) -> dict:
        if self._res is None:
            try:
                _, err = self.proc.communicate(timeout=timeout)
                if self.proc.returncode != 0:
                    raise RuntimeError(f"ref_harness failed ({self.proc.returncode}): {err[-2000:]}")
                self._res = bundle.load(self.fout)
            finally:
                shutil.rmtree(self.td, ignore_errors=True)
        return self._res


def start(workload, threads: int = 1, repeat: int = 1, save_gamma=False, dump_alpha=False, harness=None, extra_env=None) -> Pending:
    harness = harness or HARNESS
    if not (os.path.exists(harness) and os.access(harness, os.X_OK)):
        raise RuntimeError(f"{harness} is not built (make -C oracle/ref_build)")
    td = tempfile.mkdtemp(prefix="smcb_ref_")
    fin, fout = os.path.join(td, "in.smcb"), os.path.join(td, "out.smcb")
    bundle.save(fin, workload.to_bundle(save_gamma=save_gamma, dump_alpha=dump_alpha))
    env = dict(os.environ, OMP_PROC_BIND="spread", OMP_NUM_THREADS=str(threads))
    env.update(extra_env or {})
    proc = subprocess.Popen([harness, fin, fout, str(threads), str(repeat)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                            text=True, env=env)
    return Pending(proc, td, fout)


def run(workload, threads: int = 1, repeat: int = 1, save_gamma=False, dump_alpha=False, timeout=None, harness=None,
        extra_env=None) -> dict:
    return start(workload, threads, repeat, save_gamma, dump_alpha, harness, extra_env).result(timeout)

"""Shared helpers of the test-suite: golden loading and error metrics."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                      if not os.path.basename(p).startswith("model_"))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.inp = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
        self.ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
        self.npop = int(self.inp["npop"].reshape(-1)[0])
        lens = self.inp["contig_lengths"]
        offs = np.concatenate([[0], np.cumsum(lens)])
        self.contigs = [np.ascontiguousarray(self.inp["obs"][offs[i]:offs[i + 1]]) for i in range(len(lens))]
        self.M = self.ref["pi"].shape[0]


def relmax(a, b):
    """max |a - b| relative to the largest reference entry."""
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max())


# tolerances of the parity tests (BASELINE.json north_star: log-likelihood within 1e-8 relative;
# statistics within 1e-7 of the largest entry).  What we actually reach is ~1e-12 / ~1e-9.
LL_RTOL = 1e-8
STAT_RTOL = 1e-7


def load_model(config: str) -> dict:
    """Per-E-step inputs of a BASELINE config as produced by the reference (tests/golden/make_model_inputs.py)."""
    z = np.load(os.path.join(GOLDEN_DIR, f"model_{config}.npz"))
    return {k: z[k] for k in z.files}

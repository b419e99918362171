"""Generates tests/golden/model_<config>.npz: the per-E-step inputs (pi, transition, emission table, key
table, eigensystems) of each BASELINE.json config, produced by the UNMODIFIED reference's do_dirty_work /
TransitionBundle::update on a prefix of the config's synthetic contigs (oracle/_ref/ref_harness).

bench.py and the full-size GPU tests load these instead of running anything under oracle/ in the measured
arm.  The inputs depend on the model, hidden states, CSFS and the KEY UNIVERSE only -- the script asserts
that the prefix already contains every key of the full-size workload.

    python tests/golden/make_model_inputs.py      # needs /root/reference (this container)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from smcpp_b200 import synth  # noqa: E402
from oracle import refrun  # noqa: E402

CONFIGS = ["C1", "C2", "C3", "C4", "C5-16", "C5-32", "C5-64", "C5-128"]
PREFIX = 60000


def main():
    if not refrun.build():
        raise SystemExit("needs /root/reference to build oracle/_ref/ref_harness")
    for name in CONFIGS:
        full = synth.config(name)
        keys_full = np.unique(np.concatenate([np.unique(c[:, 1:], axis=0) for c in full.contigs]), axis=0)
        eig_full = np.unique(np.concatenate([np.unique(c[c[:, 0] > 1][:, 1:], axis=0) for c in full.contigs]), axis=0)
        small = synth.config(name)
        small.contigs = [np.ascontiguousarray(c[:PREFIX]) for c in full.contigs[:2]]
        ref = refrun.run(small, threads=2)
        assert np.array_equal(ref["keys"], keys_full), f"{name}: prefix misses keys of the full workload"
        assert np.array_equal(ref["keys"][ref["eig_key_idx"]], eig_full), f"{name}: prefix misses span>1 keys"
        out = {k: ref[k] for k in ("pi", "T", "keys", "E", "eig_key_idx", "eig_P", "eig_Pinv", "eig_d", "eig_dscaled",
                                   "eig_scale", "eig_cplx")}
        out["hidden_states"] = full.hidden_states
        out["total_blocks"] = np.int64(full.total_blocks)
        path = os.path.join(HERE, f"model_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: M={full.M} K={ref['keys'].shape[0]} n_eig={len(ref['eig_key_idx'])} blocks={full.total_blocks} "
              f"-> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

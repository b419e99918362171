"""Generates the committed golden vectors under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/ref_harness, built in place from /root/reference by oracle/ref_build/Makefile).

    python tests/golden/make_golden.py        # needs /root/reference (this container, not the GPU box)

Each <name>.npz holds the inputs (observation rows, hidden states, model, CSFS) and every dump of the
reference for that input: pi, T, keys, E, eigensystems, per-contig ll / xisum / gamma0 / gamma_sums,
alpha_hat / log_c, Q.  The reference has no numeric golden for this path of its own, so
these are the pins.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from smcpp_b200 import synth  # noqa: E402
from oracle import refrun  # noqa: E402


def ref_test_inference_workload(contigs=3):
    """Input recipe of the reference's own test/unit/test_inference.py:11-47 (51 hidden-state boundaries,
    n = 30, rows incl. a span-200000 row with nb > 0 and span-10 missing rows), with our synthetic model."""
    hs = np.array([0, 0.002, 0.0024992427075529156, 0.0031231070556282147, 0.0039027012668429368, 0.0048768988404573679,
                   0.0060942769312431738, 0.0076155385891087321, 0.0095165396414589112, 0.011892071150027212,
                   0.014860586049702963, 0.018570105657341358, 0.023205600571298769, 0.028998214001102113,
                   0.036236787437156658, 0.045282263373729439, 0.0565856832591419, 0.070710678118654752,
                   0.088361573317084705, 0.11041850887031313, 0.1379813265364985, 0.15879323234898887,
                   0.22690003122622915, 0.28675095012241164, 0.35039604900931776, 0.4174620285802807,
                   0.48093344839252727, 0.54048403452772453, 0.58902987679112695, 0.63973400753929655,
                   0.6661845719884536, 0.68097444812291441, 0.69652310395210704, 0.71291262669986732,
                   0.73023918985303526, 0.74861647270557707, 0.76818018497781393, 0.7890941548490632,
                   0.81155867710242946, 0.8429182938518559, 0.88146343535942318, 0.92368486081866963,
                   0.97035848127888702, 1.0225351498208244, 1.1293598575982273, 1.2553186915845398, 1.468142830257521,
                   1.7982719448467761, 2.3740247153419043, 3.2719144602927757, 4.8068671176749671, np.inf])
    n = 30
    fakeobs = [[1, -1, 0, 0], [1, 1, 0, 0], [10, 0, 0, 0], [10, -1, 0, 0], [200000, 0, 0, n - 2], [1, 1, n - 4, n - 2]]
    fakeobs = np.array(fakeobs * 20, np.int32)
    a, s = synth.model()
    M = len(hs) - 1
    return synth.Workload(name="ref_test_inference", npop=1, n=(n - 2,), na=(2,), M=M, hidden_states=hs,
                          contigs=[fakeobs.copy() for _ in range(contigs)], model_a=a, model_s=s, theta=0.0025,
                          rho=0.0031206103977654887, sfs=synth.dummy_sfs(hs, (n - 2,)))


def ragged_workload():
    """Ragged contigs: long, short, single-row, and one without any span > 1 row."""
    w = synth.make_workload("ragged", 1, 700, 16, 4)
    c0 = w.contigs[0]
    c1 = synth.make_contig(50, w.n, 77)
    c2 = c0[:1].copy()                       # a single (missing) row
    c3 = c0[c0[:, 0] == 1][:40].copy()       # only span-1 rows
    w.contigs = [c0, c1, c2, np.ascontiguousarray(c3)]
    return w


def cases():
    yield "c1_2k", synth.make_workload("c1_2k", 1, 2000, 16, 4)
    yield "c2_1500", synth.make_workload("c2_1500", 1, 1500, 32, 10)
    yield "c4_twopop_1200", synth.make_workload("c4_twopop_1200", 2, 1200, 32, (6, 6), npop=2)
    yield "m17_800", synth.make_workload("m17_800", 1, 800, 17, 6)
    yield "m64_600", synth.make_workload("m64_600", 1, 600, 64, 10)
    yield "ragged", ragged_workload()
    yield "ref_test_inference", ref_test_inference_workload()


def main():
    if not refrun.build():
        raise SystemExit("oracle/_ref/ref_harness could not be built (needs /root/reference)")
    for name, w in cases():
        ref = refrun.run(w, dump_alpha=True)
        inp = w.to_bundle(dump_alpha=True)
        out = {("in_" + k): np.asarray(v) for k, v in inp.items()}
        out.update({("ref_" + k): v for k, v in ref.items() if k not in ("estep_seconds", "ctor_seconds", "threads")})
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: L={[c.shape[0] for c in w.contigs]} M={w.M} K={ref['keys'].shape[0]} ll={ref['ll'].sum():.12g} "
              f"-> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

"""Runs a few E-steps of BASELINE config 2 (1 x 10^6 blocks, M = 32) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smcpp_b200 import capi, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"model_{cfg}.npz"))
model = {k: z[k] for k in z.files}
w = synth.config(cfg)
ctx = capi.Context(0)
ctx.set_contigs(w.contigs, w.npop, model["keys"])
for i in range(steps):
    ctx.estep_device(model["pi"], model["T"], model["E"], model, upload=(i == 0))
    print(ctx.stats())
out = ctx.fetch()
print("ll", out["ll"].sum())

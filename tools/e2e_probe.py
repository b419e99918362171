import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smcpp_b200 import capi, synth
cfg = "C3"
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"model_{cfg}.npz"))
model = {k: z[k] for k in z.files}
w = synth.config(cfg)
ctx = capi.Context(0)
ctx.set_contigs(w.contigs, w.npop, model["keys"])
for i in range(3):
    ctx.estep(model["pi"], model["T"], model["E"], model)
for i in range(10):
    t0 = time.perf_counter(); o = ctx.estep(model["pi"], model["T"], model["E"], model); t1 = time.perf_counter()
    st = ctx.stats()
    print("e2e wall %.2f ms  device total %.2f (setup %.3f fwd %.2f stats %.2f fin %.2f) sweeps %d %d" % (1e3*(t1-t0), st["ms_total"], st["ms_setup"], st["ms_forward"], st["ms_stats"], st["ms_finalize"], st["fwd_sweeps"], st["bwd_sweeps"]))
for i in range(16):
    t0 = time.perf_counter(); ctx.estep_device(model["pi"], model["T"], model["E"], model, upload=False); t1 = time.perf_counter()
    st = ctx.stats()
    print("resident wall %.2f ms  device total %.2f (fwd %.2f stats %.2f)" % (1e3*(t1-t0), st["ms_total"], st["ms_forward"], st["ms_stats"]))

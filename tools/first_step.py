import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from smcpp_b200 import capi, synth
cfg = sys.argv[1]
z = np.load(f"tests/golden/model_{cfg}.npz"); model = {k: z[k] for k in z.files}
w = synth.config(cfg)
for opts in ({}, {"burn_in_blocks": 768}):
    ctx = capi.Context(0)
    for k, v in opts.items(): ctx.set_option(k, v)
    ctx.set_contigs(w.contigs, w.npop, model["keys"])
    for i in range(3):
        ctx.estep_device(model["pi"], model["T"], model["E"], model, upload=True)
        st = ctx.stats()
        print(opts, i, {k: st[k] for k in ("n_chunks", "chunk_blocks", "burn_in_blocks", "fwd_sweeps", "bwd_sweeps", "fwd_redone", "bwd_redone", "fwd_max_mismatch", "bwd_max_mismatch", "ms_total")}, flush=True)
    ctx.close()

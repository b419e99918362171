import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smcpp_b200 import capi, synth
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "model_C3.npz"))
model = {k: z[k] for k in z.files}
ids = [int(x) for x in sys.argv[1].split(",")]
contigs = [synth.make_contig(1_000_000, (20,), 1000 + c, 1) for c in ids]
for opts in ({}, {"mma_min_chunks": 10**9}, {"target_warps": 9472 * 4}):
    ctx = capi.Context(0)
    for k, v in opts.items(): ctx.set_option(k, v)
    ctx.set_contigs(contigs, 1, model["keys"])
    o = ctx.estep(model["pi"], model["T"], model["E"], model)
    st = ctx.stats()
    print(opts, "ll", o["ll"], "chunks", st["n_chunks"], "Lc", st["chunk_blocks"], "sweeps", st["fwd_sweeps"], st["bwd_sweeps"], "redone", st["fwd_redone"], st["bwd_redone"], "mm", st["fwd_max_mismatch"], st["bwd_max_mismatch"])
    ctx.close()

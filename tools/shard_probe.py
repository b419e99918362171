"""Per-rank shard sizes of the scaling run (C3 on 2/4/8 GPUs = 11 / 6 / 3 contigs): E-step time under different planner choices."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smcpp_b200 import capi, synth
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "model_C3.npz"))
model = {k: z[k] for k in z.files}
ids = [int(x) for x in sys.argv[1].split(",")]
contigs = [synth.make_contig(1_000_000, (20,), 1000 + c, 1) for c in ids]
variants = [{}]
for opts in variants:
    ctx = capi.Context(0)
    for k, v in opts.items(): ctx.set_option(k, v)
    ctx.set_contigs(contigs, 1, model["keys"])
    best = None
    for i in range(4):
        ctx.estep_device(model["pi"], model["T"], model["E"], model, upload=(i == 0))
        st = ctx.stats()
        if i and (best is None or st["ms_total"] < best["ms_total"]): best = st
    print(len(ids), opts, "chunks", best["n_chunks"], "Lc", best["chunk_blocks"], "total %.2f fwd||bwd %.2f stats %.2f" % (best["ms_total"], best["ms_forward"], best["ms_stats"]),
          "burn", best["burn_in_blocks"], "mm %.2e %.2e" % (best["fwd_max_mismatch"], best["bwd_max_mismatch"]), "lockstep %.4f" % (best["mma_steps"] / max(1, 8 * best["mma_rounds"])), "fwd-only %.2f bwd-only %.2f" % (best["ms_forward_only"], best["ms_backward"]), flush=True)
    ctx.set_option("chunks_per_warp", 0)
    ctx.close()

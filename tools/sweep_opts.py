"""Times one E-step of a BASELINE config under several option sets in ONE process (the synthetic contigs are generated once).

    python tools/sweep_opts.py C3 "fwd_cached_keys=0" "fwd_cached_keys=3" "fwd_cached_keys=3,burn_in_blocks=384"
    python tools/sweep_opts.py C3:3 ...        # the first 3 contigs only (one rank's shard of an 8-GPU run)

Prints device times (CUDA events inside the library) per option set: total / recursions / forward alone / backward alone / statistics.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from smcpp_b200 import capi, synth  # noqa: E402


def main():
    spec = sys.argv[1]
    cfg, _, ncont = spec.partition(":")
    sets = sys.argv[2:] or [""]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests", "golden", f"model_{cfg}.npz"))
    model = {k: z[k] for k in z.files}
    w = synth.config(cfg)
    contigs = w.contigs[:int(ncont)] if ncont else w.contigs
    steps = int(os.environ.get("SWEEP_STEPS", "5"))
    upload = os.environ.get("SWEEP_UPLOAD", "0") == "1"      # 1: every step uploads the inputs and lets the library compute the eigensystems
    ll0 = None
    for s in sets:
        ctx = capi.Context(0)
        for item in s.split(","):
            if "=" in item:
                k, v = item.split("=", 1)
                ctx.set_option(k.strip(), float(v))
        ctx.set_contigs(contigs, w.npop, model["keys"])
        ctx.estep_device(model["pi"], model["T"], model["E"], model, upload=True)
        for _ in range(2):
            ctx.estep_device(model["pi"], model["T"], model["E"], model, upload=False)
        acc = {}
        for _ in range(steps):
            ctx.estep_device(model["pi"], model["T"], model["E"], None if upload else model, upload=upload)
            st = ctx.stats()
            for k in ("ms_total", "ms_forward", "ms_forward_only", "ms_backward", "ms_stats", "ms_finalize"):
                acc.setdefault(k, []).append(st[k])
        st = ctx.stats()
        ll = float(ctx.fetch()["ll"].sum())
        if ll0 is None:
            ll0 = ll
        print(f"[{s or 'default'}] chunks={st['n_chunks']}x{st['chunk_blocks']} burn={st['burn_in_blocks']} sweeps={st['fwd_sweeps']}/{st['bwd_sweeps']} "
              + " ".join(f"{k[3:]}={np.median(v):.3f}" for k, v in acc.items())
              + f" mismatch={st['fwd_max_mismatch']:.2e}/{st['bwd_max_mismatch']:.2e} ll_rel_vs_first={abs(ll - ll0) / abs(ll0):.2e}", flush=True)
        ctx.close()


if __name__ == "__main__":
    main()

"""Development check on a GPU box: GPU E-step vs the compiled reference on several configs."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smcpp_b200 import capi, synth
from oracle import refrun


def compare(name, scale, use_ref_eig=True, opts=None, ref_threads=1, check_alpha=False):
    w = synth.config(name, scale)
    t0 = time.time(); ref = refrun.run(w, threads=ref_threads, dump_alpha=check_alpha); t_ref = time.time() - t0
    ctx = capi.Context(0)
    for k, v in (opts or {}).items():
        ctx.set_option(k, v)
    t0 = time.time(); ctx.set_contigs(w.contigs, w.npop); t_set = time.time() - t0
    assert (ctx.keys == ref["keys"]).all(), "key table mismatch"
    assert (ctx.eig_keys == ref["eig_key_idx"]).all(), "eig keys mismatch"
    eig = ref if use_ref_eig else None
    out = ctx.estep(ref["pi"], ref["T"], ref["E"], eig)
    t0 = time.time(); out = ctx.estep(ref["pi"], ref["T"], ref["E"], eig); t_gpu = time.time() - t0
    st = ctx.stats()
    res = {"cfg": name, "scale": scale, "blocks": w.total_blocks, "M": w.M, "ref_eig": use_ref_eig}
    llr, llg = ref["ll"].sum(), out["ll"].sum()
    res["ll_rel"] = float(abs(llg - llr) / abs(llr))
    for k in ("xisum", "gamma0", "gamma_sums"):
        res[k] = float(np.abs(out[k] - ref[k]).max() / np.abs(ref[k]).max())
    res["present_ok"] = bool((out["key_present"] == ref["key_present"]).all())
    if check_alpha:
        a = ctx.debug_alpha_hat(0); b = ref["alpha_hat_0"]
        res["alpha_exact_frac"] = float((a == b).mean()); res["alpha_maxdiff"] = float(np.abs(a - b).max())
    res["ref_estep_s"] = float(ref["estep_seconds"][-1]); res["gpu_call_s"] = t_gpu; res["set_contigs_s"] = t_set
    res["stats"] = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()}
    print(json.dumps(res), flush=True)
    ctx.close()
    return res


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "basic"
    if which == "basic":
        compare("C1", 1.0, check_alpha=True, opts={"force_sequential": 1})
        compare("C1", 1.0, check_alpha=True)
        compare("C1", 1.0, use_ref_eig=False)
        compare("C2", 0.02, check_alpha=True, opts={"force_sequential": 1})
        compare("C2", 0.02, check_alpha=True)
        compare("C2", 0.2)
        compare("C4", 0.02)
        compare("C5-64", 0.01)
        compare("C5-128", 0.004)
        compare("C3", 0.01, ref_threads=8)
    elif which == "big":
        compare("C2", 1.0)

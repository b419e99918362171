"""Randomised parity sweep on a GPU box: random small data sets, models and planner options through the C ABI against the
oracle port (tests' tolerances).  Prints one line per failing case and a summary; exit code 1 on any failure.

    python tools/fuzz_parity.py [n_cases] [seed]
"""
import os, sys, json, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from smcpp_b200 import capi
from oracle import port
from helpers import LL_RTOL, STAT_RTOL, relmax


def random_contig(rng, L, npop, n, span_mode):
    W = 1 + 3 * npop
    o = np.zeros((L, W), np.int32)
    if span_mode == "ones":
        o[:, 0] = 1
    elif span_mode == "big":
        o[:, 0] = rng.integers(2, 100000, L)
    elif span_mode == "few":
        o[:, 0] = rng.choice([1, 1, 2, 7, 300], L)
    elif span_mode == "alt":
        o[:, 0] = np.where(np.arange(L) % 2 == 0, 1, np.minimum(2 + rng.geometric(1 / 200.0, L), 50000))
    else:
        o[:, 0] = np.where(rng.random(L) < 0.5, 1, rng.integers(2, 3000, L))
    nkeys = int(rng.integers(1, 12))
    pool = np.zeros((nkeys, 3 * npop), np.int32)
    for k in range(nkeys):
        for p in range(npop):
            nb = int(rng.integers(0, n[p] + 1))
            pool[k, 3 * p:3 * p + 3] = [int(rng.integers(-1, 3)), int(rng.integers(0, nb + 1)), nb]
    o[:, 1:] = pool[rng.integers(0, nkeys, L)]
    return o


def model(rng, M, K):
    base = rng.random((M, M)) ** 4 + np.eye(M) * rng.choice([5, 50, 500])
    S = base + base.T
    for _ in range(200):
        d = S.sum(1)
        S = S / np.sqrt(d[:, None] * d[None, :])
    T = (1 - 1e-5) * S + 1e-5 / (M + 1)
    pi = rng.random(M) + 0.1
    pi /= pi.sum()
    E = np.clip(rng.random((K, M)) * 0.9 + 0.05, 1e-3, 1.0)
    return pi, T, E


def one_case(rng, idx):
    npop = int(rng.choice([1, 1, 2]))
    n = tuple(int(x) for x in rng.integers(1, 8, npop))
    M = int(rng.choice([1, 2, 5, 13, 16, 31, 32, 33, 40, 64]))
    C = int(rng.integers(1, 5))
    span_mode = str(rng.choice(["ones", "big", "few", "alt", "mix"]))
    contigs = [random_contig(rng, int(rng.choice([1, 2, 3, 9, 100, 700, 2500])), npop, n, span_mode) for _ in range(C)]
    keys = np.unique(np.concatenate([c[:, 1:] for c in contigs]), axis=0)
    K = keys.shape[0]
    pi, T, E = model(rng, M, K)
    eig_idx = np.array([k for k in range(K) if any(((c[:, 0] > 1) & (c[:, 1:] == keys[k]).all(1)).any() for c in contigs)], np.int32)
    eig = capi.host_eigensystems(T, E, eig_idx)
    ref = {"pi": pi, "T": T, "E": E, "keys": keys, **eig}
    opts = {"chunk_blocks": int(rng.choice([0, 16, 50, 128, 999])), "burn_in_blocks": int(rng.choice([0, 64, 512])),
            "mma_min_chunks": int(rng.choice([1, 64])), "force_mma_forward": int(rng.integers(0, 2)),
            "chunks_per_warp": int(rng.choice([0, 1, 2, 4, 8])), "fwd_cached_keys": int(rng.choice([0, 2, 4])),
            "slab_blocks": int(rng.choice([32, 256, 16384])), "tiles": int(rng.choice([1, 2])), "max_restarts": int(rng.choice([0, 8])),
            "stats_streams": int(rng.choice([1, 2]))}
    if rng.random() < 0.15:
        opts = {"force_sequential": 1}
    desc = {"case": idx, "npop": npop, "M": M, "C": C, "L": [c.shape[0] for c in contigs], "K": K, "n_eig": len(eig_idx), "spans": span_mode, "opts": opts}
    if eig["eig_cplx"].any():     # irregular spectra have their own tests against the compiled reference (tests/test_irregular_spectra.py)
        return None, desc
    ctx = capi.Context(0)
    try:
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.set_contigs(contigs, npop, keys)
        out = ctx.estep(pi, T, E, ref)
        worst = 0.0
        for c, obs in enumerate(contigs):
            o = port.hmm_estep(obs, ref)
            e_ll = abs(out["ll"][c] - o["ll"]) / max(abs(o["ll"]), 1e-300)
            errs = {"ll": e_ll / LL_RTOL}
            for k in ("xisum", "gamma0", "gamma_sums"):
                errs[k] = relmax(out[k][c], o[k]) / STAT_RTOL
            errs["present"] = 0.0 if np.array_equal(out["key_present"][c].astype(bool), np.isin(np.arange(K), port.key_ids(obs, keys))) else 9.9
            worst = max(worst, max(errs.values()))
            if not np.isfinite(list(errs.values())).all():
                worst = float("inf")
        return worst, desc
    finally:
        ctx.set_option("chunks_per_warp", 0)
        ctx.set_option("fwd_cached_keys", 0)
        ctx.close()


def main():
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    bad = skipped = 0
    worst_all = 0.0
    for i in range(ncases):
        try:
            worst, desc = one_case(rng, i)
        except Exception as ex:
            print("EXC", i, repr(ex)[:300], flush=True)
            traceback.print_exc()
            bad += 1
            continue
        if worst is None:
            skipped += 1
            continue
        worst_all = max(worst_all, worst) if np.isfinite(worst) else float("inf")
        if not (worst <= 1.0):
            bad += 1
            print("FAIL", json.dumps(desc), "worst error / tolerance = %.3g" % worst, flush=True)
    print(f"fuzz: {ncases} cases, {bad} failed, {skipped} skipped (complex spectrum), worst error/tolerance {worst_all:.3g}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

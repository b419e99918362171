"""Randomised bit-exactness sweep of the device observation pipeline against the oracle restatement (GPU box).

    python tools/fuzz_obs.py [n_cases] [seed]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from smcpp_b200 import capi
from oracle import obsport


def rows(rng, L, npop, n, a, maxspan):
    W = 1 + 3 * npop
    d = np.zeros((L, W), np.int32)
    d[:, 0] = rng.integers(1, maxspan + 1, L)
    for p in range(npop):
        nb = rng.integers(0, n[p] + 1, L)
        aa = rng.integers(-1, a[p] + 1, L)
        d[:, 1 + 3 * p] = aa
        d[:, 3 + 3 * p] = nb
        d[:, 2 + 3 * p] = np.where(aa < 0, 0, (rng.random(L) * (nb + 1)).astype(np.int64))
    # runs of repeated keys so that compress has something to merge
    rep = rng.random(L) < 0.5
    for j in range(1, L):
        if rep[j]:
            d[j, 1:] = d[j - 1, 1:]
    return d


def main():
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    bad = 0
    for i in range(ncases):
        npop = int(rng.choice([1, 2]))
        a = (2,) if npop == 1 else tuple(rng.choice([(2, 0), (1, 1)]))
        n = tuple(int(x) for x in rng.integers(1, 6, npop))
        L = int(rng.choice([1, 2, 5, 40, 600, 5000]))
        maxspan = int(rng.choice([1, 3, 50, 2000]))
        thinning = int(rng.choice([1, 2, 7, 100, 5000]))
        offset = int(rng.choice([0, 0, 1, thinning - 1, thinning + 3]))
        w = int(rng.choice([1, 3, 100, 1000]))
        cut = int(rng.choice([2, 10, 1000]))
        d = rows(rng, L, npop, n, a, maxspan)
        desc = dict(case=i, npop=npop, a=a, n=n, L=L, maxspan=maxspan, thinning=thinning, offset=offset, w=w, cut=cut)
        try:
            p = capi.ObsPipeline(d)
            ok = np.array_equal(p.recode_nonseg(cut).rows(), obsport.recode_nonseg(d, cut))
            d1 = obsport.compress_repeated_obs(obsport.recode_nonseg(d, cut))
            ok &= np.array_equal(p.compress().rows(), d1)
            off = p.break_long_spans(cut)
            pieces = obsport.break_long_spans(d1, cut)
            ok &= list(np.diff(off)) == [x.shape[0] for x in pieces]
            j = int(rng.integers(0, len(pieces)))
            x = pieces[j]
            p.select_piece(j)
            t = obsport.thin_data(x, thinning, offset)
            ok &= np.array_equal(p.thin(thinning, offset).rows(), t)
            b = obsport.bin_observations(t, a, w)
            ok &= np.array_equal(p.bin(a, w).rows(), b)
            rc = obsport.recode_monomorphic(b, a)
            ok &= np.array_equal(p.recode_monomorphic(a).rows(), rc)
            ok &= np.array_equal(p.compress().rows(), obsport.compress_repeated_obs(rc))
            p.close()
        except Exception as ex:
            ok = False
            desc["exc"] = repr(ex)[:200]
        if not ok:
            bad += 1
            print("FAIL", desc, flush=True)
    print(f"fuzz_obs: {ncases} cases, {bad} failed")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

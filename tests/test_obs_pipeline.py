"""Input side of the E-step: thin_data -> bin_observations -> RecodeMonomorphic -> compress_repeated_obs.

CPU part: the oracle restatement (oracle/obs_port.c) against what the reference's own functions produced
(tests/golden/obs/obs_pipeline.npz, made by tests/golden/make_obs_golden.py from the reference's Cython / Python sources).
GPU part: the CUDA pipeline (smcpp_b200/csrc/obs_pipeline.cu through the C ABI) against the goldens and, on larger fresh
inputs, against the oracle -- bit for bit (integer work)."""
import os

import numpy as np
import pytest

from oracle import obsport

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obs", "obs_pipeline.npz"))
NAMES = [str(x) for x in GOLD["names"]]


def case(name):
    g = {k.split("__", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "__")}
    npop, thinning, w = (int(x) for x in g["params"])
    return g, npop, thinning, w


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_outputs(name):
    g, npop, thinning, w = case(name)
    thin = obsport.thin_data(g["raw"], thinning)
    assert np.array_equal(thin, g["thin"])
    binned = obsport.bin_observations(g["thin"], g["a"], w)
    assert np.array_equal(binned, g["binned"])
    rec = obsport.recode_monomorphic(g["binned"], g["a"])
    assert np.array_equal(rec, g["recoded"])
    assert np.array_equal(obsport.compress_repeated_obs(g["recoded"]), g["compressed"])
    assert np.array_equal(obsport.compress_repeated_obs(g["raw"]), g["raw_compressed"])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_front_of_chain_matches_reference(name):
    # RecodeNonseg(cutoff) -> Compress -> BreakLongSpans(cutoff): reference smcpp/analysis/base.py:50-52
    g, npop, thinning, w = case(name)
    cut_ns, cut_bl = (int(x) for x in g["cutoffs"])
    ns = obsport.recode_nonseg(g["raw"], cut_ns)
    assert np.array_equal(ns, g["nonseg"])
    pieces = obsport.break_long_spans(obsport.compress_repeated_obs(ns), cut_bl)
    assert [p.shape[0] for p in pieces] == list(g["piece_len"])
    assert np.array_equal(np.concatenate(pieces, axis=0), g["pieces"])


def test_oracle_invariants():
    g, npop, thinning, w = case("p1")
    thin = obsport.thin_data(g["raw"], thinning)
    assert thin[:, 0].astype(np.int64).sum() == g["raw"][:, 0].astype(np.int64).sum()        # reference's own assert (:82)
    comp = obsport.compress_repeated_obs(thin)
    assert comp[:, 0].astype(np.int64).sum() == thin[:, 0].astype(np.int64).sum()
    assert (np.abs(np.diff(comp[:, 1:], axis=0)).sum(axis=1) > 0).all()                    # no two neighbours share a key


# ---------------------------------------------------------------------------------------------------------------
# GPU: the CUDA pipeline through the C ABI
# ---------------------------------------------------------------------------------------------------------------
def _raw_rows(rng, L, npop, n, a):
    """Fresh seeded rows (vectorised twin of make_obs_golden.raw_rows; the oracle is the checker here)."""
    W = 1 + 3 * npop
    d = np.zeros((L, W), np.int32)
    u = rng.random(L)
    kind = np.select([u < 0.45, u < 0.55, u < 0.62], [0, 1, 2], 3)
    d[:, 0] = np.select([kind == 0, kind == 1, kind == 2], [rng.geometric(1 / 300.0, L), rng.geometric(1 / 150.0, L), 1], rng.integers(1, 3, L))
    for p in range(npop):
        nb = rng.integers(0, n[p] + 1, L)
        aa = rng.integers(0, a[p] + 1, L)
        aa = np.where(rng.random(L) < 0.1, -1, aa)
        bb = (rng.random(L) * (nb + 1)).astype(np.int64)
        col = 1 + 3 * p
        d[:, col] = np.select([kind == 0, kind == 1, kind == 2], [0, -1, a[p]], aa)
        d[:, col + 2] = np.select([kind == 0, kind == 1], [n[p], 0], nb)
        d[:, col + 1] = np.select([kind == 0, kind == 1, kind == 2], [0, 0, nb], np.where(aa < 0, 0, bb))
    return d


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_pipeline_matches_reference_outputs(name):
    from smcpp_b200 import capi
    g, npop, thinning, w = case(name)
    p = capi.ObsPipeline(g["raw"])
    assert np.array_equal(p.thin(thinning).rows(), g["thin"])
    assert np.array_equal(p.bin(g["a"], w).rows(), g["binned"])
    assert np.array_equal(p.recode_monomorphic(g["a"]).rows(), g["recoded"])
    assert np.array_equal(p.compress().rows(), g["compressed"])
    p.close()
    p = capi.ObsPipeline(g["raw"])
    assert np.array_equal(p.compress().rows(), g["raw_compressed"])
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("npop,n,a,L,thinning,w,offset", [(1, (10,), (2,), 400_000, 1243, 100, 0), (2, (6, 4), (1, 1), 200_000, 3, 7, 2),
                                                            (1, (3,), (2,), 150_000, 1, 1000, 0), (1, (8,), (2,), 100_000, 50, 100, 77)])
def test_gpu_pipeline_against_oracle_on_fresh_inputs(npop, n, a, L, thinning, w, offset):
    from smcpp_b200 import capi
    rng = np.random.default_rng(L + thinning)
    raw = _raw_rows(rng, L, npop, n, a)
    p = capi.ObsPipeline(raw)
    thin = obsport.thin_data(raw, thinning, offset)
    assert np.array_equal(p.thin(thinning, offset).rows(), thin)
    assert thin[:, 0].astype(np.int64).sum() == raw[:, 0].astype(np.int64).sum()
    binned = obsport.bin_observations(thin, a, w)
    assert np.array_equal(p.bin(a, w).rows(), binned)
    rec = obsport.recode_monomorphic(binned, a)
    assert np.array_equal(p.recode_monomorphic(a).rows(), rec)
    assert np.array_equal(p.compress().rows(), obsport.compress_repeated_obs(rec))
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_front_of_chain_matches_reference(name):
    from smcpp_b200 import capi
    g, npop, thinning, w = case(name)
    cut_ns, cut_bl = (int(x) for x in g["cutoffs"])
    p = capi.ObsPipeline(g["raw"])
    assert np.array_equal(p.recode_nonseg(cut_ns).rows(), g["nonseg"])
    off = p.compress().break_long_spans(cut_bl)
    assert list(np.diff(off)) == list(g["piece_len"])
    got = [p.select_piece(i).rows() for i in range(len(off) - 1)]
    assert np.array_equal(np.concatenate(got, axis=0), g["pieces"])
    # a piece goes on through the rest of the chain like any contig
    big = int(np.argmax(np.diff(off)))
    want = obsport.compress_repeated_obs(obsport.bin_observations(obsport.thin_data(got[big], thinning), g["a"], w))
    assert np.array_equal(p.select_piece(big).thin(thinning).bin(g["a"], w).compress().rows(), want)
    p.close()


@pytest.mark.gpu
def test_gpu_pipeline_feeds_the_estep_context():
    # the compressed rows are exactly what set_contigs() expects (span > 0, C-contiguous int32)
    from smcpp_b200 import capi
    g, npop, thinning, w = case("p1")
    p = capi.ObsPipeline(g["raw"])
    rows = p.thin(thinning).bin(g["a"], w).recode_monomorphic(g["a"]).compress().rows()
    p.close()
    ctx = capi.Context(0)
    ctx.set_contigs([rows], npop)
    assert ctx.total_blocks == rows.shape[0] and (rows[:, 0] > 0).all()
    ctx.close()


@pytest.mark.gpu
def test_gpu_pipeline_errors():
    from smcpp_b200 import capi
    with pytest.raises(ValueError):
        capi.ObsPipeline(np.zeros((4, 5), np.int32))
    p = capi.ObsPipeline(np.array([[3, 0, 0, 2]], np.int32))
    with pytest.raises(RuntimeError, match="thinning"):
        p.thin(0)
    with pytest.raises(RuntimeError, match="w >= 1"):
        p.bin([2], 0)
    p.close()


# ---------------------------------------------------------------------------------------------------------------
# Live pin (build container only): the oracle against the reference's own functions on random inputs
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not os.path.isdir("/root/reference/smcpp"), reason="the reference tree is only mounted in the build container")
def test_oracle_matches_live_reference_functions_on_random_inputs():
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_obs_golden", os.path.join(here, "golden", "make_obs_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    try:
        et = mk.build_reference_cython()
    except Exception as ex:   # no Cython / compiler on this box
        pytest.skip(f"cannot build the reference's Cython module here: {ex}")
    compress = mk.reference_function(os.path.join(mk.REF, "estimation_tools.py"), "compress_repeated_obs")
    recode = mk.reference_function(os.path.join(mk.REF, "data_filter.py"), "_recode", cls="RecodeMonomorphic")
    rng = np.random.default_rng(2024)
    for case_no in range(60):
        npop = int(rng.choice([1, 2]))
        a = (2,) if npop == 1 else tuple(int(x) for x in rng.choice([(2, 0), (1, 1)]))
        n = tuple(int(x) for x in rng.integers(1, 7, npop))
        L = int(rng.choice([1, 3, 50, 800]))
        thinning = int(rng.choice([1, 2, 9, 150, 4000]))
        w = int(rng.choice([1, 4, 100, 1000]))
        raw = mk.raw_rows(rng, L, npop, n, a, long_runs=bool(rng.integers(0, 2)))
        # offset > 0 only on longer inputs: the reference sizes its output from data[offset:] (rows, :17), too small otherwise
        offset = int(rng.choice([0, 1, thinning - 1])) if L >= 50 else 0
        try:
            thin = et.thin_data(raw.copy(), thinning, offset)
        except IndexError:      # the reference's own output buffer was too small for this offset: nothing to pin against
            offset = 0
            thin = et.thin_data(raw.copy(), thinning)
        assert np.array_equal(obsport.thin_data(raw, thinning, offset), thin), (case_no, "thin", offset)
        c = mk.FakeContig(thin.copy(), a)
        binned = np.array(et.bin_observations(c, w))
        assert np.array_equal(obsport.bin_observations(thin, a, w), binned), (case_no, "bin")
        c2 = mk.FakeContig(binned.copy(), a)
        recode(None, c2)
        assert np.array_equal(obsport.recode_monomorphic(binned, a), c2.data), (case_no, "recode")
        assert np.array_equal(obsport.compress_repeated_obs(c2.data), compress(c2.data.copy())), (case_no, "compress")

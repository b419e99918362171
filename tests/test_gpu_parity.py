"""Parity tests proper: the CUDA path, called through the C ABI (ctypes -> libsmcpp_b200.so), against
  * the golden vectors produced by the unmodified reference (tests/golden/*.npz),
  * the oracle port on fresh seeded inputs,
  * the compiled reference itself where oracle/_ref travelled to this box,
and, at BASELINE.json's full sizes, size-independent properties.

Tolerances (helpers.py): log-likelihood 1e-8 relative (BASELINE.json north_star); xi / gamma statistics 1e-7
of the largest entry.  Sequential mode (one chunk per contig) must reproduce the reference's
float alpha_hat bit for bit.
"""
import numpy as np
import pytest

from helpers import GOLDEN_NAMES, LL_RTOL, STAT_RTOL, Golden, load_model, relmax
from oracle import port, refrun
from smcpp_b200 import capi, synth
from smcpp_b200.inference import InferenceManager

pytestmark = pytest.mark.gpu


def run_ctx(contigs, npop, ref, opts=None, lib_eig=False, keys=None):
    ctx = capi.Context(0)
    for k, v in (opts or {}).items():
        ctx.set_option(k, v)
    ctx.set_contigs(contigs, npop, keys)
    out = ctx.estep(ref["pi"], ref["T"], ref["E"], None if lib_eig else ref)
    return ctx, out


def check_against(out, ref, ll_rtol=LL_RTOL, stat_rtol=STAT_RTOL):
    assert abs(out["ll"].sum() - ref["ll"].sum()) <= ll_rtol * abs(ref["ll"].sum())
    assert np.all(np.abs(out["ll"] - ref["ll"]) <= ll_rtol * np.abs(ref["ll"]))
    for k in ("xisum", "gamma0", "gamma_sums"):
        for c in range(len(ref["ll"])):
            assert relmax(out[k][c], ref[k][c]) <= stat_rtol, (k, c)
    assert np.array_equal(out["key_present"], ref["key_present"])


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_golden_sequential_bit_exact_forward(name):
    g = Golden(name)
    ctx, out = run_ctx(g.contigs, g.npop, g.ref, {"force_sequential": 1})
    assert np.array_equal(ctx.keys, g.ref["keys"])
    assert np.array_equal(ctx.eig_keys, g.ref["eig_key_idx"])
    check_against(out, g.ref, ll_rtol=1e-13, stat_rtol=1e-11)
    for c in range(len(g.contigs)):
        assert np.array_equal(ctx.debug_alpha_hat(c), g.ref[f"alpha_hat_{c}"]), "float alpha_hat differs"
    ctx.close()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
@pytest.mark.parametrize("chunk,burn", [(64, 64), (128, 512), (37, 300)])
def test_golden_chunked(name, chunk, burn):
    g = Golden(name)
    ctx, out = run_ctx(g.contigs, g.npop, g.ref, {"chunk_blocks": chunk, "burn_in_blocks": burn})
    check_against(out, g.ref)
    st = ctx.stats()
    assert st["chunk_blocks"] <= chunk
    ctx.close()


@pytest.mark.parametrize("name", GOLDEN_NAMES)     # M = 16, 17, 32 (one 32-state tile), 51, 64 (two tiles)
@pytest.mark.parametrize("chunk,burn", [(64, 64), (100, 512), (37, 300), (16, 0)])
@pytest.mark.parametrize("tiles", [1, 2])
def test_golden_chunked_tensor_path(name, chunk, burn, tiles):
    """The DMMA recursions, forced on for small inputs: 8 chunks per warp (recursion_mma.cu, tiles = 1) and, for M <= 32,
    16 chunks per warp (recursion_mma2.cu, tiles = 2)."""
    g = Golden(name)
    ctx, out = run_ctx(g.contigs, g.npop, g.ref, {"chunk_blocks": chunk, "burn_in_blocks": burn, "mma_min_chunks": 1,
                                                   "force_mma_forward": 1, "tiles": tiles})
    check_against(out, g.ref)
    ctx.close()


@pytest.mark.parametrize("name", ["c1_2k", "c2_1500", "m17_800", "c4_twopop_1200", "ragged"])
@pytest.mark.parametrize("G,fused", [(8, 0), (8, 1), (2, 0), (1, 1)])
def test_two_tile_recursions_are_bitwise_the_one_tile_recursions(name, G, fused):
    """Per chunk the two-tile kernels execute the one-tile kernels' arithmetic (same fragments, same order): identical
    results for identical chunking, whatever the chunks-per-tile count and the launch form."""
    g = Golden(name)
    opts = {"chunk_blocks": 48, "burn_in_blocks": 512, "mma_min_chunks": 1, "force_mma_forward": 1, "chunks_per_warp": G}
    ctx1, a = run_ctx(g.contigs, g.npop, g.ref, dict(opts, tiles=1))
    ctx2, b = run_ctx(g.contigs, g.npop, g.ref, dict(opts, tiles=2, fused_recursions=fused))
    for k in ("ll", "xisum", "gamma0", "gamma_sums", "reduced"):
        assert np.array_equal(a[k], b[k]), k
    for c in range(len(g.contigs)):
        assert np.array_equal(ctx1.debug_alpha_hat(c), ctx2.debug_alpha_hat(c))
    check_against(b, g.ref)
    ctx1.close()
    ctx2.close()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
@pytest.mark.parametrize("G", [8, 2])
def test_golden_tensor_path_several_chunks_per_warp(name, G):
    """Small inputs run with one chunk per warp; pin 8 (and 2) so that every row of the MMA carries a chunk."""
    g = Golden(name)
    ctx = capi.Context(0)
    try:
        ctx.set_option("chunks_per_warp", G)
        for k, v in {"chunk_blocks": 48, "burn_in_blocks": 512, "mma_min_chunks": 1, "force_mma_forward": 1}.items():
            ctx.set_option(k, v)
        ctx.set_contigs(g.contigs, g.npop)
        out = ctx.estep(g.ref["pi"], g.ref["T"], g.ref["E"], g.ref)
        check_against(out, g.ref)
    finally:
        ctx.set_option("chunks_per_warp", 0)
        ctx.close()


@pytest.mark.parametrize("name,G,cached", [("c2_1500", 1, 4), ("c2_1500", 8, 4), ("c2_1500", 8, 0), ("c1_2k", 8, 2), ("m64_600", 8, 4),
                                           ("m17_800", 4, 4), ("c4_twopop_1200", 8, 1)])
def test_tensor_path_float_step_is_bit_exact(name, G, cached):
    """A contig of span-1 blocks only exercises nothing but the float GEMV step of the tensor-path forward kernel (packed
    FFMA2 arithmetic, step matrices in shared or global memory): every alpha_hat column must equal the reference
    semantics (the port, pinned bit for bit to the reference in tests/test_oracle.py) exactly."""
    g = Golden(name)
    contigs = []
    for c in g.contigs:
        c = c.copy()
        c[:, 0] = 1
        contigs.append(c)
    ref = dict(g.ref)
    for k in ("eig_P", "eig_Pinv", "eig_d", "eig_dscaled", "eig_scale", "eig_key_idx"):
        ref[k] = g.ref[k][:0]
    ctx = capi.Context(0)
    try:
        ctx.set_option("chunks_per_warp", G)
        ctx.set_option("fwd_cached_keys", cached)
        for k, v in {"chunk_blocks": 150, "burn_in_blocks": 512, "mma_min_chunks": 1, "force_mma_forward": 1}.items():
            ctx.set_option(k, v)
        ctx.set_contigs(contigs, g.npop, g.ref["keys"])
        assert len(ctx.eig_keys) == 0
        out = ctx.estep(ref["pi"], ref["T"], ref["E"], ref)
        for c, obs in enumerate(contigs):
            o = port.hmm_estep(obs, ref, want_alpha=True)
            a = ctx.debug_alpha_hat(c)
            assert np.array_equal(a[:151], o["alpha_hat"][:151]), "chunk 0 follows the reference chain from pi: must be exact"
            assert (a == o["alpha_hat"]).mean() > 0.999      # later chunks start from a burn-in state (float-converged)
            assert abs(out["ll"][c] - o["ll"]) <= 1e-11 * abs(o["ll"])
            assert relmax(out["xisum"][c], o["xisum"]) <= STAT_RTOL
    finally:
        ctx.set_option("chunks_per_warp", 0)
        ctx.set_option("fwd_cached_keys", 0)
        ctx.close()


@pytest.mark.parametrize("name", ["c1_2k", "c2_1500", "m17_800", "m64_600", "ref_test_inference"])
def test_golden_library_eigensystems(name):
    g = Golden(name)
    ctx, out = run_ctx(g.contigs, g.npop, g.ref, lib_eig=True)
    check_against(out, g.ref)
    ctx.close()


def test_short_burn_in_is_repaired_by_sweeps():
    # burn-in far too short: boundary checks must fail and the sweeps must restore the exact chain
    g = Golden("c2_1500")
    ctx, out = run_ctx(g.contigs, g.npop, g.ref, {"chunk_blocks": 50, "burn_in_blocks": 2, "max_restarts": 0})
    st = ctx.stats()
    assert st["fwd_redone"] > 0 and st["fwd_sweeps"] > 1
    assert st["bwd_redone"] > 0 and st["bwd_sweeps"] > 1
    check_against(out, g.ref)
    # the context lengthens its burn-in for the next E-step
    out2 = ctx.estep(g.ref["pi"], g.ref["T"], g.ref["E"], g.ref)
    assert ctx.stats()["burn_in_blocks"] > 2
    check_against(out2, g.ref)
    ctx.close()


def test_mass_failures_rerun_pass_zero_with_a_longer_burn_in():
    """Many failed boundaries = the burn-in is far too short: pass 0 is run again with a doubled burn-in (fully parallel)
    instead of one dependent sweep per failed chunk in a row; the longer burn-in stays for the next E-step."""
    g = Golden("c2_1500")
    ctx, out = run_ctx(g.contigs, g.npop, g.ref, {"chunk_blocks": 25, "burn_in_blocks": 2, "mma_min_chunks": 1})
    st = ctx.stats()
    assert st["restarts"] >= 1 and st["fwd_sweeps"] + st["bwd_sweeps"] <= 6
    check_against(out, g.ref)
    out2 = ctx.estep(g.ref["pi"], g.ref["T"], g.ref["E"], g.ref)
    assert ctx.stats()["restarts"] == 0
    check_against(out2, g.ref)
    ctx.close()


def test_zero_burn_in_degenerates_to_sequential_sweeps():
    g = Golden("c1_2k")
    ctx, out = run_ctx([g.contigs[0][:400]], g.npop, g.ref, {"chunk_blocks": 40, "burn_in_blocks": 0, "max_restarts": 0}, keys=g.ref["keys"])
    seq = port.hmm_estep(g.contigs[0][:400], g.ref)
    st = ctx.stats()
    assert st["fwd_sweeps"] > 1 and st["bwd_sweeps"] > 1      # every boundary starts wrong and is repaired
    # re-run chunks are swept until their boundaries agree bit for bit (fwd_tol = 0), the backward pass to 1e-10
    assert abs(out["ll"][0] - seq["ll"]) <= LL_RTOL * abs(seq["ll"])
    assert relmax(out["xisum"][0], seq["xisum"]) <= STAT_RTOL
    ctx.close()


@pytest.mark.parametrize("M,n,L", [(16, 4, 3000), (32, 10, 2500), (8, 3, 1200), (33, 5, 700), (96, 8, 400), (128, 6, 300)])
def test_fresh_inputs_against_port(M, n, L):
    # fresh seeded inputs; pi / T / E come from a golden-independent recipe so no reference is needed on the box
    rng = np.random.default_rng(M * 1000 + L)
    w = synth.make_workload("fresh", 2, L, M, n, seed0=900 + M)
    keys = np.unique(np.concatenate([c[:, 1:] for c in w.contigs]), axis=0)
    K = keys.shape[0]
    # a symmetric doubly stochastic chain (Sinkhorn) with the reference's uniform mixing, so that diag(e) T^T is similar
    # to a symmetric matrix and every eigensystem is real; emissions in (0, 1]
    base = rng.random((M, M)) ** 4 + np.eye(M) * 50
    S = base + base.T
    for _ in range(200):
        d = S.sum(1)
        S = S / np.sqrt(d[:, None] * d[None, :])
    T = (1 - 1e-5) * S + 1e-5 / (M + 1)
    pi = rng.random(M) + 0.1
    pi /= pi.sum()
    E = np.clip(rng.random((K, M)) * 0.9 + 0.05, 1e-3, 1.0)
    eig_idx = np.array([k for k in range(K) if any(((c[:, 0] > 1) & (c[:, 1:] == keys[k]).all(1)).any() for c in w.contigs)], np.int32)
    eig = capi.host_eigensystems(T, E, eig_idx)
    assert not eig["eig_cplx"].any()
    ref = {"pi": pi, "T": T, "E": E, "keys": keys, **eig}
    ctx, out = run_ctx(w.contigs, 1, ref, {"chunk_blocks": 256, "burn_in_blocks": 512, "mma_min_chunks": 1, "force_mma_forward": M % 2})
    for c, obs in enumerate(w.contigs):
        o = port.hmm_estep(obs, ref)
        assert abs(out["ll"][c] - o["ll"]) <= LL_RTOL * abs(o["ll"])
        for k in ("xisum", "gamma0", "gamma_sums"):
            assert relmax(out[k][c], o[k]) <= STAT_RTOL, k
    ctx.close()


@pytest.mark.parametrize("spans", [(2, 5, 40), (3,), tuple(range(2, 400))])
def test_span_sorted_statistics_against_port(spans):
    """k_stats32e: long runs of equal span (one rank-1 stream per run, weighted once), runs that cross work items (> 4096
    span>1 blocks per contig) and all-different spans (rank-2 groups), against the port on the same inputs."""
    M, n, L = 32, 6, 9000
    rng = np.random.default_rng(len(spans))
    w = synth.make_workload("fresh", 1, L, M, n, seed0=77)
    obs = w.contigs[0].copy()
    big = obs[:, 0] > 1
    obs[big, 0] = rng.choice(np.asarray(spans, np.int32), size=int(big.sum()))
    keys = np.unique(obs[:, 1:], axis=0)
    K = keys.shape[0]
    base = rng.random((M, M)) ** 4 + np.eye(M) * 50
    S = base + base.T
    for _ in range(200):
        d = S.sum(1)
        S = S / np.sqrt(d[:, None] * d[None, :])
    T = (1 - 1e-5) * S + 1e-5 / (M + 1)
    pi = rng.random(M) + 0.1
    pi /= pi.sum()
    E = np.clip(rng.random((K, M)) * 0.9 + 0.05, 1e-3, 1.0)
    eig_idx = np.array([k for k in range(K) if ((obs[:, 0] > 1) & (obs[:, 1:] == keys[k]).all(1)).any()], np.int32)
    eig = capi.host_eigensystems(T, E, eig_idx)
    assert not eig["eig_cplx"].any()
    ref = {"pi": pi, "T": T, "E": E, "keys": keys, **eig}
    assert int(big.sum()) > 4096                                     # more than one work item
    ctx, out = run_ctx([obs], 1, ref, {"chunk_blocks": 512, "burn_in_blocks": 512, "mma_min_chunks": 1})
    o = port.hmm_estep(obs, ref)
    assert abs(out["ll"][0] - o["ll"]) <= LL_RTOL * abs(o["ll"])
    for k in ("xisum", "gamma0", "gamma_sums"):
        assert relmax(out[k][0], o[k]) <= STAT_RTOL, k
    ctx.close()


def test_many_keys_two_populations():
    """More distinct observation keys than k_stats32 keeps in shared memory (two-population full-SFS data): the per-key
    gamma sums then accumulate in global memory."""
    M, L = 32, 20000
    rng = np.random.default_rng(4242)
    w = synth.make_workload("fresh", 1, L, M, (14, 14), npop=2, seed0=31)
    obs = w.contigs[0]
    keys = np.unique(obs[:, 1:], axis=0)
    K = keys.shape[0]
    assert K > 288
    base = rng.random((M, M)) ** 4 + np.eye(M) * 50
    S = base + base.T
    for _ in range(200):
        d = S.sum(1)
        S = S / np.sqrt(d[:, None] * d[None, :])
    T = (1 - 1e-5) * S + 1e-5 / (M + 1)
    pi = rng.random(M) + 0.1
    pi /= pi.sum()
    E = np.clip(rng.random((K, M)) * 0.9 + 0.05, 1e-3, 1.0)
    eig_idx = np.array([k for k in range(K) if ((obs[:, 0] > 1) & (obs[:, 1:] == keys[k]).all(1)).any()], np.int32)
    eig = capi.host_eigensystems(T, E, eig_idx)
    ref = {"pi": pi, "T": T, "E": E, "keys": keys, **eig}
    ctx, out = run_ctx([obs], 2, ref, {"chunk_blocks": 1024, "burn_in_blocks": 512, "mma_min_chunks": 1})
    o = port.hmm_estep(obs, ref)
    assert abs(out["ll"][0] - o["ll"]) <= LL_RTOL * abs(o["ll"])
    for k in ("xisum", "gamma0", "gamma_sums"):
        assert relmax(out[k][0], o[k]) <= STAT_RTOL, k
    # gamma_sums holds exactly the keys that occur in the contig (reference src/hmm.cpp:51-53, 69)
    occurs = np.array([(obs[:, 1:] == keys[k]).all(1).any() for k in range(K)])
    assert np.array_equal(out["key_present"][0].astype(bool), occurs)
    assert np.array_equal((o["gamma_sums"] != 0).any(1), occurs)
    ctx.close()


def random_model(rng, M, obs_list):
    """A symmetric doubly stochastic chain (Sinkhorn) with the reference's uniform mixing (so that diag(e) T^T is similar
    to a symmetric matrix and every eigensystem is real), emissions in (0, 1], eigensystems by the library's host routine."""
    keys = np.unique(np.concatenate([c[:, 1:] for c in obs_list]), axis=0)
    K = keys.shape[0]
    base = rng.random((M, M)) ** 4 + np.eye(M) * 50
    S = base + base.T
    for _ in range(200):
        d = S.sum(1)
        S = S / np.sqrt(d[:, None] * d[None, :])
    T = (1 - 1e-5) * S + 1e-5 / (M + 1)
    pi = rng.random(M) + 0.1
    pi /= pi.sum()
    E = np.clip(rng.random((K, M)) * 0.9 + 0.05, 1e-3, 1.0)
    lut = {tuple(int(v) for v in k): i for i, k in enumerate(keys)}
    big = set()
    for c in obs_list:
        for row in np.unique(c[c[:, 0] > 1][:, 1:], axis=0):
            big.add(lut[tuple(int(v) for v in row)])
    eig_idx = np.array(sorted(big), np.int32)
    eig = capi.host_eigensystems(T, E, eig_idx)
    return {"pi": pi, "T": T, "E": E, "keys": keys, **eig}


def check_against_port(out, obs_list, ref):
    for c, obs in enumerate(obs_list):
        o = port.hmm_estep(obs, ref)
        assert abs(out["ll"][c] - o["ll"]) <= LL_RTOL * abs(o["ll"])
        for k in ("xisum", "gamma0", "gamma_sums"):
            assert relmax(out[k][c], o[k]) <= STAT_RTOL, k


def test_more_than_30_keys_with_span_above_one():
    """The reference builds one eigensystem per key that occurs with span > 1, without a cap (src/transition_bundle.cpp:14-25);
    round 1 stopped at 30 (5 bits of the block code, 32-bit slab masks)."""
    M, L = 16, 4000
    rng = np.random.default_rng(31)
    obs = np.zeros((L, 4), np.int32)
    obs[:, 0] = 1
    obs[0, 1] = -1
    run = np.arange(1, L, 2)
    obs[run, 0] = rng.integers(2, 40, size=run.size)
    pick = rng.integers(0, 45, size=run.size)                   # 45 different keys carry spans > 1
    obs[run, 1] = pick % 3
    obs[run, 2] = 1 + pick // 3
    obs[run, 3] = 20
    site = np.arange(2, L, 2)
    obs[site, 1] = rng.choice([1, 2, -1], size=site.size)
    ref = random_model(rng, M, [obs])
    assert len(ref["eig_key_idx"]) == 45 and not ref["eig_cplx"].any()
    for opts in ({"chunk_blocks": 200, "burn_in_blocks": 512}, {"chunk_blocks": 128, "burn_in_blocks": 512, "mma_min_chunks": 1, "force_mma_forward": 1},
                 {"force_sequential": 1}):
        ctx, out = run_ctx([obs], 1, ref, opts)
        assert len(ctx.eig_keys) == 45
        check_against_port(out, [obs], ref)
        ctx.close()


def test_more_than_2047_observation_keys():
    """Round 1 packed the key id into 11 bits; two-population full-SFS data exceed that."""
    M, L = 32, 7000
    rng = np.random.default_rng(2048)
    obs = np.zeros((L, 7), np.int32)
    obs[:, 0] = 1
    obs[0, 1] = -1
    obs[0, 4] = -1
    run = np.arange(1, L, 2)
    obs[run, 0] = rng.integers(2, 300, size=run.size)
    site = np.arange(2, L, 2)
    combo = rng.permutation(61 * 61 - 2)[:site.size] + 1          # distinct (b1, b2) pairs, never (0, 0) or (60, 60)
    obs[site, 1] = rng.integers(0, 3, size=site.size)
    obs[site, 2] = combo // 61
    obs[site, 3] = 60
    obs[site, 5] = combo % 61
    obs[site, 6] = 60
    ref = random_model(rng, M, [obs])
    assert ref["keys"].shape[0] > 2047
    ctx, out = run_ctx([obs], 2, ref, {"chunk_blocks": 512, "burn_in_blocks": 512, "mma_min_chunks": 1})
    assert ctx.K == ref["keys"].shape[0]
    check_against_port(out, [obs], ref)
    ctx.close()


def test_exhausted_repair_sweeps_are_an_error():
    """With boundaries still failing when max_sweeps is reached the results are wrong: estep() must say so (round 1 returned 0)."""
    g = Golden("c2_1500")
    ctx = capi.Context(0)
    for k, v in {"chunk_blocks": 50, "burn_in_blocks": 0, "max_sweeps": 2, "max_restarts": 0}.items():
        ctx.set_option(k, v)
    ctx.set_contigs(g.contigs, g.npop)
    with pytest.raises(RuntimeError, match="max_sweeps"):
        ctx.estep(g.ref["pi"], g.ref["T"], g.ref["E"], g.ref)
    assert ctx.stats()["converged"] == 0
    ctx.set_option("max_sweeps", 1 << 20)
    out = ctx.estep(g.ref["pi"], g.ref["T"], g.ref["E"], g.ref)
    assert ctx.stats()["converged"] == 1
    check_against(out, g.ref)
    ctx.close()


@pytest.mark.parametrize("fused", [0, 1])
def test_fused_and_separate_recursion_launches_agree_bitwise(fused):
    g = Golden("c4_twopop_1200")
    opts = {"chunk_blocks": 100, "burn_in_blocks": 512, "mma_min_chunks": 1, "force_mma_forward": 1, "chunks_per_warp": 8}
    ctx, a = run_ctx(g.contigs, g.npop, g.ref, dict(opts, fused_recursions=fused))
    ctx2, b = run_ctx(g.contigs, g.npop, g.ref, dict(opts, fused_recursions=1 - fused))
    for k in ("ll", "xisum", "gamma0", "gamma_sums", "reduced"):
        assert np.array_equal(a[k], b[k]), k
    check_against(a, g.ref)
    ctx.close()
    ctx2.close()


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref/ref_harness did not travel to this box")
@pytest.mark.parametrize("cfg,scale", [("C1", 1.0), ("C2", 0.05), ("C4", 0.05), ("C5-64", 0.01)])
def test_baseline_configs_against_live_reference(cfg, scale):
    w = synth.config(cfg, scale)
    ref = refrun.run(w)
    ctx, out = run_ctx(w.contigs, w.npop, ref)
    check_against(out, ref)
    ctx.close()


def test_errors_follow_the_reference():
    ctx = capi.Context(0)
    bad = np.array([[1, -1, 0, 0], [0, 0, 0, 0]], np.int32)
    with pytest.raises(RuntimeError, match="data are malformed: span <= 0"):
        ctx.set_contigs([bad], 1)
    ok = np.array([[1, -1, 0, 0], [5, 0, 0, 0], [1, 1, 0, 0]], np.int32)
    with pytest.raises(RuntimeError, match="missing from the explicit key table"):
        ctx.set_contigs([ok], 1, np.array([[-1, 0, 0], [0, 0, 0]], np.int32))
    with pytest.raises(RuntimeError, match="set_contigs"):
        ctx.estep(np.ones(4) / 4, np.eye(4), np.ones((0, 4)))     # no contigs yet (K = 0)
    ctx.set_contigs([ok], 1)
    with pytest.raises(ValueError):
        ctx.estep(np.ones(4) / 4, np.eye(4), np.ones((2, 4)))     # K = 3
    ctx.close()


def test_explicit_global_key_table_and_sharding_sum():
    # two shards with a shared key table must add up to the single-context result (the all-reduce contract)
    g = Golden("ragged")
    ctx, full = run_ctx(g.contigs, g.npop, g.ref)
    im = InferenceManager(g.contigs, g.inp["hidden_states"], npop=g.npop, devices=[0, 0], keys=g.ref["keys"])
    im.set_hmm_inputs(g.ref["pi"], g.ref["T"], g.ref["E"], g.ref)
    im.E_step()
    assert np.allclose(im.reduced, full["reduced"], rtol=1e-12, atol=0)
    assert im.loglik() == pytest.approx(g.ref["ll"].sum(), rel=1e-10)
    gs = im.gamma_sums
    for c in range(len(g.contigs)):
        present = {tuple(int(v) for v in g.ref["keys"][k]) for k in range(g.ref["keys"].shape[0]) if g.ref["key_present"][c, k]}
        assert set(gs[c].keys()) == present
        assert relmax(im.xisums[c], g.ref["xisum"][c]) < STAT_RTOL
    ctx.close()
    im.close()


def test_rerun_is_deterministic():
    g = Golden("c4_twopop_1200")
    ctx, a = run_ctx(g.contigs, g.npop, g.ref, {"chunk_blocks": 100})
    b = ctx.estep(g.ref["pi"], g.ref["T"], g.ref["E"], g.ref)
    for k in ("ll", "xisum", "gamma0", "gamma_sums", "reduced"):
        assert np.array_equal(a[k], b[k]), k
    ctx.close()


def test_full_size_properties_c2():
    """BASELINE config 2 at full size (1 x 10^6 blocks, M = 32): properties that need no reference run."""
    w = synth.config("C2")
    ref = load_model("C2")     # pi / T / E / eigensystems of this config, produced by the reference
    keys = np.unique(w.contigs[0][:, 1:], axis=0)
    assert np.array_equal(keys, ref["keys"])
    ctx, out = run_ctx(w.contigs, 1, ref)
    total_span = int(w.contigs[0][:, 0].astype(np.int64).sum())
    # every block's posterior sums to its span (src/hmm.cpp:120-121, 135-136)
    assert out["gamma_sums"].sum() == pytest.approx(total_span, rel=1e-9)
    assert np.isfinite(out["ll"]).all() and out["ll"][0] < 0
    assert (out["xisum"] >= 1e-20).all()
    assert np.array_equal(out["reduced"][1 + 32 + 1024:].reshape(-1, 32), out["gamma_sums"][0])
    # chunked == sequential (to the tolerance), at full size
    ctx2, seq = run_ctx(w.contigs, 1, ref, {"force_sequential": 1})
    assert abs(out["ll"][0] - seq["ll"][0]) <= 1e-10 * abs(seq["ll"][0])
    assert relmax(out["xisum"][0], seq["xisum"][0]) <= 1e-8
    assert relmax(out["gamma_sums"][0], seq["gamma_sums"][0]) <= 1e-8
    # and the first 3000 blocks agree with the port run on that prefix through alpha_hat
    a_gpu = ctx2.debug_alpha_hat(0)[:3001]
    o = port.hmm_estep(w.contigs[0][:3000], ref, want_alpha=True)
    assert np.array_equal(a_gpu, o["alpha_hat"])
    ctx.close()
    ctx2.close()


@pytest.mark.parametrize("name,opts", [("c1_2k", {"force_sequential": 1}), ("c2_1500", {"chunk_blocks": 64, "burn_in_blocks": 512}),
                                       ("c4_twopop_1200", {"chunk_blocks": 128, "burn_in_blocks": 512, "mma_min_chunks": 1}),
                                       ("m64_600", {"chunk_blocks": 64, "burn_in_blocks": 512, "mma_min_chunks": 1}),
                                       ("m17_800", {"chunk_blocks": 100, "burn_in_blocks": 512})])
def test_posterior_decoding_against_port(name, opts):
    # save_gamma: every column of the posterior (reference src/hmm.cpp:116-121, 134-136, 147-150) against the port,
    # which is pinned to the reference (tests/test_oracle.py)
    g = Golden(name)
    ref = g.ref
    ctx = capi.Context(0)
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.set_contigs(g.contigs, g.npop, ref["keys"])
    ctx.set_save_gamma(True)
    out = ctx.estep(ref["pi"], ref["T"], ref["E"], ref)
    check_against(out, ref)
    for c, obs in enumerate(g.contigs):
        o = port.hmm_estep(obs, ref, save_gamma=True)
        got = ctx.fetch_gamma(c)
        assert got.shape == o["gamma_full"].shape
        assert np.allclose(got[1:].sum(1), obs[:, 0], rtol=1e-9)          # a column sums to the block's span
        span = np.concatenate([[1], obs[:, 0]])[:, None]
        # columns compared as distributions.  The reference rounds alpha_hat to float every step (ulp 6e-8): with one
        # chunk per contig the float trajectory is reproduced bit for bit and the columns agree to fp64 round-off;
        # chunked runs start each chunk from a burn-in state that equals the reference's to a few float ulps, so a
        # single column carries that float-level noise (it averages out of the summed statistics, checked above).
        assert relmax(got / span, o["gamma_full"] / span) <= (1e-10 if opts.get("force_sequential") else 1e-6)
        # the per-key sums are the column sums of the full posterior (src/hmm.cpp:125-128, 140-143)
        tot = got[1:].sum(0)
        assert relmax(tot, out["gamma_sums"][c].sum(0)) <= 1e-9
    # the column normalisation of `smc++ posterior` (smcpp/commands/posterior.py:104-106) on the device
    ctx.set_save_gamma(True, normalise=True)
    ctx.estep(ref["pi"], ref["T"], ref["E"], ref)
    for c, obs in enumerate(g.contigs):
        o = port.hmm_estep(obs, ref, save_gamma=True)
        want = o["gamma_full"] / o["gamma_full"].sum(1, keepdims=True)
        got = ctx.fetch_gamma(c)
        assert np.allclose(got.sum(1), 1.0, rtol=1e-13)
        assert relmax(got, want) <= (1e-10 if opts.get("force_sequential") else 1e-6)
    ctx.set_save_gamma(False)
    ctx.estep(ref["pi"], ref["T"], ref["E"], ref)
    with pytest.raises(RuntimeError, match="save_gamma"):
        ctx.fetch_gamma(0)
    ctx.close()


def test_inference_manager_gammas_with_save_gamma():
    g = Golden("c2_1500")
    ref = g.ref
    im = InferenceManager(g.contigs, np.arange(ref["pi"].shape[0] + 1.0), g.npop, keys=ref["keys"])
    im.set_hmm_inputs(ref["pi"], ref["T"], ref["E"], ref)
    im.save_gamma = True
    im.E_step()
    gam = im.gammas
    assert len(gam) == len(g.contigs)
    for c, obs in enumerate(g.contigs):
        o = port.hmm_estep(obs, ref, save_gamma=True)
        assert gam[c].shape == (ref["pi"].shape[0], obs.shape[0] + 1)      # M x (L+1), as _PyInferenceManager.gammas
        span = np.concatenate([[1], obs[:, 0]])[:, None]
        assert relmax(gam[c].T / span, o["gamma_full"] / span) <= 1e-6


def test_randomised_parity_sweep():
    """tools/fuzz_parity.py: random small data sets (one/two populations, M = 1..64, degenerate span patterns, contigs of a
    single block) under random planner options, against the port."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                               "tools", "fuzz_parity.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    rng = np.random.default_rng(11)
    worst = 0.0
    for i in range(40):
        w, desc = fz.one_case(rng, i)
        if w is not None:
            assert w <= 1.0, desc
            worst = max(worst, w)
    assert worst > 0.0


def test_plain_c_consumer_agrees_with_the_python_binding(tmp_path):
    """examples/cabi_estep.c run as a program: same toy inputs through ctypes must give the same numbers."""
    import math
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "cabi_estep"
    subprocess.check_call(["gcc", "-std=c99", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "cabi_estep.c"),
                           "-L" + os.path.join(root, "smcpp_b200"), "-lsmcpp_b200", "-Wl,-rpath," + os.path.join(root, "smcpp_b200"), "-lm",
                           "-o", str(exe)])
    txt = subprocess.check_output([str(exe)], text=True)
    ll_c = [float(x) for x in txt.splitlines()[1].split()[1:3]]
    M = 4
    c0 = np.zeros((41, 4), np.int32)
    for i in range(41):
        c0[i] = [3 + 7 * (i % 5), -1 if i == 20 else 0, 0, 0] if i % 2 == 0 else [1, 1 + (i % 3 == 0), 2 if i % 7 == 1 else 0, 3 if i % 7 == 1 else 0]
    c1 = np.zeros((23, 4), np.int32)
    for i in range(23):
        c1[i] = [2 + 11 * (i % 3), 0, 0, 0] if i % 2 == 0 else [1, 1, 0, 0]
    ctx = capi.Context(0)
    ctx.set_contigs([c0, c1], 1)
    keys = ctx.keys
    pi = np.arange(1, M + 1) / (M * (M + 1) / 2.0)
    T = np.array([[50.0 if i == j else 1.0 / (1.0 + abs(i - j)) for j in range(M)] for i in range(M)])
    T /= T.sum(1, keepdims=True)
    E = np.array([[1.0 if a < 0 else math.exp(-0.02 * (i + 1)) if (a == 0 and b == 0) else 0.01 * (i + 1) * (1 + a) / (1.0 + b)
                   for i in range(M)] for a, b, nb in keys])
    out = ctx.estep(pi, T, E, None)
    assert np.allclose(out["ll"], ll_c, rtol=1e-11, atol=0)
    # and both agree with the port
    eig = ctx.eigensystems(T, E)
    ref = {"pi": pi, "T": T, "E": E, "keys": keys, **eig}
    for c, obs in enumerate((c0, c1)):
        assert abs(port.hmm_estep(obs, ref)["ll"] - ll_c[c]) <= LL_RTOL * abs(ll_c[c])
    ctx.close()

"""Edge cases of the statistics kernels' operand ring (k_stats32 / k_stats32e, smcpp_b200/csrc/stats32.cu): group counts
around the ring depth, partial last groups, slabs and items far smaller than a warp's share, one-block chunks (the
multiply-high column index of alpha_hat), and more keys than the per-slab gamma sums hold in shared memory.  The checker is the
oracle port on the same inputs (reference formulas: src/hmm.cpp:100-153)."""
import numpy as np
import pytest

from helpers import LL_RTOL, STAT_RTOL, relmax
from oracle import port
from smcpp_b200 import capi, synth
from test_gpu_parity import check_against_port, random_model, run_ctx

pytestmark = pytest.mark.gpu


def _workload(L, contigs=2, M=32, n=6, seed=5):
    return synth.make_workload("ring", contigs, L, M, n, seed0=3100 + seed + L)


@pytest.mark.parametrize("L", [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 33, 63, 65, 130])
def test_tiny_contigs(L):
    """1 ... 130 blocks per contig: every warp of a slab / item sees between zero and a few groups."""
    w = _workload(L)
    ref = random_model(np.random.default_rng(L), w.M, w.contigs)
    ctx, out = run_ctx(w.contigs, 1, ref, {"mma_min_chunks": 1, "chunk_blocks": 16, "burn_in_blocks": 512})
    check_against_port(out, w.contigs, ref)
    ctx.close()


@pytest.mark.parametrize("slab", [32, 36, 100, 1000])
def test_small_and_odd_slabs(slab):
    """Slab sizes that are no multiple of the group size: the partial group of every slab takes the masked path."""
    w = _workload(2311, contigs=3)
    ref = random_model(np.random.default_rng(slab), w.M, w.contigs)
    ctx, out = run_ctx(w.contigs, 1, ref, {"mma_min_chunks": 1, "chunk_blocks": 200, "burn_in_blocks": 512, "slab_blocks": slab})
    check_against_port(out, w.contigs, ref)
    ctx.close()


def test_one_block_chunks():
    """chunk_blocks = 1: one alpha_hat column pair per chunk (column index = 2 b), repaired by sweeps where the burn-in is short."""
    w = _workload(300, contigs=1)
    ref = random_model(np.random.default_rng(11), w.M, w.contigs)
    ctx, out = run_ctx(w.contigs, 1, ref, {"mma_min_chunks": 1, "chunk_blocks": 1, "burn_in_blocks": 512})
    check_against_port(out, w.contigs, ref)
    ctx.close()


@pytest.mark.parametrize("chunk", [2, 3, 7, 255, 257])
def test_column_index_for_odd_chunk_lengths(chunk):
    w = _workload(1500, contigs=2)
    ref = random_model(np.random.default_rng(chunk), w.M, w.contigs)
    ctx, out = run_ctx(w.contigs, 1, ref, {"mma_min_chunks": 1, "chunk_blocks": chunk, "burn_in_blocks": 512})
    check_against_port(out, w.contigs, ref)
    ctx.close()


def test_more_keys_than_fit_in_shared_memory():
    """> 128 distinct keys: the per-slab gamma sums accumulate in global memory (same code, same order)."""
    rng = np.random.default_rng(17)
    M, L = 32, 6000
    w = synth.make_workload("ring-keys", 2, L, M, 6, seed0=4242)
    contigs = []
    for c in w.contigs:
        c = c.copy()
        c[:, 1] = rng.integers(0, 3, size=L)           # a
        c[:, 2] = rng.integers(0, 60, size=L)          # b
        c[:, 3] = 60                                   # nb
        contigs.append(c)
    ref = random_model(rng, M, contigs)
    assert ref["keys"].shape[0] > 128
    ctx, out = run_ctx(contigs, 1, ref, {"mma_min_chunks": 1, "chunk_blocks": 500, "burn_in_blocks": 512})
    check_against_port(out, contigs, ref)
    ctx.close()


def test_all_span_one_and_all_span_above_one():
    """A contig without span>1 blocks (no work items) next to one without span-1 blocks (no span-1 groups in any slab)."""
    w = _workload(900, contigs=2)
    a, b = w.contigs[0].copy(), w.contigs[1].copy()
    a[:, 0] = 1
    b[:, 0] = np.maximum(b[:, 0], 2)
    ref = random_model(np.random.default_rng(3), w.M, [a, b])
    ctx, out = run_ctx([a, b], 1, ref, {"mma_min_chunks": 1, "chunk_blocks": 128, "burn_in_blocks": 512})
    check_against_port(out, [a, b], ref)
    ctx.close()

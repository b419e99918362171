"""Several GPUs in one process: smcpp_b200_multi_* (in-library contig sharding + one ncclAllReduce of the packed statistics).

With one visible GPU the handle runs on a single device (no NCCL is loaded); with two or more (gpurun --gpus 2) the
all-reduce path runs.  Either way the result must equal the single-context result and the golden vectors of the reference."""
import numpy as np
import pytest

from helpers import Golden, relmax, LL_RTOL, STAT_RTOL
from smcpp_b200 import capi

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("name", ["ragged", "ref_test_inference", "c4_twopop_1200"])
def test_multi_handle_matches_the_golden_vectors(name):
    g = Golden(name)
    ndev = max(1, min(n_gpus(), len(g.contigs), 8))
    mc = capi.MultiContext(list(range(ndev)))
    mc.set_contigs(g.contigs, g.npop)
    assert np.array_equal(mc.keys, g.ref["keys"])
    owned = sorted(int(c) for i in range(ndev) for c in mc.shard(i))
    assert owned == list(range(len(g.contigs)))
    out = mc.estep(g.ref["pi"], g.ref["T"], g.ref["E"])
    assert np.all(np.abs(out["ll"] - g.ref["ll"]) <= LL_RTOL * np.abs(g.ref["ll"]))
    for k in ("xisum", "gamma0", "gamma_sums"):
        for c in range(len(g.contigs)):
            assert relmax(out[k][c], g.ref[k][c]) <= STAT_RTOL, (k, c)
    assert np.array_equal(out["key_present"], g.ref["key_present"])
    # the all-reduced vector is the sum over ALL contigs, whatever the number of devices
    M, K = g.M, g.ref["keys"].shape[0]
    want = np.concatenate([[out["ll"].sum()], out["gamma0"].sum(0), out["xisum"].sum(0).ravel(), out["gamma_sums"].sum(0).ravel()])
    assert out["reduced"].shape == (1 + M + M * M + K * M,)
    assert np.allclose(out["reduced"], want, rtol=1e-12, atol=0)
    # per-device context: M-step objective of that device's contigs, summed over devices = the reference's Q
    q = sum(mc.context(i).q(g.ref["pi"], g.ref["T"], g.ref["E"]) for i in range(ndev) if len(mc.shard(i)))
    assert np.allclose(q, g.ref["Q"], rtol=1e-8, atol=0)
    mc.close()


def test_multi_handle_errors():
    with pytest.raises(RuntimeError, match="device"):
        capi.MultiContext([0, 99])
    mc = capi.MultiContext([0])
    with pytest.raises(RuntimeError, match="multi_set_contigs"):
        mc.estep(np.ones(4) / 4, np.eye(4), np.ones((0, 4)))
    bad = np.array([[1, -1, 0, 0], [0, 0, 0, 0]], np.int32)
    with pytest.raises(RuntimeError, match="span <= 0"):
        mc.set_contigs([bad], 1)
    mc.close()

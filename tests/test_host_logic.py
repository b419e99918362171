"""Host-side logic that needs no GPU: workload generator, bundles, sharding, key order, and the
world_size-2 all-reduce contract over gloo (the transport the multi-GPU path uses with NCCL)."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import Golden
from oracle import port
from smcpp_b200 import bundle, parallel, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synthetic_workload_is_deterministic_and_well_formed():
    a = synth.config("C2", 0.01).contigs[0]
    b = synth.config("C2", 0.01).contigs[0]
    assert np.array_equal(a, b)
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"] and a.shape[1] == 4
    assert tuple(a[0]) == (1, -1, 0, 0)                       # the pipeline's leading missing row
    assert (a[:, 0] > 0).all()
    assert (a[1::2, 0] >= 2).all() and (a[2::2, 0] == 1).all()  # alternating run / site rows
    assert ((a[:, 2] <= a[:, 3]) & (a[:, 3] <= 10)).all()
    # monomorphic rows never appear
    assert not (((a[:, 1] == 0) & (a[:, 2] == 0) & (a[:, 3] > 0)).any())
    assert not (((a[:, 1] == 2) & (a[:, 2] == a[:, 3]) & (a[:, 3] > 0)).any())
    h = hashlib.sha256(synth.config("C1").contigs[0].tobytes()).hexdigest()
    assert h == hashlib.sha256(synth.config("C1").contigs[0].tobytes()).hexdigest()
    w4 = synth.config("C4", 0.01)
    assert w4.contigs[0].shape[1] == 7 and w4.npop == 2 and w4.sfs.shape == (32, 3, 49)


def test_bundle_roundtrip(tmp_path):
    d = {"a": np.arange(6, dtype=np.int32).reshape(2, 3), "b": np.float64(3.5), "c": np.zeros((0, 4), np.float32),
         "d": np.array([1, 2, 3], np.uint8)}
    p = tmp_path / "x.smcb"
    bundle.save(p, d)
    r = bundle.load(p)
    assert np.array_equal(r["a"], d["a"]) and r["b"][0] == 3.5 and r["c"].shape == (0, 4) and r["d"].dtype == np.uint8


def test_key_order_is_the_references_map_order():
    g = Golden("c4_twopop_1200")
    assert np.array_equal(parallel.local_keys(g.contigs), g.ref["keys"])
    g = Golden("ragged")
    assert np.array_equal(parallel.local_keys(g.contigs), g.ref["keys"])


def test_shard_contigs_lpt():
    s = parallel.shard_contigs([10] * 22, 8)
    assert sorted(len(x) for x in s) == [2, 2, 3, 3, 3, 3, 3, 3]          # 3,3,3,3,3,3,2,2
    assert sorted(sum(s, [])) == list(range(22))
    s = parallel.shard_contigs([100, 1, 1, 1, 50, 49], 2)
    loads = [sum([100, 1, 1, 1, 50, 49][i] for i in r) for r in s]
    assert max(loads) <= 102
    assert parallel.shard_contigs([5, 4], 4) == [[0], [1], [], []]


def test_pack_unpack_reduced():
    rng = np.random.default_rng(1)
    C, M, K = 3, 5, 4
    ll, g0, xi, gs = rng.random(C), rng.random((C, M)), rng.random((C, M, M)), rng.random((C, K, M))
    v = parallel.pack_reduced(ll, g0, xi, gs)
    u = parallel.unpack_reduced(v, M, K)
    assert u["ll"] == pytest.approx(ll.sum()) and np.allclose(u["xisum"], xi.sum(0)) and np.allclose(u["gamma_sums"], gs.sum(0))


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from helpers import Golden
from oracle import port
from smcpp_b200 import parallel
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
g = Golden("ragged")
owned = parallel.shard_contigs([c.shape[0] for c in g.contigs], 2)[rank]
mine = [g.contigs[i] for i in owned]
keys = parallel.union_keys(parallel.local_keys(mine))
assert np.array_equal(keys, g.ref["keys"]), "union of shard key tables = global table"
M, K = g.M, keys.shape[0]
# the compute here is the oracle port (no GPU in this test); what is under test is shard -> pack -> all-reduce
outs = [port.hmm_estep(c, g.ref) for c in mine]
vec = parallel.pack_reduced([o["ll"] for o in outs], [o["gamma0"] for o in outs], [o["xisum"] for o in outs],
                            [o["gamma_sums"] for o in outs])
t = torch.from_numpy(vec.copy())
parallel.allreduce_sum_(t)
full = parallel.pack_reduced(g.ref["ll"], g.ref["gamma0"], g.ref["xisum"], g.ref["gamma_sums"])
err = float(np.abs(t.numpy() - full).max() / np.abs(full).max())
assert err < 1e-12, err
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", err)
"""


def test_two_rank_gloo_allreduce_of_packed_statistics(tmp_path):
    script = tmp_path / "worker.py"
    port_no = 29500 + (os.getpid() % 400)
    script.write_text(_WORKER.format(root=ROOT, port=port_no))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o

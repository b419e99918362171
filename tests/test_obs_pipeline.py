"""Input side of the E-step: thin_data -> bin_observations -> RecodeMonomorphic -> compress_repeated_obs.

CPU part: the oracle restatement (oracle/obs_port.c) against what the reference's own functions produced
(tests/golden/obs_pipeline.npz, made by tests/golden/make_obs_golden.py from the reference's Cython / Python sources).
GPU part: the CUDA pipeline (smcpp_b200/csrc/obs_pipeline.cu through the C ABI) against the goldens and, on larger fresh
inputs, against the oracle -- bit for bit (integer work)."""
import os

import numpy as np
import pytest

from oracle import obsport

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obs_pipeline.npz"))
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


def test_oracle_invariants():
    g, npop, thinning, w = case("p1")
    thin = obsport.thin_data(g["raw"], thinning)
    assert thin[:, 0].astype(np.int64).sum() == g["raw"][:, 0].astype(np.int64).sum()        # reference's own assert (:82)
    comp = obsport.compress_repeated_obs(thin)
    assert comp[:, 0].astype(np.int64).sum() == thin[:, 0].astype(np.int64).sum()
    assert (np.abs(np.diff(comp[:, 1:], axis=0)).sum(axis=1) > 0).all()                    # no two neighbours share a key

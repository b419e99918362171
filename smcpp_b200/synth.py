"""Deterministic synthetic workloads of SURVEY.md section 8(d) (the BASELINE.json configs).

Everything here is *input generation*: observation rows in the reference's int32 layout
``[span, a1, b1, nb1 (, a2, b2, nb2)]`` (reference README.rst:519-566, smcpp/_smcpp.pyx:133-151),
the piecewise-constant size history, hidden-state boundaries and a synthetic conditioned SFS that the
reference consumes through its own ``DummySFS`` backend (reference include/conditioned_sfs.h:46-67).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Workload:
    name: str
    npop: int
    n: tuple            # undistinguished sample size per population
    na: tuple           # distinguished lineages per population
    M: int
    hidden_states: np.ndarray      # [M+1], last = inf
    contigs: list                  # list of int32 [L, 1+3P], C-contiguous
    model_a: np.ndarray
    model_s: np.ndarray
    theta: float = 2.5e-3
    rho: float = 1e-3
    alpha: float = 1.0
    pol_err: float = 0.0
    sfs: np.ndarray = field(default=None)   # [M, 3, dim]
    overrides: dict = field(default_factory=dict)   # test harness only: override_T / override_E / override_pi

    @property
    def total_blocks(self) -> int:
        return int(sum(c.shape[0] for c in self.contigs))

    def to_bundle(self, save_gamma=False, dump_alpha=False) -> dict:
        return {
            "npop": np.int32(self.npop),
            "n": np.asarray(self.n, np.int32),
            "na": np.asarray(self.na, np.int32),
            "hidden_states": np.asarray(self.hidden_states, np.float64),
            "contig_lengths": np.asarray([c.shape[0] for c in self.contigs], np.int32),
            "obs": np.concatenate(self.contigs, axis=0).astype(np.int32),
            "model_a": self.model_a, "model_s": self.model_s,
            "theta": np.float64(self.theta), "rho": np.float64(self.rho),
            "alpha": np.float64(self.alpha), "pol_err": np.float64(self.pol_err),
            "sfs": self.sfs,
            "save_gamma": np.int32(bool(save_gamma)), "dump_alpha": np.int32(bool(dump_alpha)),
            **{k: np.asarray(v, np.float64) for k, v in self.overrides.items()},
        }


def hidden_states(M: int) -> np.ndarray:
    """[0, logspace(-2, 1, M-1), inf]  (same recipe as the survey probe harness)."""
    hs = np.empty(M + 1)
    hs[0] = 0.0
    if M > 1:
        m = np.arange(1, M)
        hs[1:M] = 0.01 * (10.0 / 0.01) ** ((m - 1.0) / max(M - 2.0, 1.0))
    hs[M] = np.inf
    return hs


def model(K: int = 24):
    """24 log-spaced pieces on [0.002, 5], a_k = 1 + 0.8 sin(0.7 k)."""
    k = np.arange(K)
    a = 1.0 + 0.8 * np.sin(0.7 * k)
    t0 = 0.002 * (5.0 / 0.002) ** (k / K)
    t1 = 0.002 * (5.0 / 0.002) ** ((k + 1) / K)
    s = np.where(k == 0, t1, t1 - t0)
    return a.astype(np.float64), s.astype(np.float64)


def dummy_sfs(hs: np.ndarray, n: tuple, seed: int = 7) -> np.ndarray:
    """Seeded positive 3 x dim CSFS per hidden state, scaled by the state's time.

    dim = prod(n_p + 1); entries (a=0, all b=0) and (a=2, all b=n) are zero as the reference asserts
    (reference src/conditioned_sfs.cpp:108-109)."""
    M = len(hs) - 1
    rng = np.random.default_rng(seed)
    dims = tuple(int(x) + 1 for x in n)
    dim = int(np.prod(dims))
    out = np.zeros((M, 3, dim))
    bsum = np.add.outer(np.arange(dims[0]), np.arange(dims[1])).reshape(-1) if len(dims) == 2 else np.arange(dims[0])
    for m in range(M):
        hi = 2.0 * hs[m] if np.isinf(hs[m + 1]) else hs[m + 1]
        t = 0.5 * (hs[m] + hi)
        if t == 0.0:
            t = 0.5 * hs[1] if M > 1 else 1.0
        u = rng.uniform(0.05, 1.0, size=(3, dim))
        out[m] = t * u / (1.0 + np.arange(3)[:, None] + bsum[None, :])
        out[m, 0, 0] = 0.0
        out[m, 2, dim - 1] = 0.0
    return out


def make_contig(L: int, n: tuple, seed: int, npop: int = 1, mean_run: float = 500.0) -> np.ndarray:
    """One contig of L rows following SURVEY 8(d): row 0 missing, then alternating run / site rows."""
    rng = np.random.default_rng(seed)
    W = 1 + 3 * npop
    obs = np.zeros((L, W), np.int32)
    obs[:, 0] = 1
    obs[0, 1] = -1
    if npop == 2:
        obs[0, 4] = -1
    idx = np.arange(1, L)
    run = idx[(idx % 2) == 1]
    site = idx[(idx % 2) == 0]
    # run rows
    spans = np.minimum(2 + rng.geometric(1.0 / mean_run, size=run.size), 50000)
    obs[run, 0] = spans
    miss = (np.arange(run.size) % 1000) == 999
    obs[run[miss], 1] = -1
    # site rows
    full = rng.random(site.size) < 0.125
    ared = rng.choice(np.array([1, 2, -1]), size=site.size, p=[0.90, 0.05, 0.05])
    obs[site, 1] = ared
    nf = int(full.sum())
    if nf:
        a = rng.integers(0, 3, size=nf)
        bs = [rng.integers(0, int(n[p]) + 1, size=nf) for p in range(npop)]
        for _ in range(64):
            all0 = np.all([b == 0 for b in bs], axis=0)
            alln = np.all([bs[p] == n[p] for p in range(npop)], axis=0)
            bad = ((a == 0) & all0) | ((a == 2) & alln)
            if not bad.any():
                break
            nb = int(bad.sum())
            a[bad] = rng.integers(0, 3, size=nb)
            for p in range(npop):
                bs[p][bad] = rng.integers(0, int(n[p]) + 1, size=nb)
        rows = site[full]
        obs[rows, 1] = a
        for p in range(npop):
            obs[rows, 2 + 3 * p] = bs[p]
            obs[rows, 3 + 3 * p] = n[p]
    return np.ascontiguousarray(obs)


def make_workload(name: str, contigs: int, L: int, M: int, n, npop: int = 1, seed0: int = 1000) -> Workload:
    n = tuple(n) if isinstance(n, (tuple, list)) else (int(n),) * npop
    na = (2,) if npop == 1 else (2, 0)
    hs = hidden_states(M)
    a, s = model()
    cs = [make_contig(L, n, seed0 + c, npop) for c in range(contigs)]
    return Workload(name=name, npop=npop, n=n, na=na, M=M, hidden_states=hs, contigs=cs,
                    model_a=a, model_s=s, sfs=dummy_sfs(hs, n))


# BASELINE.json configs (C1..C5); `scale` shrinks L for parity-test sized versions
def config(which: str, scale: float = 1.0) -> Workload:
    def LL(x):
        return max(int(round(x * scale)), 8)
    if which == "C1":
        return make_workload("C1", 1, LL(10_000), 16, 4)
    if which == "C2":
        return make_workload("C2", 1, LL(1_000_000), 32, 10)
    if which == "C3":
        return make_workload("C3", 22, LL(1_000_000), 32, 20)
    if which == "C4":
        return make_workload("C4", 2, LL(500_000), 32, (6, 6), npop=2)
    if which.startswith("C5-"):
        return make_workload(which, 1, LL(1_000_000), int(which[3:]), 10)
    raise KeyError(which)

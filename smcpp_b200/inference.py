"""Host-side mirror of the reference's inference-manager interface for the E-step path.

Names and meanings follow the reference's Cython class `_PyInferenceManager`
(smcpp/_smcpp.pyx:97-308) and C++ `InferenceManager` (include/inference_manager.h:18-82):
`E_step()`, `loglik()`, `xisums`, `gammas`, `gamma_sums`, `save_gamma`, plus the observation/hidden-state
constructor arguments.  Everything numeric is done by libsmcpp_b200.so on the GPU through the C ABI
(include/smcpp_b200.h); this class only owns buffers, shards contigs over GPUs and reshapes outputs.

What is *not* here (out of the hot path): the CSFS and the autodiff'd M-step.  The inputs
the reference derives from its model before the forward-backward -- pi, the transition matrix and the
per-key emission vectors (do_dirty_work, src/inference_manager.cpp:213-229) -- are handed in with
`set_hmm_inputs`.
"""
from __future__ import annotations

import threading

import numpy as np

from . import capi, parallel


class InferenceManager:
    """E-step engine over a list of contigs.

    observations : list of int32 arrays [L, 1+3P]  (reference smcpp/_smcpp.pyx:133-151)
    hidden_states: M+1 boundaries, first 0, last inf (reference include/inference_manager.h:41)
    devices      : CUDA device indices of THIS process; contigs are sharded over them
                   (one context per device, driven from one host thread each).
    """

    def __init__(self, observations, hidden_states, npop: int = 1, devices=(0,), keys=None):
        self.hidden_states = np.asarray(hidden_states, np.float64)
        self.M = len(self.hidden_states) - 1
        self.npop = int(npop)
        self._obs = [np.ascontiguousarray(o, np.int32) for o in observations]
        for ob in self._obs:
            if ob.ndim != 2 or ob.shape[1] != 1 + 3 * self.npop:
                raise RuntimeError("observations must be int32 arrays of shape [L, 1 + 3*npop]")
            if (ob[:, 0] <= 0).any():
                raise RuntimeError("data are malformed: span <= 0")   # reference src/inference_manager.cpp:243-244
        self.devices = list(devices)
        self.keys = parallel.sort_keys(keys) if keys is not None else parallel.local_keys(self._obs)
        self.K = self.keys.shape[0]
        lengths = [o.shape[0] for o in self._obs]
        self._shards = parallel.shard_contigs(lengths, len(self.devices))
        self._ctx = []
        for dev, idx in zip(self.devices, self._shards):
            if not idx:
                self._ctx.append(None)
                continue
            c = capi.Context(dev)
            c.set_contigs([self._obs[i] for i in idx], self.npop, self.keys)
            self._ctx.append(c)
        live = [c for c in self._ctx if c is not None]
        self.eig_keys = np.unique(np.concatenate([c.eig_keys for c in live])) if live else np.zeros(0, np.int32)
        self.save_gamma = False
        self._inputs = None
        self._out = None

    # -- inputs of the forward-backward (what do_dirty_work() leaves behind in the reference)
    def set_hmm_inputs(self, pi, transition, emission_probs, eigensystems: dict | None = None):
        pi = np.asarray(pi, np.float64)
        T = np.asarray(transition, np.float64)
        E = np.asarray(emission_probs, np.float64)
        if pi.shape != (self.M,) or T.shape != (self.M, self.M) or E.shape != (self.K, self.M):
            raise RuntimeError("set_hmm_inputs: expected pi[M], transition[M,M], emission_probs[K,M]")
        self._inputs = (pi, T, E, eigensystems)

    def set_model(self, model_a, model_s, theta, rho, alpha, sfs, n, na, polarization_error: float = 0.0):
        """Build pi / transition / emission table from the model like the reference's do_dirty_work()
        (src/inference_manager.cpp:213-229), through the library's host routines (include/smcpp_b200.h:
        smcpp_b200_host_initial_distribution / _transition / _emission).

        model_a, model_s : piecewise-constant sizes and piece lengths (setParams, smcpp/_smcpp.pyx:205-221)
        theta, rho, alpha: setTheta / setRho / setAlpha (src/inference_manager.cpp:71-87)
        sfs              : conditioned SFS per hidden state, [M, na[0]+1, sfs_dim] -- an INPUT of this path
                           (the reference computes it in OnePopConditionedSFS / JointCSFS, out of scope here)."""
        mi = capi.host_model_inputs(self.hidden_states, model_a, model_s, theta, rho, alpha, polarization_error, sfs, n, na,
                                    self.keys)
        self.set_hmm_inputs(mi["pi"], mi["T"], mi["E"])
        return mi

    def Q(self, dpi=None, dT=None, dE=None):
        """InferenceManager::Q() (reference src/inference_manager.cpp:116-126, src/hmm.cpp:155-193) on the device
        (smcpp_b200_q): [sum log(pi) gamma0, sum_{nb=0 keys} log(e) gamma_sums, sum_{nb>0 keys} ..., sum log(T) xisum], per contig
        in the reference's order with its doubly compensated summation, keys present in the contig only, summed over
        contigs (and over this process's devices).  A present key with a non-positive emission entry makes its class
        -inf (the reference warns there).  With the derivative arrays dpi[D,M], dT[D,M,M], dE[D,K,M] of the current
        inputs the gradient dq[4,D] is returned as well (the reference evaluates Q on autodiff scalars).

        Before the first E-step the reference's HMM constructor has pre-filled gamma_sums with span * pi
        (src/hmm.cpp:16-27) and zeroed the rest; so do we (set_statistics)."""
        if self._inputs is None:
            raise RuntimeError("Q: set_hmm_inputs() / set_model() has not been called")
        pi, T, E, _ = self._inputs
        if self._out is None:
            lut = {tuple(int(v) for v in k): i for i, k in enumerate(self.keys)}
            for ctx, idx in zip(self._ctx, self._shards):
                if ctx is None:
                    continue
                gs = np.zeros((len(idx), self.K, self.M))
                for local, glob in enumerate(idx):
                    ob = self._obs[glob]
                    uniq, inv = np.unique(ob[:, 1:], axis=0, return_inverse=True)
                    tot = np.bincount(inv.reshape(-1), weights=ob[:, 0].astype(np.float64), minlength=len(uniq))
                    for u, t in zip(uniq, tot):
                        gs[local, lut[tuple(int(v) for v in u)]] += t * pi
                ctx.set_statistics(np.zeros((len(idx), self.M, self.M)), np.zeros((len(idx), self.M)), gs)
        q = np.zeros(4)
        dq = None
        for ctx in self._ctx:
            if ctx is None:
                continue
            r = ctx.q(pi, T, E, dpi, dT, dE)
            if dpi is None:
                q += r
            else:
                q += r[0]
                dq = r[1] if dq is None else dq + r[1]
        return q if dpi is None else (q, dq)

    def set_option(self, name, value):
        for c in self._ctx:
            if c is not None:
                c.set_option(name, value)

    def _eig_for(self, ctx, T, E, eig):
        """Eigensystems in the order of ctx.eig_keys (a shard may see a subset of the global eigen keys)."""
        if eig is None:
            return None
        pos = {int(k): i for i, k in enumerate(eig["eig_key_idx"])}
        sel = [pos[int(k)] for k in ctx.eig_keys]
        return {k: np.ascontiguousarray(np.asarray(eig[k])[sel]) for k in ("eig_P", "eig_Pinv", "eig_d", "eig_dscaled", "eig_scale")}

    def E_step(self, forward_backward_only: bool = False):
        """Reference: _PyInferenceManager.E_step (smcpp/_smcpp.pyx:185-191) -> InferenceManager::Estep."""
        if self._inputs is None:
            raise RuntimeError("E_step: set_hmm_inputs() has not been called")
        pi, T, E, eig = self._inputs
        for c in self._ctx:
            if c is not None:
                c.set_save_gamma(bool(self.save_gamma))
        results = [None] * len(self._ctx)
        errors = []

        def run(i, ctx):
            try:
                results[i] = ctx.estep(pi, T, E, self._eig_for(ctx, T, E, eig))
            except Exception as ex:  # surfaced on the calling thread below
                errors.append(ex)

        threads = [threading.Thread(target=run, args=(i, c)) for i, c in enumerate(self._ctx) if c is not None]
        if len(threads) == 1:
            threads[0].run()
        else:
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if errors:
            raise errors[0]
        C = len(self._obs)
        out = {"ll": np.zeros(C), "xisum": np.zeros((C, self.M, self.M)), "gamma0": np.zeros((C, self.M)),
               "gamma_sums": np.zeros((C, self.K, self.M)), "key_present": np.zeros((C, self.K), np.uint8)}
        red = np.zeros(1 + self.M + self.M * self.M + self.K * self.M)
        for idx, r in zip(self._shards, results):
            if r is None:
                continue
            for local, glob in enumerate(idx):
                for k in ("ll", "xisum", "gamma0", "gamma_sums", "key_present"):
                    out[k][glob] = r[k][local]
            red += r["reduced"]
        out["reduced"] = red
        self._out = out

    # -- outputs, shaped like the reference's Python properties
    def _need(self):
        if self._out is None:
            raise RuntimeError("E_step() has not been run")
        return self._out

    def loglik(self) -> float:
        """Reference: sum of InferenceManager::loglik() (smcpp/_smcpp.pyx:303-308)."""
        return float(self._need()["ll"].sum())

    @property
    def logliks(self):
        return self._need()["ll"].copy()

    @property
    def xisums(self):
        """Reference: _PyInferenceManager.xisums (smcpp/_smcpp.pyx:257-263): one M x M matrix per contig."""
        return [x.copy() for x in self._need()["xisum"]]

    @property
    def gammas(self):
        """Reference: _PyInferenceManager.gammas (smcpp/_smcpp.pyx:233-239): per contig an M x (L+1) matrix with
        save_gamma (the `smc++ posterior` path), else only column 0 (src/hmm.cpp:150), returned as [M, 1]."""
        o = self._need()
        if not self.save_gamma:
            return [g.reshape(-1, 1).copy() for g in o["gamma0"]]
        ret = [None] * len(self._obs)
        for ctx, idx in zip(self._ctx, self._shards):
            if ctx is None:
                continue
            for local, glob in enumerate(idx):
                ret[glob] = ctx.fetch_gamma(local).T
        return ret

    @property
    def gamma_sums(self):
        """Reference: _PyInferenceManager.gamma_sums (smcpp/_smcpp.pyx:241-255): per contig a dict
        {key tuple: vector[M]} holding exactly the keys present in that contig."""
        o = self._need()
        ret = []
        for c in range(len(self._obs)):
            ret.append({tuple(int(v) for v in self.keys[k]): o["gamma_sums"][c, k].copy()
                        for k in range(self.K) if o["key_present"][c, k]})
        return ret

    @property
    def reduced(self):
        """[ll | gamma0 | xisum | gamma_sums] summed over contigs: the all-reduce payload."""
        return self._need()["reduced"].copy()

    def stats(self):
        return [c.stats() if c is not None else None for c in self._ctx]

    def close(self):
        for c in self._ctx:
            if c is not None:
                c.close()
        self._ctx = []

"""Throughput of the device-resident observation pre-processing (thin -> bin -> recode -> compress) on a chromosome-sized
contig, against the oracle's sequential CPU restatement of the reference functions on a bounded sample."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from smcpp_b200 import capi
from test_obs_pipeline import _raw_rows

L = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
n, a, thinning, w = (20,), (2,), 1521, 100
raw = _raw_rows(np.random.default_rng(5), L, 1, n, a)
W = raw.shape[1]
res = {"rows_in": L, "bases": int(raw[:, 0].astype(np.int64).sum())}
p = capi.ObsPipeline(raw)
for rep in range(3):          # the first pass allocates the device buffers; later passes reuse them
    p.upload(raw)
    sizes = [L]
    p.thin(thinning); sizes.append(p.n_rows)
    p.bin(a, w); sizes.append(p.n_rows)
    p.recode_monomorphic(a); sizes.append(p.n_rows)
    p.compress(); sizes.append(p.n_rows)
    ms = dict(p.ms)
p.close()
# algorithmic bytes: rows read + rows written (4 W bytes each) per step; the prefix sums add 8 B/row read + written per scan
by = {"thin": 4 * W * (sizes[0] + sizes[1]), "bin": 4 * W * (sizes[1] + sizes[2]), "recode_monomorphic": 2 * 4 * W * sizes[2],
      "compress": 4 * W * (sizes[3] + sizes[4])}
res["steps"] = {k: {"ms": ms[k], "alg_GBps": by[k] / ms[k] / 1e6} for k in ms}
res["rows"] = sizes
res["total_ms"] = sum(ms.values())
# CPU baseline (measurement only, like bench.py's cpu_baseline leg): the oracle's sequential restatement of the reference
# functions, one host core, on a bounded sample of the same rows
from oracle import obsport
sample = raw[:min(L, 2_000_000)]
t0 = time.perf_counter()
t = obsport.thin_data(sample, thinning); b = obsport.bin_observations(t, a, w); r = obsport.recode_monomorphic(b, a); c = obsport.compress_repeated_obs(r)
res["cpu_port_rows_per_s"] = sample.shape[0] / (time.perf_counter() - t0)
res["gpu_rows_per_s"] = L / (res["total_ms"] * 1e-3)
print(json.dumps(res))

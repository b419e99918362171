"""Generates tests/golden/obs/obs_pipeline.npz: seeded raw observation rows and what the REFERENCE makes of them in the
`smc++ estimate` pre-processing chain (smcpp/analysis/analysis.py:60-63: Thin -> BinObservations -> RecodeMonomorphic
-> Compress).  Run in the build container only (needs /root/reference):

    python tests/golden/make_obs_golden.py

thin_data / bin_observations are Cython (smcpp/_estimation_tools.pyx): the file is compiled unmodified in a scratch
directory except that `beta_de_avg_pdf` -- the one function that needs the GSL header, not on this path -- is cut off.
compress_repeated_obs (smcpp/estimation_tools.py:51-61) and RecodeMonomorphic._recode (smcpp/data_filter.py:331-336) are
pure Python / NumPy and are executed from the reference's source text (importing the modules would pull in the full
package).  Nothing of the reference is copied into this repository."""
import ast
import os
import subprocess
import sys
import tempfile
import textwrap

import numpy as np

REF = "/root/reference/smcpp"
HERE = os.path.dirname(os.path.abspath(__file__))


def build_reference_cython():
    d = tempfile.mkdtemp(prefix="smcpp_ref_et_")
    src = open(os.path.join(REF, "_estimation_tools.pyx")).read()
    src = src.replace('cdef extern from "<gsl/gsl_sf_gamma.h>":\n    double gsl_sf_lnbeta(double, double) nogil\n', "")
    src = src[:src.index("def beta_de_avg_pdf")]
    open(os.path.join(d, "ref_estimation_tools.pyx"), "w").write(src)
    open(os.path.join(d, "setup.py"), "w").write(textwrap.dedent("""
        from setuptools import setup, Extension
        from Cython.Build import cythonize
        import numpy
        setup(ext_modules=cythonize([Extension("ref_estimation_tools", ["ref_estimation_tools.pyx"],
                                               include_dirs=[numpy.get_include()])], language_level=3))
    """))
    subprocess.check_call([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=d, stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    sys.path.insert(0, d)
    import ref_estimation_tools
    return ref_estimation_tools


def reference_function(path, name, cls=None):
    """Compile one function (or method) out of a reference source file."""
    tree = ast.parse(open(path).read())
    body = tree.body
    if cls:
        body = next(n for n in body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == name)
    fn.decorator_list = []
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np}
    exec(compile(mod, path, "exec"), ns)
    return ns[name]


class FakeContig:   # what bin_observations / _recode touch: .data, .a, len()
    def __init__(self, data, a):
        self.data = data
        self.a = np.asarray(a, np.int64)

    def __len__(self):   # reference smcpp/contig.py: total base pairs
        return int(self.data[:, 0].sum())


def raw_rows(rng, L, npop, n, a, long_runs=True):
    """Raw .smc-like rows: monomorphic / missing runs with occasional segregating or full-SFS sites."""
    W = 1 + 3 * npop
    d = np.zeros((L, W), np.int32)
    for l in range(L):
        u = rng.random()
        if u < 0.45:      # run of non-segregating bases, full sample observed
            d[l, 0] = rng.geometric(1 / 300.0) if long_runs else rng.integers(1, 6)
            for p in range(npop):
                d[l, 1 + 3 * p:4 + 3 * p] = [0, 0, n[p]]
        elif u < 0.55:    # missing run
            d[l, 0] = rng.geometric(1 / 150.0) if long_runs else rng.integers(1, 4)
            for p in range(npop):
                d[l, 1 + 3 * p:4 + 3 * p] = [-1, 0, 0]
        elif u < 0.62:    # fully derived site (a = a_p, b = nb): RecodeMonomorphic's target
            d[l, 0] = 1
            for p in range(npop):
                nb = rng.integers(0, n[p] + 1)
                d[l, 1 + 3 * p:4 + 3 * p] = [a[p], nb, nb]
        else:             # segregating site, partly missing undistinguished sample
            d[l, 0] = rng.integers(1, 3)
            for p in range(npop):
                nb = rng.integers(0, n[p] + 1)
                d[l, 1 + 3 * p:4 + 3 * p] = [rng.integers(-1 if rng.random() < 0.1 else 0, a[p] + 1), rng.integers(0, nb + 1), nb]
                if d[l, 1 + 3 * p] < 0:
                    d[l, 2 + 3 * p] = 0
    return d


def main():
    et = build_reference_cython()
    compress = reference_function(os.path.join(REF, "estimation_tools.py"), "compress_repeated_obs")
    recode = reference_function(os.path.join(REF, "data_filter.py"), "_recode", cls="RecodeMonomorphic")
    recode_nonseg = reference_function(os.path.join(REF, "estimation_tools.py"), "recode_nonseg")
    break_long_spans = reference_function(os.path.join(REF, "estimation_tools.py"), "break_long_spans")
    contig_mod = {}
    exec(compile(open(os.path.join(REF, "contig.py")).read(), os.path.join(REF, "contig.py"), "exec"), contig_mod)
    import logging
    for f in (recode_nonseg, break_long_spans):     # module globals the two functions use
        f.__globals__["Contig"] = contig_mod["Contig"]
        f.__globals__["logger"] = logging.getLogger("ref")
    out = {}
    cases = [("p1", 1, (8,), (2,), 4000, 37, 100, True), ("p1_dense", 1, (5,), (2,), 3000, 7, 10, False),
             ("p1_w1000", 1, (20,), (2,), 5000, 1521, 1000, True), ("p2_20", 2, (6, 4), (2, 0), 3000, 53, 100, True),
             ("p2_11", 2, (3, 5), (1, 1), 3000, 2, 25, False), ("exact_fill", 1, (4,), (2,), 0, 10, 10, True)]
    names = []
    for name, npop, n, a, L, thinning, w, long_runs in cases:
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        if name == "exact_fill":   # spans that land exactly on bin / thinning boundaries
            d = np.array([[10, 0, 0, 4], [1, 1, 2, 4], [9, 0, 0, 4], [20, -1, 0, 0], [5, 0, 0, 4], [5, 2, 4, 4], [1, 1, 0, 0], [29, 0, 0, 4],
                          [10, 0, 0, 0]], np.int32)
        else:
            d = raw_rows(rng, L, npop, n, a, long_runs)
        # base.py:50-52 front of the chain on the raw rows: RecodeNonseg(cutoff) -> Compress -> BreakLongSpans(cutoff)
        cut_ns, cut_bl = (400, 300) if long_runs else (4, 3)
        c0 = contig_mod["Contig"](pid=("p",) * npop, data=d.copy(), n=n, a=a, fn="golden")
        recode_nonseg(c0, cut_ns)
        out[f"{name}__nonseg"] = c0.data.copy()
        c0.data = compress(c0.data)
        pieces = break_long_spans(c0, cut_bl)
        out[f"{name}__pieces"] = np.concatenate([p.data for p in pieces], axis=0)
        out[f"{name}__piece_len"] = np.array([p.data.shape[0] for p in pieces], np.int64)
        out[f"{name}__cutoffs"] = np.array([cut_ns, cut_bl], np.int64)
        thin = et.thin_data(d.copy(), thinning)
        c = FakeContig(thin.copy(), a)
        binned = np.array(et.bin_observations(c, w))
        c2 = FakeContig(binned.copy(), a)
        recode(None, c2)
        comp = compress(c2.data.copy())
        comp_raw = compress(d.copy())
        for k, v in (("raw", d), ("thin", thin), ("binned", binned), ("recoded", c2.data), ("compressed", comp), ("raw_compressed", comp_raw),
                     ("params", np.array([npop, thinning, w], np.int64)), ("a", np.asarray(a, np.int64)), ("n", np.asarray(n, np.int64))):
            out[f"{name}__{k}"] = np.asarray(v)
        names.append(name)
        print(name, d.shape, "->", thin.shape, "->", binned.shape, "->", comp.shape)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "obs", "obs_pipeline.npz"), **out)


if __name__ == "__main__":
    main()

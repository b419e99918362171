// TEST INFRASTRUCTURE (oracle build only) -- not part of the product, never linked into it.
//
// Drives the UNMODIFIED reference implementation of the E-step path, compiled from the sources
// where they lie under /root/reference (see Makefile in this directory), and dumps every
// intermediate and output of the path into an SMCB1 bundle:
//   InferenceManager::Estep            reference src/inference_manager.cpp:108-114
//   -> do_dirty_work                   :213-229   (pi, emission table, transition)
//   -> TransitionBundle::update        reference src/transition_bundle.cpp:3-61
//   -> HMM::Estep                      reference src/hmm.cpp:45-153
// The CSFS (an *input* of the path) is supplied through the reference's own DummySFS
// (reference include/conditioned_sfs.h:46-67).
//
// usage: ref_harness <in.smcb> <out.smcb> [threads] [repeat]
#include <chrono>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <utility>
#include <vector>
#include <omp.h>

#include "common.h"
#include "sparsepp/spp.h"
#include "block_key.h"
#include "piecewise_constant_rate_function.h"
#include "conditioned_sfs.h"
#include "transition.h"
// open up the reference classes so the harness can read their state (layout is unaffected)
#define private public
#define protected public
#include "transition_bundle.h"
#include "inference_bundle.h"
#include "hmm.h"
#include "inference_manager.h"
#undef private
#undef protected

#include "bundle_io.h"

static void quiet_logger(const std::string name, const std::string level, const std::string msg)
{
    if (std::getenv("SMCB_REF_VERBOSE")) std::fprintf(stderr, "[%s %s] %.300s\n", name.c_str(), level.c_str(), msg.c_str());
}

template <size_t P>
static void dump_bins(smcb::Bundle &out, NPopInferenceManager<P> &im, const std::vector<block_key> &keys)
{
    std::vector<int32_t> kidx, cells;
    std::vector<double> w;
    for (size_t k = 0; k < keys.size(); ++k) {
        auto it = im.bins.find(keys[k]);
        if (it == im.bins.end()) continue;
        for (const auto &p : it->second) {
            kidx.push_back((int32_t)k);
            for (int q = 0; q < p.first.size(); ++q) cells.push_back(p.first(q));
            w.push_back(p.second);
        }
    }
    out.put_i4("bins_key_idx", {(long)kidx.size()}, kidx.data());
    out.put_i4("bins_cell", {(long)kidx.size(), (long)(2 * P)}, cells.data());
    out.put_f8("bins_w", {(long)w.size()}, w.data());
}

static int run(const smcb::Bundle &in, smcb::Bundle &out, int threads, int repeat)
{
    const int P = in.scalar_i4("npop");
    const int32_t *n = in.get("n").as<int32_t>();
    const int32_t *na = in.get("na").as<int32_t>();
    const smcb::Array &hsA = in.get("hidden_states");
    std::vector<double> hs(hsA.as<double>(), hsA.as<double>() + hsA.count());
    const int M = (int)hs.size() - 1;
    const smcb::Array &LA = in.get("contig_lengths");
    const int C = (int)LA.count();
    std::vector<int> Ls(LA.as<int32_t>(), LA.as<int32_t>() + C);
    const int width = 1 + 3 * P;
    // the reference keeps raw int* into caller-owned storage (reference smcpp/_smcpp.pyx:133-151)
    std::vector<int> obs_store(in.get("obs").as<int32_t>(), in.get("obs").as<int32_t>() + in.get("obs").count());
    std::vector<int *> obs_ptrs;
    {
        long off = 0;
        for (int c = 0; c < C; ++c) { obs_ptrs.push_back(obs_store.data() + off * width); off += Ls[c]; }
    }
    const double theta = in.scalar_f8("theta"), rho = in.scalar_f8("rho"), alpha = in.scalar_f8("alpha");
    const double pol_err = in.scalar_f8("pol_err");
    const int save_gamma = in.scalar_i4("save_gamma", 0);
    const int dump_alpha = in.scalar_i4("dump_alpha", 0);

    // CSFS input: sfs[M][na0+1][sfs_dim] row-major
    const smcb::Array &sfsA = in.get("sfs");
    if (sfsA.shape.size() != 3 || sfsA.shape[0] != M || sfsA.shape[1] != 3)
        throw std::runtime_error("sfs must be [M,3,dim] (DummySFS stores 3 x dim matrices)");
    const int dim = (int)sfsA.shape[2];
    std::vector<double> sfs_store(sfsA.as<double>(), sfsA.as<double>() + sfsA.count());
    std::vector<double *> sfs_ptrs;
    for (int m = 0; m < M; ++m) sfs_ptrs.push_back(sfs_store.data() + (long)m * 3 * dim);

    const smcb::Array &aA = in.get("model_a"), &sA = in.get("model_s");
    std::vector<adouble> pa, ps;
    for (long k = 0; k < aA.count(); ++k) { pa.push_back(adouble(aA.as<double>()[k])); ps.push_back(adouble(sA.as<double>()[k])); }
    ParameterVector pv{pa, ps};

    omp_set_num_threads(threads);
    std::unique_ptr<InferenceManager> im;
    auto t_c0 = std::chrono::steady_clock::now();
    if (P == 1) {
        FixedVector<int, 1> nn, nna;
        nn << n[0];
        nna << na[0];
        im.reset(new NPopInferenceManager<1>(nn, nna, Ls, obs_ptrs, hs, pol_err, new DummySFS<adouble>(dim, M, sfs_ptrs)));
    } else if (P == 2) {
        FixedVector<int, 2> nn, nna;
        nn << n[0], n[1];
        nna << na[0], na[1];
        im.reset(new NPopInferenceManager<2>(nn, nna, Ls, obs_ptrs, hs, pol_err, new DummySFS<adouble>(dim, M, sfs_ptrs)));
    } else
        throw std::runtime_error("npop must be 1 or 2");
    auto t_c1 = std::chrono::steady_clock::now();
    im->setParams(pv);
    im->setTheta(theta);
    im->setRho(rho);
    im->setAlpha(alpha);
    im->saveGamma = save_gamma != 0;
    // rebuilds of pi / emission table / T are not part of the timed E-step window
    im->do_dirty_work();
    // Optional overrides of what do_dirty_work() left behind (tests of irregular spectra: the reference's own model never
    // produces a transition matrix whose diag(e) Td^T has complex or negative eigenvalues, HMM::Estep handles any input):
    //   override_T [M][M] row-major, override_E [K][M] in the std::map order of emission_probs, override_pi [M]
    if (in.has("override_T")) {
        const double *t = in.get("override_T").as<double>();
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < M; ++j) im->transition(i, j) = adouble(t[(size_t)i * M + j]);
        im->tb.update(im->transition, false);
    }
    if (in.has("override_E")) {
        const double *e = in.get("override_E").as<double>();
        size_t k = 0;
        for (auto &pr : im->emission_probs) {
            for (int mm = 0; mm < M; ++mm) pr.second(mm) = adouble(e[k * M + mm]);
            ++k;
        }
    }
    if (in.has("override_pi")) {
        const double *v = in.get("override_pi").as<double>();
        for (int mm = 0; mm < M; ++mm) im->pi(mm) = adouble(v[mm]);
    }

    std::vector<double> secs;
    for (int r = 0; r < repeat; ++r) {
        auto t0 = std::chrono::steady_clock::now();
        im->Estep(false);
        auto t1 = std::chrono::steady_clock::now();
        secs.push_back(std::chrono::duration<double>(t1 - t0).count());
    }
    out.put_f8("estep_seconds", {(long)secs.size()}, secs.data());
    double ctor_s = std::chrono::duration<double>(t_c1 - t_c0).count();
    out.put_f8("ctor_seconds", {1}, &ctor_s);
    int32_t thr = threads;
    out.put_i4("threads", {1}, &thr);

    // ---- inputs of the forward-backward, as the reference built them
    std::vector<double> buf;
    buf.resize(M);
    for (int m = 0; m < M; ++m) buf[m] = im->pi(m).value();
    out.put_f8("pi", {M}, buf.data());
    buf.resize((size_t)M * M);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) buf[(size_t)i * M + j] = im->tb.Td(i, j);
    out.put_f8("T", {M, M}, buf.data());

    std::vector<block_key> keys;
    for (const auto &p : im->emission_probs) keys.push_back(p.first);
    const int K = (int)keys.size();
    std::map<block_key, int> key_index;
    std::vector<int32_t> kv;
    for (int k = 0; k < K; ++k) {
        key_index[keys[k]] = k;
        for (int q = 0; q < 3 * P; ++q) kv.push_back(keys[k](q));
    }
    out.put_i4("keys", {K, 3 * P}, kv.data());
    buf.resize((size_t)K * M);
    for (int k = 0; k < K; ++k) {
        const Vector<adouble> &e = im->emission_probs.at(keys[k]);
        for (int m = 0; m < M; ++m) buf[(size_t)k * M + m] = e(m).value();
    }
    out.put_f8("E", {K, M}, buf.data());

    // eigensystems (reference include/transition_bundle.h:9-30)
    {
        std::vector<int32_t> eidx, cplx;
        std::vector<double> Pr, Pir, d, ds, sc;
        for (const auto &p : im->tb.eigensystems) {
            eidx.push_back(key_index.at(p.first));
            const eigensystem &es = p.second;
            for (int i = 0; i < M; ++i)
                for (int j = 0; j < M; ++j) { Pr.push_back(es.P_r(i, j)); }
            for (int i = 0; i < M; ++i)
                for (int j = 0; j < M; ++j) { Pir.push_back(es.Pinv_r(i, j)); }
            for (int i = 0; i < M; ++i) { d.push_back(es.d_r(i)); ds.push_back(es.d_r_scaled(i)); }
            sc.push_back(es.scale);
            cplx.push_back(es.cplx ? 1 : 0);
        }
        long ne = (long)eidx.size();
        out.put_i4("eig_key_idx", {ne}, eidx.data());
        out.put_f8("eig_P", {ne, M, M}, Pr.data());
        out.put_f8("eig_Pinv", {ne, M, M}, Pir.data());
        out.put_f8("eig_d", {ne, M}, d.data());
        out.put_f8("eig_dscaled", {ne, M}, ds.data());
        out.put_f8("eig_scale", {ne}, sc.data());
        out.put_i4("eig_cplx", {ne}, cplx.data());
        int64_t nt = (int64_t)im->targets.size();
        out.put<int64_t>("n_span_key_targets", "i8", {1}, &nt);
    }

    // ---- rate-function grid, emission tensor and bins (inputs of rows a3-a5, a14)
    {
        const PiecewiseConstantRateFunction<adouble> &eta = *im->eta;
        const std::vector<double> &ts = eta.getTs();
        out.put_f8("eta_ts", {(long)ts.size()}, ts.data());
        std::vector<double> ada, rr, act;
        for (const adouble &x : eta.getAda()) ada.push_back(x.value());
        for (const adouble &x : eta.getRrng()) rr.push_back(x.value());
        for (const adouble &x : eta.average_coal_times()) act.push_back(x.value());
        out.put_f8("eta_ada", {(long)ada.size()}, ada.data());
        out.put_f8("eta_Rrng", {(long)rr.size()}, rr.data());
        out.put_f8("eta_avg_coal_times", {(long)act.size()}, act.data());
        std::vector<int32_t> hi(eta.getHsIndices().begin(), eta.getHsIndices().end());
        out.put_i4("eta_hs_indices", {(long)hi.size()}, hi.data());
        const Matrix<adouble> &em = im->emission;
        buf.resize((size_t)em.rows() * em.cols());
        for (int i = 0; i < em.rows(); ++i)
            for (int j = 0; j < em.cols(); ++j) buf[(size_t)i * em.cols() + j] = em(i, j).value();
        out.put_f8("emission_tensor", {(long)em.rows(), (long)em.cols()}, buf.data());
        if (P == 1) dump_bins<1>(out, *dynamic_cast<NPopInferenceManager<1> *>(im.get()), keys);
        else dump_bins<2>(out, *dynamic_cast<NPopInferenceManager<2> *>(im.get()), keys);
    }

    // ---- outputs of HMM::Estep per contig (reference include/hmm.h:33-37)
    std::vector<double> ll(C), xis((size_t)C * M * M), g0((size_t)C * M), gs((size_t)C * K * M, 0.0);
    std::vector<uint8_t> present((size_t)C * K, 0);
    for (int c = 0; c < C; ++c) {
        HMM &h = *im->hmms[c];
        ll[c] = h.ll;
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < M; ++j) xis[((size_t)c * M + i) * M + j] = h.xisum(i, j);
        for (int m = 0; m < M; ++m) g0[(size_t)c * M + m] = h.gamma(m, 0);
        for (const auto &p : h.gamma_sums) {
            int k = key_index.at(p.first);
            present[(size_t)c * K + k] = 1;
            for (int m = 0; m < M; ++m) gs[((size_t)c * K + k) * M + m] = p.second(m);
        }
        if (save_gamma) {
            const long L = Ls[c];
            std::vector<double> g((size_t)(L + 1) * M);
            for (long l = 0; l <= L; ++l)
                for (int m = 0; m < M; ++m) g[(size_t)l * M + m] = h.gamma(m, l);
            out.put_f8("gamma_full_" + std::to_string(c), {L + 1, M}, g.data());
        }
        if (dump_alpha) {
            const long L = Ls[c];
            std::vector<float> a((size_t)(L + 1) * M);
            for (long l = 0; l <= L; ++l)
                for (int m = 0; m < M; ++m) a[(size_t)l * M + m] = h.alpha_hat(m, l);
            out.put_f4("alpha_hat_" + std::to_string(c), {L + 1, M}, a.data());
            std::vector<double> lc(L + 1);
            for (long l = 0; l <= L; ++l) lc[l] = h.log_c(l);
            out.put_f8("log_c_" + std::to_string(c), {L + 1}, lc.data());
        }
    }
    out.put_f8("ll", {C}, ll.data());
    out.put_f8("xisum", {C, M, M}, xis.data());
    out.put_f8("gamma0", {C, M}, g0.data());
    out.put_f8("gamma_sums", {C, K, M}, gs.data());
    out.put_u1("key_present", {C, K}, present.data());

    // M-step objective pieces (reference src/hmm.cpp:155-193, src/inference_manager.cpp:116-126)
    std::vector<adouble> q = im->Q();
    double qv[4] = {q[0].value(), q[1].value(), q[2].value(), q[3].value()};
    out.put_f8("Q", {4}, qv);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <in.smcb> <out.smcb> [threads] [repeat]\n", argv[0]);
        return 2;
    }
    init_logger_cb(quiet_logger);
    int threads = argc > 3 ? std::atoi(argv[3]) : 1;
    int repeat = argc > 4 ? std::atoi(argv[4]) : 1;
    if (threads < 1) threads = 1;
    if (repeat < 1) repeat = 1;
    try {
        smcb::Bundle in = smcb::Bundle::load(argv[1]);
        smcb::Bundle out;
        int rc = run(in, out, threads, repeat);
        out.save(argv[2]);
        return rc;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ref_harness: %s\n", e.what());
        return 1;
    }
}

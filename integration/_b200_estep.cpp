// integration/_b200_estep.cpp -- the reference-side binding of libsmcpp_b200.so, as a real translation unit.
//
// Drop this file into the reference as src/_b200_estep.cpp (setup.py: add it to `sources`, link smcpp_b200).  It
// REPLACES the body of InferenceManager::Estep (reference src/inference_manager.cpp:108-114):
//
//     do_dirty_work(); tb.update(transition, true); parallel_do(hmm->Estep(fbonly));
//
// by one batched call into the CUDA library, and afterwards fills the members that the untouched rest of the reference
// reads -- HMM::ll, HMM::xisum, HMM::gamma, HMM::gamma_sums (include/hmm.h:33-37; InferenceManager is a friend,
// include/hmm.h:11) -- so that HMM::Q (src/hmm.cpp:155-193), InferenceManager::Q / loglik / getXisums / getGammas /
// getGammaSums (src/inference_manager.cpp:116-150, 174-177) and therefore smcpp/_smcpp.pyx work unchanged.
//
// Everything here is written against the reference's UNMODIFIED headers:
//   * the library context is created lazily on the first E-step -- bpm_keys / emission_probs are filled by
//     populate_emission_probs() in the body of the DERIVED NPopInferenceManager constructor
//     (include/inference_manager.h:100-103), i.e. after the base constructor has returned, so the base constructor is too
//     early to hand the key table to smcpp_b200_set_contigs;
//   * the context lives in a side table keyed by `this` (a maintainer editing the header would add a member
//     `smcpp_b200_ctx *b200` and release it in the destructor instead; b200_release() below is that release);
//   * InferenceManager::Estep is called under `with nogil` and declared WITHOUT `except +` (smcpp/_smcpp.pxd:49,
//     smcpp/_smcpp.pyx:189-190): nothing may throw out of it.  On a library error the message goes through the
//     reference's logger at CRITICAL level, every HMM::ll becomes NaN and b200_failed(this) turns true (a maintainer
//     rethrows it from Q(), which IS `except +`).
//
// In this repository the file is compiled against the reference's own objects by oracle/ref_build/Makefile (target
// ref_harness_b200: the reference's inference_manager.o with its own Estep symbol localised + this file +
// libsmcpp_b200.so) and tested on the GPU against the unmodified ref_harness (tests/test_dropin.py).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "inference_manager.h"
#include "smcpp_b200.h"

namespace {

struct B200State {
    smcpp_b200_ctx *ctx = nullptr;
    std::vector<block_key> keys;      // global key table in the reference's std::map order (= bpm_keys)
    bool failed = false;
    std::string error;
};

std::mutex g_mu;
std::map<const InferenceManager *, B200State> g_state;

B200State &state_of(const InferenceManager *im)
{
    std::lock_guard<std::mutex> lock(g_mu);
    return g_state[im];
}

}  // namespace

// release hook (a maintainer calls it from ~InferenceManager)
void b200_release(const InferenceManager *im)
{
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_state.find(im);
    if (it == g_state.end()) return;
    if (it->second.ctx) smcpp_b200_destroy(it->second.ctx);
    g_state.erase(it);
}

bool b200_failed(const InferenceManager *im, std::string *why)
{
    B200State &st = state_of(im);
    if (why) *why = st.error;
    return st.failed;
}

void InferenceManager::Estep(bool /* fbonly: never read by the reference either, src/hmm.cpp:45 */)
{
    DEBUG1 << "E step (smcpp_b200)";
    B200State &st = state_of(this);
    const int C = (int)hmms.size();
    auto give_up = [&](const std::string &msg) {
        st.failed = true;
        st.error = msg;
        CRITICAL << "smcpp_b200 E-step failed: " << msg;
        for (int c = 0; c < C; ++c) hmms[c]->ll = std::numeric_limits<double>::quiet_NaN();
    };
    try {
        do_dirty_work();                     // pi, emission_probs, transition (+ tb.T / tb.Td for HMM::Q): unchanged reference code
        if (tb.Td.rows() != M) tb.update(transition, false);
        // ---- one-time: context + observations (replaces the span tables / eigensystem targets of the constructor)
        if (!st.ctx) {
            const char *dev = std::getenv("SMCPP_B200_DEVICE");
            if (smcpp_b200_create(&st.ctx, dev ? std::atoi(dev) : 0)) { give_up(smcpp_b200_last_error(nullptr)); return; }
            st.keys.clear();
            std::vector<int32_t> key_table;
            for (const auto &p : emission_probs) {            // std::map order = block_key::operator< = bpm_keys
                st.keys.push_back(p.first);
                for (int q = 0; q < p.first.size(); ++q) key_table.push_back(p.first(q));
            }
            std::vector<const int32_t *> ptrs;
            std::vector<int32_t> lens;
            for (auto &ob : obs) {                            // int32 row-major [L][1 + 3 npop], as mapped by map_obs()
                ptrs.push_back(ob.data());
                lens.push_back((int32_t)ob.rows());
            }
            if (smcpp_b200_set_contigs(st.ctx, C, ptrs.data(), lens.data(), npop, key_table.data(), (int)st.keys.size())) {
                give_up(smcpp_b200_last_error(st.ctx));
                smcpp_b200_destroy(st.ctx);
                st.ctx = nullptr;
                return;
            }
        }
        // ---- per E-step inputs: value parts of what do_dirty_work() left behind
        const int K = (int)st.keys.size();
        std::vector<double> pi_d(M), T_d((size_t)M * M), E_d((size_t)K * M);
        for (int m = 0; m < M; ++m) pi_d[m] = toDouble(pi(m));
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < M; ++j) T_d[(size_t)i * M + j] = tb.Td(i, j);               // the C ABI is row-major
        for (int k = 0; k < K; ++k) {
            const Vector<adouble> &e = emission_probs.at(st.keys[k]);
            for (int m = 0; m < M; ++m) E_d[(size_t)k * M + m] = toDouble(e(m));
        }
        std::vector<double> ll(C), xis((size_t)C * M * M), g0((size_t)C * M), gs((size_t)C * K * M);
        std::vector<uint8_t> present((size_t)C * K);
        smcpp_b200_set_save_gamma(st.ctx, saveGamma ? 1 : 0);
        // P == NULL: the library computes the eigensystems of diag(e_key) Td^T (TransitionBundle::update(T, true) of the
        // reference, src/transition_bundle.cpp:14-25) itself; tb.eigensystems / tb.span_Qs are not needed any more.
        if (smcpp_b200_estep(st.ctx, M, pi_d.data(), T_d.data(), E_d.data(), 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                             ll.data(), xis.data(), g0.data(), gs.data(), nullptr) ||
            smcpp_b200_get_key_present(st.ctx, present.data())) {
            give_up(smcpp_b200_last_error(st.ctx));
            return;
        }
        // ---- results -> the members HMM::Q and the getters read
        typedef Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> RowMat;
        std::vector<double> gbuf;
        for (int c = 0; c < C; ++c) {
            HMM &h = *hmms[c];
            h.ll = ll[c];
            h.xisum = Eigen::Map<const RowMat>(&xis[(size_t)c * M * M], M, M);
            if (saveGamma) {                                   // M x (L+1), the `smc++ posterior` path (src/hmm.cpp:48-49, 147-150)
                gbuf.resize((size_t)(h.L + 1) * M);
                if (smcpp_b200_fetch_gamma(st.ctx, c, gbuf.data())) { give_up(smcpp_b200_last_error(st.ctx)); return; }
                h.gamma = Eigen::Map<const RowMat>(gbuf.data(), h.L + 1, M).transpose();
            } else {
                h.gamma = Eigen::Map<const Vector<double> >(&g0[(size_t)c * M], M);     // M x 1: gamma.col(0), src/hmm.cpp:150
            }
            h.gamma_sums.clear();                              // exactly the keys that occur in this contig (src/hmm.cpp:51-53, 69)
            for (int k = 0; k < K; ++k)
                if (present[(size_t)c * K + k])
                    h.gamma_sums.emplace(st.keys[k], Eigen::Map<const Vector<double> >(&gs[((size_t)c * K + k) * M], M));
        }
        st.failed = false;
        st.error.clear();
    } catch (const std::exception &e) {
        give_up(e.what());
    } catch (...) {
        give_up("unknown exception");
    }
}

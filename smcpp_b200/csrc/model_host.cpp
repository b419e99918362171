// smcpp_b200 -- host-side builders of the per-E-step HMM inputs: the initial
// distribution pi, the transition matrix and the per-key emission vectors, from the piecewise-constant size
// history eta, rho, theta and caller-supplied conditioned-SFS matrices.  They are O(M^2 + K M) per E-step and
// independent of the sequence length, so they stay on the host (the reference runs them under MPFR / autodiff);
// value parts only -- the derivative vectors of the reference's adouble are an M-step concern.
//
// Re-statement, in our own structure, of what these reference routines compute (cited per function); quirks
// that change numbers are kept and marked "verbatim".
#include "model_host.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <set>
#include <vector>

extern "C" {
__float128 expq(__float128);
__float128 sinhq(__float128);
__float128 coshq(__float128);
__float128 sqrtq(__float128);
}

namespace smcb {
namespace {

// ---- eta: breakpoints merged with the hidden-state boundaries (reference src/piecewise_constant_rate_function.cpp:31-84)
struct Rate {
    std::vector<double> ada, ts, Rrng, hs;
    std::vector<int> hs_idx;

    Rate(const std::vector<double> &a, const std::vector<double> &s, const std::vector<double> &hidden) : hs(hidden)
    {
        const int K0 = (int)a.size();
        ada.resize(K0);
        ts.assign(K0 + 1, 0.0);
        for (int k = 0; k < K0; ++k) {
            ada[k] = 1.0 / a[k];
            ts[k + 1] = ts[k] + s[k];
        }
        ts[K0] = INFINITY;   // the last piece is flat to infinity
        for (double h : hs) {
            if (std::isinf(h)) {
                hs_idx.push_back((int)ts.size() - 1);
                continue;
            }
            const int ip = (int)(std::upper_bound(ts.begin(), ts.end(), h) - ts.begin()) - 1;
            if (std::fabs(ts[ip] - h) < 1e-8) hs_idx.push_back(ip);
            else if (ip + 1 < (int)ts.size() && std::fabs(ts[ip + 1] - h) < 1e-8) hs_idx.push_back(ip + 1);
            else {
                ts.insert(ts.begin() + ip + 1, h);
                ada.insert(ada.begin() + ip + 1, ada[ip]);
                hs_idx.push_back(ip + 1);
            }
        }
        const int K = (int)ada.size();
        Rrng.assign(K + 1, 0.0);
        for (int k = 0; k < K; ++k) Rrng[k + 1] = Rrng[k] + ada[k] * (ts[k + 1] - ts[k]);   // :165-170 (0*inf never occurs: last ada > 0 gives inf)
    }
    // cumulative hazard R(t); reference :406-411
    double R(double t) const
    {
        const int ip = (int)(std::upper_bound(ts.begin(), ts.end(), t) - ts.begin()) - 1;
        return Rrng[ip] + ada[ip] * (t - ts[ip]);
    }
    // int_a^b exp(-(R(t) + log_denom)) dt; reference :173-203
    double R_integral(double a, double b, double log_denom) const
    {
        int ip_a = (int)(std::upper_bound(ts.begin(), ts.end(), a) - ts.begin()) - 1;
        int ip_b = (int)(std::upper_bound(ts.begin(), ts.end(), b) - ts.begin()) - 1;
        if (std::isinf(b)) ip_b = (int)ts.size() - 2;
        double ret = 0.0;
        for (int i = ip_a; i < ip_b + 1; ++i) {
            const double left = std::max(a, ts[i]), right = std::min(b, ts[i + 1]), diff = right - left;
            double r = std::exp(-(R(left) + log_denom));
            if (ada[i] > 0.0) {
                if (!std::isinf(diff)) r *= -std::expm1(-diff * ada[i]);
                r /= ada[i];
            } else
                r *= diff;
            ret += r;
        }
        return ret;
    }
    // expected coalescence time within each hidden state; reference :372-403 (NaN marks "no coalescence possible")
    std::vector<double> average_coal_times() const
    {
        std::vector<double> ret;
        for (size_t i = 1; i < hs.size(); ++i) {
            if (Rrng[hs_idx[i - 1]] == Rrng[hs_idx[i]]) {
                ret.push_back(std::numeric_limits<double>::quiet_NaN());
                continue;
            }
            double log_denom = -Rrng[hs_idx[i - 1]];
            const bool inf = std::isinf(ts[hs_idx[i]]);
            if (!inf) log_denom += std::log(-std::expm1(-(Rrng[hs_idx[i]] - Rrng[hs_idx[i - 1]])));
            double x = hs[i - 1] * std::exp(-(Rrng[hs_idx[i - 1]] + log_denom)) + R_integral(ts[hs_idx[i - 1]], ts[hs_idx[i]], log_denom);
            if (!inf) x -= hs[i] * std::exp(-(Rrng[hs_idx[i]] + log_denom));
            ret.push_back(x);
        }
        return ret;
    }
};

// ---- 3x3 generator exponential of the {no recombination, floating, coalesced} chain; reference src/transition.cpp:112-130
template <typename T> struct M3 { T v[3][3]; };
template <typename T> M3<T> eye3()
{
    M3<T> r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.v[i][j] = i == j ? T(1) : T(0);
    return r;
}
template <typename T> M3<T> mul3(const M3<T> &a, const M3<T> &b)
{
    M3<T> r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            T s = T(0);
            for (int k = 0; k < 3; ++k) s += a.v[i][k] * b.v[k][j];
            r.v[i][j] = s;
        }
    return r;
}
inline double fsqrt(double x) { return std::sqrt(x); }
inline double fsinh(double x) { return std::sinh(x); }
inline double fcosh(double x) { return std::cosh(x); }
inline double fexp(double x) { return std::exp(x); }
inline __float128 fsqrt(__float128 x) { return sqrtq(x); }
inline __float128 fsinh(__float128 x) { return sinhq(x); }
inline __float128 fcosh(__float128 x) { return coshq(x); }
inline __float128 fexp(__float128 x) { return expq(x); }

template <typename T> M3<T> matrix_exp3(T c_rho, T c_eta)
{
    const T sq = fsqrt(T(4) * c_eta * c_eta + c_rho * c_rho);
    const T s = fsinh(T(0.5) * sq) / sq, c = fcosh(T(0.5) * sq), e = fexp(-c_eta - c_rho / T(2));
    M3<T> Q;
    Q.v[0][0] = e * (c + (T(2) * c_eta - c_rho) * s);
    Q.v[0][1] = T(2) * e * c_rho * s;
    Q.v[0][2] = T(1) - Q.v[0][0] - Q.v[0][1];
    Q.v[1][0] = T(2) * e * c_eta * s;
    Q.v[1][1] = e * (c - (T(2) * c_eta - c_rho) * s);
    Q.v[1][2] = T(1) - Q.v[1][0] - Q.v[1][1];
    Q.v[2][0] = T(0); Q.v[2][1] = T(0); Q.v[2][2] = T(1);
    return Q;
}

std::vector<double> to_vec(const double *p, int n) { return std::vector<double>(p, p + n); }

// closed-form hypergeometric pmf (the reference calls gsl_ran_hypergeometric_pdf, include/marginalize_key.h:47)
double lchoose(double n, double r) { return std::lgamma(n + 1.) - std::lgamma(r + 1.) - std::lgamma(n - r + 1.); }
double hypergeom_pdf(int k, int n1, int n2, int t)
{
    if (t > n1 + n2) t = n1 + n2;
    if (k > n1 || k > t) return 0.;
    if (t > n2 && k + n2 < t) return 0.;
    return std::exp(lchoose(n1, k) + lchoose(n2, t - k) - lchoose(n1 + n2, t));
}

typedef std::vector<int> Key;   // 3 ints per population: a, b, nb

}  // namespace

int host_initial_distribution(int M, const double *hidden_states, int n_pieces, const double *a, const double *s, double *pi)
{
    // reference src/inference_manager.cpp:56-69
    Rate eta(to_vec(a, n_pieces), to_vec(s, n_pieces), to_vec(hidden_states, M + 1));
    double sum = 0.0;
    for (int m = 0; m < M - 1; ++m) pi[m] = std::exp(-eta.R(hidden_states[m])) - std::exp(-eta.R(hidden_states[m + 1]));
    pi[M - 1] = std::exp(-eta.R(hidden_states[M - 1]));
    for (int m = 0; m < M; ++m) {
        if (pi[m] < 1e-20) pi[m] = 1e-20;
        sum += pi[m];
    }
    for (int m = 0; m < M; ++m) pi[m] /= sum;
    return 0;
}

int host_average_coal_times(int M, const double *hidden_states, int n_pieces, const double *a, const double *s, double *out)
{
    Rate eta(to_vec(a, n_pieces), to_vec(s, n_pieces), to_vec(hidden_states, M + 1));
    const std::vector<double> v = eta.average_coal_times();
    for (int m = 0; m < M; ++m) out[m] = v[m];
    return 0;
}

int host_transition(int M, const double *hidden_states, int n_pieces, const double *a, const double *s, double rho, double *T)
{
    // reference src/transition.cpp:133-253 (HJTransition); Mh = number of hidden-state BOUNDARIES (verbatim: the uniform
    // mixing below is sized with it, so rows sum to 1 - 1e-5/(M+1))
    Rate eta(to_vec(a, n_pieces), to_vec(s, n_pieces), to_vec(hidden_states, M + 1));
    const std::vector<double> &ts = eta.ts, &ada = eta.ada;
    const std::vector<int> &hi = eta.hs_idx;
    const int Mh = M + 1, nts = (int)ts.size();
    const std::vector<double> avg = eta.average_coal_times();
    // running products of the interval exponentials in 113-bit arithmetic (the reference uses 256-bit MPFR; both are
    // indistinguishable after the cast to double, SURVEY probe P5), then cast to double
    typedef __float128 Q;
    std::vector<M3<double>> expms(nts), prods(nts);
    {
        std::vector<M3<Q>> eU(nts, eye3<Q>()), pU(nts, eye3<Q>());
        for (int i = hi[0] + 1; i < nts; ++i) {
            if (!std::isinf(ts[i])) {
                const double delta = ts[i] - ts[i - 1];
                const Q c_eta = (Q)(ada[i - 1] * delta);
                Q c_rho = (Q)delta;
                c_rho *= (Q)rho;
                eU[i] = matrix_exp3<Q>(c_rho, c_eta);
            }   // verbatim: for the infinite last interval the reference appends its matrix instead of assigning it, so slot i stays I
            pU[i] = mul3(pU[i - 1], eU[i]);
        }
        for (int i = 0; i < nts; ++i)
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    expms[i].v[r][c] = (double)eU[i].v[r][c];
                    prods[i].v[r][c] = (double)pU[i].v[r][c];
                }
    }
    std::vector<int> avc_ip(M);
    for (int j = 0; j < M; ++j) avc_ip[j] = (int)(std::upper_bound(ts.begin(), ts.end(), avg[j]) - ts.begin()) - 1;
    std::vector<double> expm_diff(Mh - 2);
    for (int k = 1; k < Mh - 1; ++k) expm_diff[k - 1] = prods[hi[k]].v[0][2] - prods[hi[k - 1]].v[0][2];
    std::vector<double> Phi((size_t)M * M, 0.0);
    for (int j = 1; j < Mh; ++j) {
        double *row = &Phi[(size_t)(j - 1) * M];
        for (int k = 0; k < j - 1; ++k) row[k] = expm_diff[k];
        const double rct = avg[j - 1];
        const int rct_ip = avc_ip[j - 1];
        M3<double> A = eye3<double>();
        for (int ell = hi[j - 1]; ell < rct_ip; ++ell) A = mul3(A, expms[ell]);   // verbatim index range
        const double delta = rct - ts[rct_ip];
        const double c_eta = ada[rct_ip] * delta;
        const double c_rho = delta * rho;
        A = mul3(A, matrix_exp3<double>(c_rho, c_eta));
        const M3<double> B = mul3(prods[hi[j - 1]], A);
        double Rj = c_eta;
        Rj += ada[rct_ip] * (ts[rct_ip + 1] - rct);
        for (int jj = rct_ip + 2; jj < hi[j]; ++jj) Rj += ada[jj] * (ts[jj + 1] - ts[jj]);   // verbatim: starts at rct_ip + 2
        const double p_float = B.v[0][1] * std::exp(-Rj);
        double Rjk1 = 0.0;
        for (int k = j + 1; k < Mh; ++k) {
            double inc = 0.0;
            for (int jj = hi[k - 1]; jj < hi[k]; ++jj) inc += ada[jj] * (ts[jj + 1] - ts[jj]);
            double p_coal = std::exp(-Rjk1);
            Rjk1 += inc;
            if (!std::isinf(inc)) p_coal *= -std::expm1(-inc);
            row[k - 1] += p_float * p_coal;
        }
        row[j - 1] = 0.0;
        double sum = 0.0;
        for (int k = 0; k < M; ++k) sum += row[k];
        row[j - 1] = 1.0 - sum;
    }
    const double beta = 1e-5, p2 = beta / Mh;
    for (size_t x = 0; x < Phi.size(); ++x) {
        double v = Phi[x] < 1e-20 ? 1e-20 : Phi[x];
        v *= (1 - beta);
        T[x] = v + p2;
    }
    return 0;
}

int host_emission(int npop, const int *n, const int *na, int M, const double *hidden_states, int n_pieces, const double *a,
                  const double *s, double theta, double alpha, double pol_err, const double *sfs, int K, const int32_t *keys,
                  double *E, std::string *msg)
{
    const int P = npop;
    if (theta <= 0) { if (msg) *msg = "mutation rate theta <= 0"; return 1; }   // reference src/conditioned_sfs.cpp:102-103
    int sfs_dim = 1;
    for (int p = 1; p < P; ++p) sfs_dim *= na[p] + 1;
    for (int p = 0; p < P; ++p) sfs_dim *= n[p] + 1;
    const int rows = na[0] + 1, cells = rows * sfs_dim;
    // incorporate_theta + flattening; reference src/conditioned_sfs.cpp:100-148, src/inference_manager.cpp:397-407
    std::vector<double> em((size_t)M * cells);
    for (int m = 0; m < M; ++m) {
        const double *c = sfs + (size_t)m * cells;
        double tauh = 0.0;
        for (int x = 0; x < cells; ++x) tauh += c[x];
        const double f = -std::expm1(-theta * tauh) / tauh;
        double *o = &em[(size_t)m * cells];
        double t2 = 0.0;
        for (int x = 0; x < cells; ++x) { o[x] = c[x] * f; t2 += o[x]; }
        o[0] = 1.0 - t2;
        for (int x = 0; x < cells; ++x)
            if (o[x] < 1e-10) o[x] = 1e-10;
    }
    // two-column table for keys without undistinguished lineages; reference :410-431
    Rate eta(to_vec(a, n_pieces), to_vec(s, n_pieces), to_vec(hidden_states, M + 1));
    const std::vector<double> avg = eta.average_coal_times();
    std::vector<double> e2((size_t)M * 2);
    for (int m = 0; m < M; ++m) {
        if (std::isnan(avg[m])) e2[2 * m] = e2[2 * m + 1] = 1e-20;
        else {
            const double l = -2. * alpha * theta * avg[m];
            e2[2 * m] = std::exp(l);
            e2[2 * m + 1] = -std::expm1(l);
        }
    }
    std::vector<int> dims(2 * P);
    for (int p = 0; p < P; ++p) { dims[2 * p] = na[p] + 1; dims[2 * p + 1] = n[p] + 1; }
    auto is_mono = [&](const Key &k) {   // reference :287-297
        for (int p = 0; p < P; ++p)
            if (k[3 * p] != na[p] || k[3 * p + 1] != k[3 * p + 2]) return false;
        return true;
    };
    for (int kk = 0; kk < K; ++kk) {
        const int32_t *key = keys + (size_t)kk * 3 * P;
        double *out = E + (size_t)kk * M;
        bool reduced = true, miss = true;
        int amin = 1 << 30, asum = 0;
        for (int p = 0; p < P; ++p) {
            reduced = reduced && key[3 * p + 2] == 0;
            if (na[p] > 0) miss = miss && key[3 * p] == -1;
            amin = std::min(amin, (int)key[3 * p]);
            asum += key[3 * p];
        }
        if (reduced && (miss || amin >= 0)) {   // reference :454-460
            for (int m = 0; m < M; ++m) out[m] = miss ? 1.0 : e2[2 * m + (asum % 2)];
        } else {
            // bins of this key: construct_bins, reference :329-386 (+ include/bin_key.h, include/marginalize_key.h)
            std::vector<Key> binned(1, Key());
            for (int p = 0; p < P; ++p) {   // bin_key with cutoff 1.0: a == -1 expands over 0..na, nothing else
                std::vector<Key> next;
                for (const Key &pre : binned) {
                    const int a0 = key[3 * p], lo = a0 == -1 ? 0 : a0, hi = a0 == -1 ? na[p] : a0;
                    for (int aa = lo; aa <= hi; ++aa) {
                        Key k2 = pre;
                        k2.push_back(aa); k2.push_back(key[3 * p + 1]); k2.push_back(key[3 * p + 2]);
                        next.push_back(k2);
                    }
                }
                binned.swap(next);
            }
            std::set<Key> bset(binned.begin(), binned.end());
            std::map<Key, double> mm;
            for (const Key &bk : bset) {
                // marginalize_key: hypergeometric up-sampling of each population to sample size n
                std::map<Key, double> probs;
                probs[Key()] = 1.0;
                for (int p = 0; p < P; ++p) {
                    std::map<Key, double> next;
                    const int ap = bk[3 * p], bp = bk[3 * p + 1], nbp = bk[3 * p + 2];
                    for (const auto &pre : probs)
                        for (int n1 = bp; n1 <= n[p] + bp - nbp; ++n1) {
                            Key k2 = pre.first;
                            k2.push_back(ap); k2.push_back(n1); k2.push_back(n[p]);
                            next[k2] += pre.second * hypergeom_pdf(bp, n1, n[p] - n1, nbp);
                        }
                    probs.swap(next);
                }
                for (const auto &pr : probs) {
                    Key mbk = pr.first;
                    if (is_mono(mbk))
                        for (int p = 0; p < P; ++p) { mbk[3 * p] = 0; mbk[3 * p + 1] = 0; }   // convert_monomorphic :299-312
                    Key fold = mbk;                                                                 // folded_key :315-327
                    for (int p = 0; p < P; ++p) { fold[3 * p] = na[p] - mbk[3 * p]; fold[3 * p + 1] = mbk[3 * p + 2] - mbk[3 * p + 1]; }
                    mm[mbk] += (1. - pol_err) * pr.second;
                    mm[fold] += pol_err * pr.second;
                }
            }
            double ssum = 0.0;
            std::map<Key, double> cellw;   // (a_p, b_p) pairs -> weight
            for (const auto &pr : mm) {
                if (pr.second <= 0 || is_mono(pr.first)) continue;
                ssum += pr.second;
            }
            if (ssum <= 0) { if (msg) *msg = "s<=0"; return 1; }   // reference :375-379
            for (const auto &pr : mm) {
                if (pr.second <= 0 || is_mono(pr.first)) continue;
                Key cell;
                for (int p = 0; p < P; ++p) { cell.push_back(pr.first[3 * p]); cell.push_back(pr.first[3 * p + 1]); }
                cellw[cell] += pr.second / ssum;
            }
            for (int m = 0; m < M; ++m) out[m] = 0.0;
            for (const auto &cw : cellw) {   // tensorRef, reference include/tensorslice.h
                long idx = 0;
                for (int d = 0; d < 2 * P; ++d) idx = idx * dims[d] + cw.first[d];
                for (int m = 0; m < M; ++m) out[m] += cw.second * em[(size_t)m * cells + idx];
            }
        }
        double mx = out[0], mn = out[0];
        for (int m = 1; m < M; ++m) { mx = std::max(mx, out[m]); mn = std::min(mn, out[m]); }
        if (mx > 1.0 || mn <= 0.0) { if (msg) *msg = "probability vector not in [0, 1]"; return 1; }   // reference :466-474
    }
    return 0;
}

}  // namespace smcb

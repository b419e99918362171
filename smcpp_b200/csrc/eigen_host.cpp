// smcpp_b200 -- real non-symmetric eigensolver on the host (Householder reduction to Hessenberg form,
// Francis double-shift QR with accumulation, back-substitution): the classical EISPACK orthes/ortran/hqr2
// sequence, which is also what the reference's Eigen::EigenSolver descends from.  Results are compared
// through P diag(d) Pinv (tests/test_eigen.py), never eigenvector by eigenvector.
#include "eigen_host.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace smcb {
namespace {

// a few persistent helper threads for the per-key decompositions (process-wide, created on first use)
class Pool {
public:
    static Pool &instance()
    {
        static Pool p;
        return p;
    }
    int size() const { return (int)threads_.size(); }
    void submit(std::function<void()> f)
    {
        {
            std::lock_guard<std::mutex> lock(mu_);
            q_.push_back(std::move(f));
        }
        cv_.notify_one();
    }

private:
    Pool()
    {
        const int hw = (int)std::thread::hardware_concurrency();
        const int n = std::max(0, std::min(hw - 1, 7));
        for (int i = 0; i < n; ++i)
            threads_.emplace_back([this]() {
                for (;;) {
                    std::function<void()> f;
                    {
                        std::unique_lock<std::mutex> lock(mu_);
                        cv_.wait(lock, [this]() { return stop_ || !q_.empty(); });
                        if (stop_ && q_.empty()) return;
                        f = std::move(q_.front());
                        q_.pop_front();
                    }
                    f();
                }
            });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> lock(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    bool stop_ = false;
};

struct Mat {
    int n;
    std::vector<double> a;
    explicit Mat(int n_) : n(n_), a((size_t)n_ * n_, 0.0) {}
    double &operator()(int i, int j) { return a[(size_t)i * n + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * n + j]; }
};

// The accumulated transformation: element (i, j) at a[j * n + i], so that the loops over i below (one column of V at a time,
// or two / three adjacent columns of the QR sweep) run over contiguous memory and vectorise.  Element-wise the arithmetic and
// its order are those of the row-major version: results are bit-identical.
struct MatT {
    int n;
    std::vector<double> a;
    explicit MatT(int n_) : n(n_), a((size_t)n_ * n_, 0.0) {}
    double &operator()(int i, int j) { return a[(size_t)j * n + i]; }
    double operator()(int i, int j) const { return a[(size_t)j * n + i]; }
    double *col(int j) { return a.data() + (size_t)j * n; }
};

void hessenberg(Mat &H, MatT &V)
{
    const int n = H.n, low = 0, high = n - 1;
    std::vector<double> ort(n, 0.0), fbuf(n, 0.0);
    for (int m = low + 1; m <= high - 1; ++m) {
        double scale = 0.0;
        for (int i = m; i <= high; ++i) scale += std::fabs(H(i, m - 1));
        if (scale == 0.0) continue;
        double h = 0.0;
        for (int i = high; i >= m; --i) {
            ort[i] = H(i, m - 1) / scale;
            h += ort[i] * ort[i];
        }
        double g = std::sqrt(h);
        if (ort[m] > 0) g = -g;
        h -= ort[m] * g;
        ort[m] -= g;
        {
            // f_j = sum_i ort_i H(i, j) for all j at once (i descending as before, j contiguous)
            double *__restrict fj = fbuf.data();
            for (int j = m; j < n; ++j) fj[j] = 0.0;
            for (int i = high; i >= m; --i) {
                const double oi = ort[i];
                const double *__restrict hr = &H(i, 0);
                for (int j = m; j < n; ++j) fj[j] += oi * hr[j];
            }
            for (int j = m; j < n; ++j) fj[j] /= h;
            for (int i = m; i <= high; ++i) {
                const double oi = ort[i];
                double *__restrict hr = &H(i, 0);
                for (int j = m; j < n; ++j) hr[j] -= fj[j] * oi;
            }
        }
        {
            // four rows at a time: four independent summation chains instead of one (each in its original order)
            int i = 0;
            for (; i + 3 <= high; i += 4) {
                double *__restrict h0 = &H(i, 0), *__restrict h1 = &H(i + 1, 0), *__restrict h2 = &H(i + 2, 0), *__restrict h3 = &H(i + 3, 0);
                double f0 = 0.0, f1 = 0.0, f2 = 0.0, f3 = 0.0;
                for (int j = high; j >= m; --j) {
                    const double o = ort[j];
                    f0 += o * h0[j]; f1 += o * h1[j]; f2 += o * h2[j]; f3 += o * h3[j];
                }
                f0 /= h; f1 /= h; f2 /= h; f3 /= h;
                for (int j = m; j <= high; ++j) {
                    const double o = ort[j];
                    h0[j] -= f0 * o; h1[j] -= f1 * o; h2[j] -= f2 * o; h3[j] -= f3 * o;
                }
            }
            for (; i <= high; ++i) {
                double f = 0.0;
                for (int j = high; j >= m; --j) f += ort[j] * H(i, j);
                f /= h;
                for (int j = m; j <= high; ++j) H(i, j) -= f * ort[j];
            }
        }
        ort[m] *= scale;
        H(m, m - 1) = scale * g;
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V(i, j) = i == j ? 1.0 : 0.0;
    for (int m = high - 1; m >= low + 1; --m) {
        if (H(m, m - 1) == 0.0) continue;
        for (int i = m + 1; i <= high; ++i) ort[i] = H(i, m - 1);
        {
            const double om = ort[m], hm = H(m, m - 1);
            int j = m;
            for (; j + 3 <= high; j += 4) {                       // four columns at a time (independent summation chains)
                double *__restrict v0 = V.col(j), *__restrict v1 = V.col(j + 1), *__restrict v2 = V.col(j + 2), *__restrict v3 = V.col(j + 3);
                double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
                for (int i = m; i <= high; ++i) {
                    const double o = ort[i];
                    g0 += o * v0[i]; g1 += o * v1[i]; g2 += o * v2[i]; g3 += o * v3[i];
                }
                g0 = (g0 / om) / hm; g1 = (g1 / om) / hm; g2 = (g2 / om) / hm; g3 = (g3 / om) / hm;
                for (int i = m; i <= high; ++i) {
                    const double o = ort[i];
                    v0[i] += g0 * o; v1[i] += g1 * o; v2[i] += g2 * o; v3[i] += g3 * o;
                }
            }
            for (; j <= high; ++j) {
                double *__restrict vc = V.col(j);
                double g = 0.0;
                for (int i = m; i <= high; ++i) g += ort[i] * vc[i];
                g = (g / om) / hm;
                for (int i = m; i <= high; ++i) vc[i] += g * ort[i];
            }
        }
    }
}

inline void cdiv(double xr, double xi, double yr, double yi, double &zr, double &zi)
{
    double r, d;
    if (std::fabs(yr) > std::fabs(yi)) {
        r = yi / yr;
        d = yr + r * yi;
        zr = (xr + r * xi) / d;
        zi = (xi - r * xr) / d;
    } else {
        r = yr / yi;
        d = yi + r * yr;
        zr = (r * xr + xi) / d;
        zi = (r * xi - xr) / d;
    }
}

// H upper Hessenberg (destroyed: becomes quasi-triangular T, then holds the triangular eigenvectors),
// V accumulates; on return the columns of V are the (real / real+imag pair) eigenvectors.
int schur_vectors(Mat &H, MatT &V, std::vector<double> &d, std::vector<double> &e)
{
    const int nn = H.n, low = 0, high = nn - 1;
    int n = nn - 1;
    const double eps = std::ldexp(1.0, -52);
    double exshift = 0.0, p = 0, q = 0, r = 0, s = 0, z = 0, t, w, x, y;
    double norm = 0.0;
    for (int i = 0; i < nn; ++i)
        for (int j = std::max(i - 1, 0); j < nn; ++j) norm += std::fabs(H(i, j));
    int iter = 0, total_iter = 0;
    while (n >= low) {
        int l = n;
        while (l > low) {
            s = std::fabs(H(l - 1, l - 1)) + std::fabs(H(l, l));
            if (s == 0.0) s = norm;
            if (std::fabs(H(l, l - 1)) < eps * s) break;
            --l;
        }
        if (l == n) {
            H(n, n) += exshift;
            d[n] = H(n, n);
            e[n] = 0.0;
            --n;
            iter = 0;
        } else if (l == n - 1) {
            w = H(n, n - 1) * H(n - 1, n);
            p = (H(n - 1, n - 1) - H(n, n)) / 2.0;
            q = p * p + w;
            z = std::sqrt(std::fabs(q));
            H(n, n) += exshift;
            H(n - 1, n - 1) += exshift;
            x = H(n, n);
            if (q >= 0) {
                z = p >= 0 ? p + z : p - z;
                d[n - 1] = x + z;
                d[n] = d[n - 1];
                if (z != 0.0) d[n] = x - w / z;
                e[n - 1] = 0.0;
                e[n] = 0.0;
                x = H(n, n - 1);
                s = std::fabs(x) + std::fabs(z);
                p = x / s;
                q = z / s;
                r = std::sqrt(p * p + q * q);
                p /= r;
                q /= r;
                for (int j = n - 1; j < nn; ++j) {
                    z = H(n - 1, j);
                    H(n - 1, j) = q * z + p * H(n, j);
                    H(n, j) = q * H(n, j) - p * z;
                }
                for (int i = 0; i <= n; ++i) {
                    z = H(i, n - 1);
                    H(i, n - 1) = q * z + p * H(i, n);
                    H(i, n) = q * H(i, n) - p * z;
                }
                {
                    double *__restrict va = V.col(n - 1), *__restrict vb = V.col(n);
                    for (int i = low; i <= high; ++i) {
                        const double zz = va[i];
                        va[i] = q * zz + p * vb[i];
                        vb[i] = q * vb[i] - p * zz;
                    }
                }
            } else {
                d[n - 1] = x + p;
                d[n] = x + p;
                e[n - 1] = z;
                e[n] = -z;
            }
            n -= 2;
            iter = 0;
        } else {
            x = H(n, n);
            y = 0.0;
            w = 0.0;
            if (l < n) {
                y = H(n - 1, n - 1);
                w = H(n, n - 1) * H(n - 1, n);
            }
            if (iter == 10) {  // Wilkinson's exceptional shift
                exshift += x;
                for (int i = low; i <= n; ++i) H(i, i) -= x;
                s = std::fabs(H(n, n - 1)) + std::fabs(H(n - 1, n - 2));
                x = y = 0.75 * s;
                w = -0.4375 * s * s;
            }
            if (iter == 30) {  // second exceptional shift
                s = (y - x) / 2.0;
                s = s * s + w;
                if (s > 0) {
                    s = std::sqrt(s);
                    if (y < x) s = -s;
                    s = x - w / ((y - x) / 2.0 + s);
                    for (int i = low; i <= n; ++i) H(i, i) -= s;
                    exshift += s;
                    x = y = w = 0.964;
                }
            }
            ++iter;
            if (++total_iter > 60 * nn + 200) return 1;
            int m = n - 2;
            while (m >= l) {
                z = H(m, m);
                r = x - z;
                s = y - z;
                p = (r * s - w) / H(m + 1, m) + H(m, m + 1);
                q = H(m + 1, m + 1) - z - r - s;
                r = H(m + 2, m + 1);
                s = std::fabs(p) + std::fabs(q) + std::fabs(r);
                p /= s;
                q /= s;
                r /= s;
                if (m == l) break;
                if (std::fabs(H(m, m - 1)) * (std::fabs(q) + std::fabs(r)) <
                    eps * (std::fabs(p) * (std::fabs(H(m - 1, m - 1)) + std::fabs(z) + std::fabs(H(m + 1, m + 1)))))
                    break;
                --m;
            }
            for (int i = m + 2; i <= n; ++i) {
                H(i, i - 2) = 0.0;
                if (i > m + 2) H(i, i - 3) = 0.0;
            }
            for (int k = m; k <= n - 1; ++k) {
                const bool notlast = k != n - 1;
                if (k != m) {
                    p = H(k, k - 1);
                    q = H(k + 1, k - 1);
                    r = notlast ? H(k + 2, k - 1) : 0.0;
                    x = std::fabs(p) + std::fabs(q) + std::fabs(r);
                    if (x == 0.0) continue;
                    p /= x;
                    q /= x;
                    r /= x;
                }
                s = std::sqrt(p * p + q * q + r * r);
                if (p < 0) s = -s;
                if (s != 0) {
                    if (k != m) H(k, k - 1) = -s * x;
                    else if (l != m) H(k, k - 1) = -H(k, k - 1);
                    p += s;
                    x = p / s;
                    y = q / s;
                    z = r / s;
                    q /= p;
                    r /= p;
                    {
                        double *__restrict h0 = &H(k, 0), *__restrict h1 = &H(k + 1, 0);
                        if (notlast) {
                            double *__restrict h2 = &H(k + 2, 0);
                            for (int j = k; j < nn; ++j) {
                                double pp = h0[j] + q * h1[j];
                                pp += r * h2[j];
                                h2[j] -= pp * z;
                                h0[j] -= pp * x;
                                h1[j] -= pp * y;
                            }
                        } else {
                            for (int j = k; j < nn; ++j) {
                                const double pp = h0[j] + q * h1[j];
                                h0[j] -= pp * x;
                                h1[j] -= pp * y;
                            }
                        }
                    }
                    for (int i = 0; i <= std::min(n, k + 3); ++i) {
                        p = x * H(i, k) + y * H(i, k + 1);
                        if (notlast) {
                            p += z * H(i, k + 2);
                            H(i, k + 2) -= p * r;
                        }
                        H(i, k) -= p;
                        H(i, k + 1) -= p * q;
                    }
                    {
                        double *__restrict v0 = V.col(k), *__restrict v1 = V.col(k + 1);
                        if (notlast) {
                            double *__restrict v2 = V.col(k + 2);
                            for (int i = low; i <= high; ++i) {
                                double pp = x * v0[i] + y * v1[i];
                                pp += z * v2[i];
                                v2[i] -= pp * r;
                                v0[i] -= pp;
                                v1[i] -= pp * q;
                            }
                        } else {
                            for (int i = low; i <= high; ++i) {
                                const double pp = x * v0[i] + y * v1[i];
                                v0[i] -= pp;
                                v1[i] -= pp * q;
                            }
                        }
                    }
                }
            }
        }
    }
    if (norm == 0.0) return 0;
    // back-substitution: eigenvectors of the quasi-triangular matrix
    for (n = nn - 1; n >= 0; --n) {
        p = d[n];
        q = e[n];
        if (q == 0) {
            int l = n;
            H(n, n) = 1.0;
            for (int i = n - 1; i >= 0; --i) {
                w = H(i, i) - p;
                r = 0.0;
                for (int j = l; j <= n; ++j) r += H(i, j) * H(j, n);
                if (e[i] < 0.0) {
                    z = w;
                    s = r;
                } else {
                    l = i;
                    if (e[i] == 0.0) {
                        H(i, n) = w != 0.0 ? -r / w : -r / (eps * norm);
                    } else {
                        x = H(i, i + 1);
                        y = H(i + 1, i);
                        q = (d[i] - p) * (d[i] - p) + e[i] * e[i];
                        t = (x * s - z * r) / q;
                        H(i, n) = t;
                        H(i + 1, n) = std::fabs(x) > std::fabs(z) ? (-r - w * t) / x : (-s - y * t) / z;
                    }
                    t = std::fabs(H(i, n));
                    if ((eps * t) * t > 1)
                        for (int j = i; j <= n; ++j) H(j, n) /= t;
                }
            }
        } else if (q < 0) {
            int l = n - 1;
            if (std::fabs(H(n, n - 1)) > std::fabs(H(n - 1, n))) {
                H(n - 1, n - 1) = q / H(n, n - 1);
                H(n - 1, n) = -(H(n, n) - p) / H(n, n - 1);
            } else {
                cdiv(0.0, -H(n - 1, n), H(n - 1, n - 1) - p, q, H(n - 1, n - 1), H(n - 1, n));
            }
            H(n, n - 1) = 0.0;
            H(n, n) = 1.0;
            for (int i = n - 2; i >= 0; --i) {
                double ra = 0.0, sa = 0.0, vr, vi;
                for (int j = l; j <= n; ++j) {
                    ra += H(i, j) * H(j, n - 1);
                    sa += H(i, j) * H(j, n);
                }
                w = H(i, i) - p;
                if (e[i] < 0.0) {
                    z = w;
                    r = ra;
                    s = sa;
                } else {
                    l = i;
                    if (e[i] == 0) {
                        cdiv(-ra, -sa, w, q, H(i, n - 1), H(i, n));
                    } else {
                        x = H(i, i + 1);
                        y = H(i + 1, i);
                        vr = (d[i] - p) * (d[i] - p) + e[i] * e[i] - q * q;
                        vi = (d[i] - p) * 2.0 * q;
                        if (vr == 0.0 && vi == 0.0)
                            vr = eps * norm * (std::fabs(w) + std::fabs(q) + std::fabs(x) + std::fabs(y) + std::fabs(z));
                        cdiv(x * r - z * ra + q * sa, x * s - z * sa - q * ra, vr, vi, H(i, n - 1), H(i, n));
                        if (std::fabs(x) > (std::fabs(z) + std::fabs(q))) {
                            H(i + 1, n - 1) = (-ra - w * H(i, n - 1) + q * H(i, n)) / x;
                            H(i + 1, n) = (-sa - w * H(i, n) - q * H(i, n - 1)) / x;
                        } else {
                            cdiv(-r - y * H(i, n - 1), -s - y * H(i, n), z, q, H(i + 1, n - 1), H(i + 1, n));
                        }
                    }
                    t = std::max(std::fabs(H(i, n - 1)), std::fabs(H(i, n)));
                    if ((eps * t) * t > 1)
                        for (int j = i; j <= n; ++j) {
                            H(j, n - 1) /= t;
                            H(j, n) /= t;
                        }
                }
            }
        }
    }
    // back-transform with the accumulated orthogonal matrix
    {
        std::vector<double> acc(nn);
        for (int j = nn - 1; j >= low; --j) {
            double *__restrict out = acc.data();
            for (int i = low; i <= high; ++i) out[i] = 0.0;
            for (int k = low; k <= std::min(j, high); ++k) {          // k ascending per element, as before
                const double hk = H(k, j);
                const double *__restrict vk = V.col(k);
                for (int i = low; i <= high; ++i) out[i] += vk[i] * hk;
            }
            double *__restrict vj = V.col(j);
            for (int i = low; i <= high; ++i) vj[i] = out[i];
        }
    }
    return 0;
}

// complex inverse by LU with partial pivoting (n <= 128)
int complex_inverse(int n, std::vector<std::complex<double>> &A, std::vector<std::complex<double>> &Inv)
{
    typedef std::complex<double> cd;
    std::vector<int> piv(n);
    for (int i = 0; i < n; ++i) piv[i] = i;
    for (int k = 0; k < n; ++k) {
        int best = k;
        double bv = std::abs(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            double v = std::abs(A[(size_t)i * n + k]);
            if (v > bv) { bv = v; best = i; }
        }
        if (bv == 0.0) return 1;
        if (best != k) {
            for (int j = 0; j < n; ++j) std::swap(A[(size_t)k * n + j], A[(size_t)best * n + j]);
            std::swap(piv[k], piv[best]);
        }
        const cd pivv = A[(size_t)k * n + k];
        for (int i = k + 1; i < n; ++i) {
            const cd f = A[(size_t)i * n + k] / pivv;
            A[(size_t)i * n + k] = f;
            if (f != cd(0.0))
                for (int j = k + 1; j < n; ++j) A[(size_t)i * n + j] -= f * A[(size_t)k * n + j];
        }
    }
    Inv.assign((size_t)n * n, cd(0.0));
    std::vector<cd> col(n);
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) col[i] = piv[i] == c ? cd(1.0) : cd(0.0);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < i; ++j) col[i] -= A[(size_t)i * n + j] * col[j];
        for (int i = n - 1; i >= 0; --i) {
            for (int j = i + 1; j < n; ++j) col[i] -= A[(size_t)i * n + j] * col[j];
            col[i] /= A[(size_t)i * n + i];
        }
        for (int i = 0; i < n; ++i) Inv[(size_t)i * n + c] = col[i];
    }
    return 0;
}

}  // namespace

// the same elimination in real arithmetic (a real spectrum has real eigenvectors: a quarter of the complex flops)
int real_inverse(int n, std::vector<double> &A, double *Inv)
{
    std::vector<int> piv(n);
    for (int i = 0; i < n; ++i) piv[i] = i;
    for (int k = 0; k < n; ++k) {
        int best = k;
        double bv = std::fabs(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            const double v = std::fabs(A[(size_t)i * n + k]);
            if (v > bv) { bv = v; best = i; }
        }
        if (bv == 0.0) return 1;
        if (best != k) {
            for (int j = 0; j < n; ++j) std::swap(A[(size_t)k * n + j], A[(size_t)best * n + j]);
            std::swap(piv[k], piv[best]);
        }
        const double pivv = A[(size_t)k * n + k];
        for (int i = k + 1; i < n; ++i) {
            const double f = A[(size_t)i * n + k] / pivv;
            A[(size_t)i * n + k] = f;
            if (f != 0.0)
                for (int j = k + 1; j < n; ++j) A[(size_t)i * n + j] -= f * A[(size_t)k * n + j];
        }
    }
    // all n right-hand sides at once: X(i, :) -= A(i, j) X(j, :) with the columns contiguous (per element the same
    // operations in the same order as one column at a time)
    double *__restrict X = Inv;
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < n; ++c) X[(size_t)i * n + c] = piv[i] == c ? 1.0 : 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            const double a = A[(size_t)i * n + j];
            const double *__restrict xj = X + (size_t)j * n;
            double *__restrict xi = X + (size_t)i * n;
            for (int c = 0; c < n; ++c) xi[c] -= a * xj[c];
        }
    for (int i = n - 1; i >= 0; --i) {
        double *__restrict xi = X + (size_t)i * n;
        for (int j = i + 1; j < n; ++j) {
            const double a = A[(size_t)i * n + j];
            const double *__restrict xj = X + (size_t)j * n;
            for (int c = 0; c < n; ++c) xi[c] -= a * xj[c];
        }
        const double dd = A[(size_t)i * n + i];
        for (int c = 0; c < n; ++c) xi[c] /= dd;
    }
    return 0;
}

int host_eig_real_general(int n, const double *A, double *P_r, double *Pinv_r, double *d_r, double *d_i, std::string *msg)
{
    typedef std::complex<double> cd;
    Mat H(n);
    MatT V(n);
    std::copy(A, A + (size_t)n * n, H.a.begin());
    std::vector<double> d(n, 0.0), e(n, 0.0);
    if (n == 1) {
        P_r[0] = 1.0; Pinv_r[0] = 1.0; d_r[0] = A[0]; d_i[0] = 0.0;
        return 0;
    }
    hessenberg(H, V);
    if (schur_vectors(H, V, d, e)) {
        if (msg) *msg = "QR iteration did not converge";
        return 1;
    }
    bool real_spectrum = true;
    for (int j = 0; j < n; ++j) real_spectrum = real_spectrum && e[j] == 0.0;
    if (real_spectrum) {
        // real eigenvectors, unit 2-norm columns (as Eigen's EigenSolver::eigenvectors()), real inverse
        std::vector<double> LUr((size_t)n * n);
        for (int j = 0; j < n; ++j) {
            double nrm = 0.0;
            for (int i = 0; i < n; ++i) nrm += V(i, j) * V(i, j);
            nrm = std::sqrt(nrm);
            for (int i = 0; i < n; ++i) P_r[(size_t)i * n + j] = LUr[(size_t)i * n + j] = V(i, j) / nrm;
        }
        if (real_inverse(n, LUr, Pinv_r)) {
            if (msg) *msg = "eigenvector matrix is singular";
            return 1;
        }
        for (int i = 0; i < n; ++i) { d_r[i] = d[i]; d_i[i] = 0.0; }
        return 0;
    }
    // complex eigenvector matrix, unit 2-norm columns (as Eigen's EigenSolver::eigenvectors())
    std::vector<cd> Pc((size_t)n * n), Pinv;
    for (int j = 0; j < n; ++j) {
        if (e[j] == 0.0) {
            double nrm = 0.0;
            for (int i = 0; i < n; ++i) nrm += V(i, j) * V(i, j);
            nrm = std::sqrt(nrm);
            for (int i = 0; i < n; ++i) Pc[(size_t)i * n + j] = cd(V(i, j) / nrm, 0.0);
        } else if (e[j] > 0.0 && j + 1 < n) {
            double nrm = 0.0;
            for (int i = 0; i < n; ++i) nrm += V(i, j) * V(i, j) + V(i, j + 1) * V(i, j + 1);
            nrm = std::sqrt(nrm);
            for (int i = 0; i < n; ++i) {
                Pc[(size_t)i * n + j] = cd(V(i, j) / nrm, V(i, j + 1) / nrm);
                Pc[(size_t)i * n + j + 1] = cd(V(i, j) / nrm, -V(i, j + 1) / nrm);
            }
            ++j;
        }
    }
    std::vector<cd> LU = Pc;
    if (complex_inverse(n, LU, Pinv)) {
        if (msg) *msg = "eigenvector matrix is singular";
        return 1;
    }
    for (size_t x = 0; x < (size_t)n * n; ++x) {
        P_r[x] = Pc[x].real();
        Pinv_r[x] = Pinv[x].real();
    }
    for (int i = 0; i < n; ++i) {
        d_r[i] = d[i];
        d_i[i] = e[i];
    }
    return 0;
}

int host_eigensystems(int M, int K, int n_eig, const int32_t *eig_keys, const double *T, const double *E, double *P,
                      double *Pinv, double *d, double *d_scaled, double *scale, int32_t *cplx, std::string *msg)
{
    for (int e = 0; e < n_eig; ++e)
        if (eig_keys[e] < 0 || eig_keys[e] >= K) {
            if (msg) *msg = "eigen key index out of range";
            return 1;
        }
    // one decomposition per key (reference src/transition_bundle.cpp:14-25 runs them one after the other); the keys are
    // independent, so they go to host threads -- the E-step waits for this before its first kernel
    std::vector<int> rc(std::max(1, n_eig), 0);
    std::vector<std::string> msgs(std::max(1, n_eig));
    auto one = [&](int e) {
        std::vector<double> A((size_t)M * M), di(M);
        const double *ek = E + (size_t)eig_keys[e] * M;
        // diag(e_key) Td^T, reference src/transition_bundle.cpp:19-20
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < M; ++j) A[(size_t)i * M + j] = ek[i] * T[(size_t)j * M + i];
        double *de = d + (size_t)e * M;
        if (host_eig_real_general(M, A.data(), P + (size_t)e * M * M, Pinv + (size_t)e * M * M, de, di.data(), &msgs[e])) {
            rc[e] = 1;
            return;
        }
        double sc = 0.0, im = 0.0;
        for (int i = 0; i < M; ++i) {
            sc = std::max(sc, std::hypot(de[i], di[i]));
            im = std::max(im, std::fabs(di[i]));
        }
        scale[e] = sc;
        for (int i = 0; i < M; ++i) d_scaled[(size_t)e * M + i] = de[i] / sc;
        if (cplx) cplx[e] = im > 0.0;
    };
    // helper threads are kept between calls (starting one costs ~0.1 ms, a 32 x 32 decomposition 0.3 ms)
    const int workers = std::min(n_eig - 1, Pool::instance().size());
    if (workers <= 0) {
        for (int e = 0; e < n_eig; ++e) one(e);
    } else {
        std::atomic<int> next(0), pending(workers);
        auto drain = [&]() { for (int e = next.fetch_add(1); e < n_eig; e = next.fetch_add(1)) one(e); };
        std::mutex mu;
        std::condition_variable cv;
        for (int t = 0; t < workers; ++t)
            Pool::instance().submit([&]() {
                drain();
                std::lock_guard<std::mutex> lock(mu);
                if (--pending == 0) cv.notify_one();
            });
        drain();
        // the helpers finish within a decomposition's time of this thread: poll briefly before sleeping on the condition
        // variable (a sleep + wake-up costs ~50 us of a ~200 us call)
        for (int spin = 0; spin < 200000 && pending.load(std::memory_order_acquire) != 0; ++spin) {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
        std::unique_lock<std::mutex> lock(mu);
        cv.wait(lock, [&]() { return pending.load() == 0; });
    }
    for (int e = 0; e < n_eig; ++e)
        if (rc[e]) {
            if (msg) *msg = msgs[e];
            return 1;
        }
    return 0;
}

}  // namespace smcb

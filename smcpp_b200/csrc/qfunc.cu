// smcpp_b200 -- the M-step objective on the device: HMM::Q / InferenceManager::Q of the reference
// (src/hmm.cpp:155-193, src/inference_manager.cpp:116-126) evaluated from the E-step statistics that are already
// resident on the GPU, values and -- given the derivative arrays of pi / emission table / transition, which the host's
// autodiff produces -- the gradient:
//     q0 = sum_m log pi_m gamma0_m                         dq0/dp = sum_m (dpi_m/dp / pi_m) gamma0_m
//     q1, q2 = sum_{key, m} log e_key(m) gamma_sums_key(m)  split by nb(key) == 0 / > 0, keys present in the contig only
//     q3 = sum_{ij} log T_ij xisum_ij
// Per contig every term is accumulated in the reference's element order with its doubly compensated summation
// (include/common.h:27-46; q0 is a plain sum there, src/hmm.cpp:161), then the contigs are added in order
// (src/inference_manager.cpp:121-125).  One thread per (contig, term, derivative slot): the sums are short
// (M, K M, M^2 elements) and order-dependent.
#include "estep_kernels.cuh"

namespace smcb {

struct Dcs {            // doubly compensated summation, reference include/common.h:27-46
    double s = 0.0, c = 0.0;
    bool first = true;
    __device__ void add(double x)
    {
        if (first) { s = x; c = 0.0; first = false; return; }
        const double y = c + x;
        const double u = x - (y - c);
        const double t = y + s;
        const double v = y - (t - s);
        const double z = u + v;
        s = t + z;
        c = z - (s - t);
    }
    __device__ double value() const { return first ? 0.0 : s; }
};

// out[(c * 4 + term) * (1 + D) + slot]; slot 0 = value, slot 1 + p = derivative p
__global__ void k_q_terms(int C, int M, int K, int D, const double *pi, const double *T, const double *E, const double *dpi,
                          const double *dT, const double *dE, const uint8_t *present, const int32_t *key_nb,
                          const double *gamma0, const double *xisum, const double *gamma_sums, double *out)
{
    const long x = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const long n = (long)C * 4 * (1 + D);
    if (x >= n) return;
    const int slot = (int)(x % (1 + D));
    const int term = (int)((x / (1 + D)) % 4);
    const int c = (int)(x / ((long)(1 + D) * 4));
    const int p = slot - 1;
    double r;
    if (term == 0) {
        // (pi.array().log() * gamma.col(0).array()).sum(): plain, in order (Eigen does not vectorise adouble)
        const double *g0 = gamma0 + (size_t)c * M;
        double acc = 0.0;
        for (int m = 0; m < M; ++m) {
            const double w = slot == 0 ? log(pi[m]) : dpi[(size_t)p * M + m] / pi[m];
            acc = m == 0 ? w * g0[m] : acc + w * g0[m];
        }
        r = acc;
    } else if (term == 3) {
        // prod = log_T.cwiseProduct(xisum), summed over prod.data(): column-major, i fastest
        const double *xs = xisum + (size_t)c * M * M;
        Dcs acc;
        for (int j = 0; j < M; ++j)
            for (int i = 0; i < M; ++i) {
                const double t = T[(size_t)i * M + j];
                const double w = slot == 0 ? log(t) : dT[((size_t)p * M + i) * M + j] / t;
                acc.add(w * xs[(size_t)i * M + j]);
            }
        r = acc.value();
    } else {
        // keys in std::map order; class 0: nb == 0 (term 1), class 1: nb > 0 (term 2)
        const int cls = term - 1;
        Dcs acc;
        bool bad = false;
        for (int k = 0; k < K && !bad; ++k) {
            if (!present[(size_t)c * K + k]) continue;
            const double *e = E + (size_t)k * M;
            double mn = e[0];
            for (int m = 1; m < M; ++m) mn = fmin(mn, e[m]);
            if (mn <= 0.0) { bad = (key_nb[k] > 0) == (cls == 1); break; }   // the reference stops at the first such key (src/hmm.cpp:172-178)
            if ((key_nb[k] > 0) != (cls == 1)) continue;
            const double *gs = gamma_sums + ((size_t)c * K + k) * M;
            for (int m = 0; m < M; ++m) {
                const double w = slot == 0 ? log(e[m]) : dE[((size_t)p * K + k) * M + m] / e[m];
                acc.add(w * gs[m]);
            }
        }
        // zeros in an emission vector: the reference warns and means -infinity for this class
        r = bad ? (slot == 0 ? -INFINITY : 0.0) : acc.value();
    }
    out[x] = r;
}

// sum over contigs in order (src/inference_manager.cpp:121-125): q[(term) * (1 + D) + slot]
__global__ void k_q_sum(int C, int D, const double *terms, double *q)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = 4 * (1 + D);
    if (x >= n) return;
    double acc = 0.0;
    for (int c = 0; c < C; ++c) acc += terms[(size_t)c * n + x];
    q[x] = acc;
}

void launch_q(int C, int M, int K, int D, const double *pi, const double *T, const double *E, const double *dpi, const double *dT,
              const double *dE, const uint8_t *present, const int32_t *key_nb, const double *gamma0, const double *xisum,
              const double *gamma_sums, double *terms, double *q, cudaStream_t st)
{
    const long n = (long)C * 4 * (1 + D);
    k_q_terms<<<(int)((n + 63) / 64), 64, 0, st>>>(C, M, K, D, pi, T, E, dpi, dT, dE, present, key_nb, gamma0, xisum, gamma_sums, terms);
    k_q_sum<<<(4 * (1 + D) + 63) / 64, 64, 0, st>>>(C, D, terms, q);
}

}  // namespace smcb

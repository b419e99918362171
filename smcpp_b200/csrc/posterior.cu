// smcpp_b200 -- full posterior decoding (`save_gamma`, the path of `smc++ posterior`): gamma[:, l] for every block.
//
// Reference: HMM::Estep with *(ib->saveGamma) (src/hmm.cpp:48-49, 116-121, 134-136, 147-150):
//   span-1 block l : v = alpha_l o beta_l / sum(alpha_l o beta_l)
//   span>1 block l : v = span |diag(P_r diag(d_r) Q_r Pinv_r)| / sum|.|,  Q_r = (u w^T) o sq_span,
//                    u = Pinv_r alpha_{l-1}, w = P_r^T beta_l                      (O(M^3) per block, inherent)
//   column 0       : alpha_0 o beta_0
// Runs after the recursions (alpha_hat columns and the beta / w vectors are already in HBM): one warp per block,
//   v_i = C sum_b w_b Pinv(b,i) [ sum_a P(i,a) d_a u_a sq(a,b) ],  C = 1 / (scale sum_a d~_a^s u_a w_a),
// with sq(a,b) = (d~_a^s - d~_b^s) / (d~_a - d~_b), sq(a,a) = s d~_a^(s-1), from the per-E-step power table.
#include "device_utils.cuh"
#include "estep_kernels.cuh"

namespace smcb {

constexpr int kPostWarps = 4;

template <int R>
__global__ void __launch_bounds__(kPostWarps * 32) k_posterior(Model m, Plan p, Work w, double *gamma, const int64_t *gcol_off, int normalise)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = m.M, Mp = m.Mp;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *sm = reinterpret_cast<double *>(smem_raw) + (size_t)warp * 5 * Mp;
    double *s_x = sm, *s_h = sm + Mp, *s_pw = sm + 2 * Mp, *s_w = sm + 3 * Mp, *s_c = sm + 4 * Mp;
    const int64_t nwork = p.total_blocks + p.n_contigs;   // one column per block + column 0 per contig
    const int Lc = p.chunk_blocks;
    for (int64_t item = (int64_t)blockIdx.x * kPostWarps + warp; item < nwork; item += (int64_t)gridDim.x * kPostWarps) {
        // item -> (contig t, column l): columns are laid out contig after contig, L_t + 1 each
        int t = 0;
        {
            int lo = 0, hi = p.n_contigs - 1;
            while (lo < hi) {   // largest t with gcol_off[t] <= item
                const int mid = (lo + hi + 1) >> 1;
                if (gcol_off[mid] <= item) lo = mid; else hi = mid - 1;
            }
            t = lo;
        }
        const int l = (int)(item - gcol_off[t]);
        const int64_t g0 = p.blk_off[t];
        double *out = gamma + item * M;
        const int64_t colbase = p.col_off[t];
        if (l == 0) {   // alpha_hat_0 o beta_0, reference src/hmm.cpp:150
            const float *a0 = w.alpha + colbase * Mp;
            const double *b0 = w.beta_out + (size_t)p.chunk_off[t] * Mp;
            double v0[R], tot = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int j = lane + 32 * r;
                v0[r] = j < M ? (double)a0[j] * b0[j] : 0.0;
                tot += v0[r];
            }
            // normalise: every column divided by its sum, as `smc++ posterior` does on the host (smcpp/commands/posterior.py:104-106)
            const double inv0 = normalise ? 1.0 / warp_sum(tot) : 1.0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int j = lane + 32 * r;
                if (j < M) out[j] = normalise ? v0[r] * inv0 : v0[r];
            }
            continue;
        }
        const int b = l - 1;
        const int cb = b / Lc;
        const float *ap = w.alpha + (colbase + (int64_t)cb * (Lc + 1) + (b - cb * Lc)) * Mp;   // alpha_{l-1}; alpha_l = ap + Mp
        const double *bv = w.bvec + (size_t)(g0 + b) * Mp;
        const int kc = p.kcode[g0 + b];
        const int e = (kc >> kKeyBits) - 1;
        if (e < 0) {
            double v[R], part = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int j = lane + 32 * r;
                v[r] = (double)ap[Mp + j] * bv[j];
                part += v[r];
            }
            const double pp = warp_sum(part);
            double tot = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) { v[r] = v[r] / pp; tot += lane + 32 * r < M ? v[r] : 0.0; }
            const double inv1 = normalise ? 1.0 / warp_sum(tot) : 1.0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int j = lane + 32 * r;
                if (j < M) out[j] = normalise ? v[r] * inv1 : v[r];
            }
            continue;
        }
        const int span = p.span[g0 + b];
        const double *pwr = m.pwtab + ((size_t)e * m.n_span + p.span_id[g0 + b]) * Mp;
        const double *PinvT = m.PinvT + (size_t)e * Mp * Mp, *PT = m.PT + (size_t)e * Mp * Mp, *Pinv = m.Pinv + (size_t)e * Mp * Mp;
        const double *idf = m.invdiff + (size_t)e * Mp * Mp;
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; ++r) s_x[lane + 32 * r] = (double)ap[lane + 32 * r];
        __syncwarp();
        double u[R], pw[R], wv[R], part = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = 0.0;
        for (int i = 0; i < M; ++i) {
            const double xi = s_x[i];
#pragma unroll
            for (int r = 0; r < R; ++r) u[r] = fma(__ldg(PinvT + (size_t)i * Mp + lane + 32 * r), xi, u[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int a = lane + 32 * r;
            pw[r] = pwr[a];
            wv[r] = bv[a];
            part += pw[r] * u[r] * wv[r];
            s_h[a] = m.dr[(size_t)e * Mp + a] * u[r];
            s_pw[a] = pw[r];
            s_w[a] = wv[r];
        }
        const double C = 1.0 / (m.scale[e] * warp_sum(part));
        __syncwarp();
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.0;
        for (int bb = 0; bb < M; ++bb) {
            const double pwb = s_pw[bb];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int a = lane + 32 * r;
                double sq;
                if (a == bb) {
                    const double da = m.dsc[(size_t)e * Mp + a];
                    sq = da != 0.0 ? (double)span * pw[r] / da : 0.0;
                } else {
                    sq = (pw[r] - pwb) * __ldg(idf + (size_t)bb * Mp + a);
                }
                s_c[a] = s_h[a] * sq;
            }
            __syncwarp();
            const double wb = s_w[bb];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = lane + 32 * r;
                double gsum = 0.0;
                for (int a = 0; a < M; ++a) gsum = fma(__ldg(PT + (size_t)a * Mp + i), s_c[a], gsum);
                acc[r] = fma(wb * __ldg(Pinv + (size_t)bb * Mp + i), gsum, acc[r]);
            }
            __syncwarp();
        }
        double tot = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r] = fabs(C * acc[r]); tot += lane + 32 * r < M ? acc[r] : 0.0; }   // the reference takes |.| (src/hmm.cpp:116)
        const double inv2 = normalise ? 1.0 / warp_sum(tot) : 1.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = lane + 32 * r;
            if (j < M) out[j] = normalise ? acc[r] * inv2 : acc[r];
        }
    }
}

// 1 / (d~_a - d~_b) tables for the span tables (0 on the diagonal)
__global__ void k_setup_invdiff(Model m)
{
    const int Mp = m.Mp;
    const long n = (long)m.n_eig * Mp * Mp;
    double *t = const_cast<double *>(m.invdiff);
    for (long x = blockIdx.x * (long)blockDim.x + threadIdx.x; x < n; x += (long)gridDim.x * blockDim.x) {
        const int a = (int)(x % Mp), b = (int)((x / Mp) % Mp), e = (int)(x / ((long)Mp * Mp));
        const double da = m.dsc[(size_t)e * Mp + a], db = m.dsc[(size_t)e * Mp + b];
        t[x] = (a != b && a < m.M && b < m.M) ? 1.0 / (da - db) : 0.0;   // [e][b][a]
    }
}

void launch_posterior(const Model &m, const Plan &p, const Work &w, double *gamma, const int64_t *gcol_off, int normalise, int n_sm, cudaStream_t st)
{
    {
        const long n = (long)m.n_eig * m.Mp * m.Mp;
        if (n > 0) k_setup_invdiff<<<(int)((n + 255) / 256), 256, 0, st>>>(m);
    }
    const size_t smem = (size_t)kPostWarps * 5 * m.Mp * sizeof(double);
    const int blocks = n_sm * 8;
    switch (m.Mp / 32) {
    case 1: k_posterior<1><<<blocks, kPostWarps * 32, smem, st>>>(m, p, w, gamma, gcol_off, normalise); break;
    case 2: k_posterior<2><<<blocks, kPostWarps * 32, smem, st>>>(m, p, w, gamma, gcol_off, normalise); break;
    case 3: k_posterior<3><<<blocks, kPostWarps * 32, smem, st>>>(m, p, w, gamma, gcol_off, normalise); break;
    default: k_posterior<4><<<blocks, kPostWarps * 32, smem, st>>>(m, p, w, gamma, gcol_off, normalise); break;
    }
}

}  // namespace smcb

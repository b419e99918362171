// smcpp_b200 -- forward / backward recursions specialised for M <= 32 hidden states (Mp == 32).
//
// One warp per chunk, lane j owns state j.  What makes these fast relative to the generic kernels:
//   * the operands of the dominant ("hot") eigen key -- rows of Pinv_r / P_r (forward), columns of P_r /
//     Pinv_r and the rows of Td (backward) -- live in registers for the whole chunk: a GEMV step is 32 DFMA
//     fed by 16 broadcast LDS.128, no global or per-element address arithmetic;
//   * every inner loop has a compile-time trip count of 32 (tables are zero padded);
//   * (span, key code) of 32 consecutive blocks are fetched by one coalesced load per warp, one batch ahead,
//     and handed out with shuffles, so no dependent global load sits on the per-step critical path;
//   * d~^span of the next step is computed one step ahead, log() of the normalisers is taken once per 8 steps,
//     and the backward pass rescales beta by exact powers of two instead of dividing by its sum every step.
// Numerics are identical to the generic kernels (same reference semantics, reference src/hmm.cpp:58-149).
#include "device_utils.cuh"
#include "estep_kernels.cuh"

namespace smcb {

constexpr int kWarps32 = 4;
constexpr unsigned kFull = 0xffffffffu;

// y = sum_i reg[i] * bcast[i] with bcast read as 16 broadcast double2 from shared memory
__device__ __forceinline__ double dot32_regs(const double (&reg)[32], const double *sm)
{
    const double2 *v2 = reinterpret_cast<const double2 *>(sm);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int q = 0; q < 16; q += 2) {
        const double2 u = v2[q], v = v2[q + 1];
        a0 = fma(reg[2 * q], u.x, a0);
        a1 = fma(reg[2 * q + 1], u.y, a1);
        a2 = fma(reg[2 * q + 2], v.x, a2);
        a3 = fma(reg[2 * q + 3], v.y, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

// same product with the matrix column read through the read-only path (non-hot eigen keys)
__device__ __forceinline__ double dot32_gmem(const double *__restrict__ col, const double *sm)
{
    const double2 *v2 = reinterpret_cast<const double2 *>(sm);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int q = 0; q < 16; q += 2) {
        const double2 u = v2[q], v = v2[q + 1];
        a0 = fma(__ldg(col + (2 * q) * 32), u.x, a0);
        a1 = fma(__ldg(col + (2 * q + 1) * 32), u.y, a1);
        a2 = fma(__ldg(col + (2 * q + 2) * 32), v.x, a2);
        a3 = fma(__ldg(col + (2 * q + 3) * 32), v.y, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

__global__ void __launch_bounds__(kWarps32 * 32) k_forward32(Model m, Plan p, Work w, int pass)
{
    __shared__ __align__(16) double xd_s[kWarps32][32];
    __shared__ __align__(16) float xf_s[kWarps32][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * kWarps32 + warp;
    if (c >= p.n_chunks) return;
    const int M = m.M;
    double *xd = xd_s[warp];
    float *xf = xf_s[warp];

    const int t = p.ch_contig[c], s = p.ch_start[c], len = p.ch_len[c];
    const int64_t g0 = p.blk_off[t];
    const int cl = c - p.chunk_off[t];
    float *acol = w.alpha + (p.col_off[t] + (int64_t)cl * (p.chunk_blocks + 1)) * 32;

    float x;
    int b0;
    if (pass == 0) {
        b0 = s - p.burn_in_fwd;
        if (b0 < 0) b0 = 0;
        x = (float)m.pi[lane];
    } else {
        if (!w.fwd_flag[c]) return;
        b0 = s;
        x = w.end_alpha_prev[(size_t)(c - 1) * 32 + lane];
    }
    if (b0 == s) {
        acol[lane] = x;
        w.start_used[(size_t)c * 32 + lane] = x;
    }
    // hot eigen key operands -> registers
    const int hot = m.hot_eig;
    double pinv_r[32], p_r[32];
    double dsc_h = 0.0, logd_h = 0.0, logscale_h = 0.0;
    if (hot >= 0) {
        const double *PinvT = m.PinvT + (size_t)hot * 1024, *PT = m.PT + (size_t)hot * 1024;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            pinv_r[i] = PinvT[i * 32 + lane];   // Pinv_r(lane, i)
            p_r[i] = PT[i * 32 + lane];         // P_r(lane, i)
        }
        dsc_h = m.dsc[hot * 32 + lane];
        logd_h = m.logd[hot * 32 + lane];
        logscale_h = m.logscale[hot];
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) { pinv_r[i] = 0.0; p_r[i] = 0.0; }
    }
    double llsum = 0.0, lprod = 1.0;   // ll = llsum + log(lprod): one log() per 8 stored steps
    int lcnt = 0;
    double pw_hot = 1.0;                // d~^span of the hot key for the CURRENT step, computed one step ahead
    const int bend = s + len;
    // (span, code) batches of 32 blocks, fetched one batch ahead
    int sp_n = 1, kc_n = 0;
    {
        const int idx = b0 + lane;
        if (idx < bend) { sp_n = p.span[g0 + idx]; kc_n = p.kcode[g0 + idx]; }
    }
    {
        const int sp0 = __shfl_sync(kFull, sp_n, 0), kc0 = __shfl_sync(kFull, kc_n, 0);
        if ((kc0 >> kKeyBits) - 1 == hot && hot >= 0) pw_hot = pow_span(dsc_h, logd_h, sp0);
    }
    for (int base = b0; base < bend; base += 32) {
        const int sp_l = sp_n, kc_l = kc_n;
        sp_n = 1; kc_n = 0;
        {
            const int idx = base + 32 + lane;
            if (idx < bend) { sp_n = p.span[g0 + idx]; kc_n = p.kcode[g0 + idx]; }
        }
        const int cnt = min(32, bend - base);
        for (int tt = 0; tt < cnt; ++tt) {
            const int b = base + tt;
            const int span = __shfl_sync(kFull, sp_l, tt);
            const int kc = __shfl_sync(kFull, kc_l, tt);
            const int k = kc & kKeyMask, e = (kc >> kKeyBits) - 1;
            // d~^span of the NEXT step (independent of this step's dependency chain)
            double pw_next = 1.0;
            {
                const int spx = tt + 1 < 32 ? __shfl_sync(kFull, sp_l, (tt + 1) & 31) : __shfl_sync(kFull, sp_n, 0);
                const int kcx = tt + 1 < 32 ? __shfl_sync(kFull, kc_l, (tt + 1) & 31) : __shfl_sync(kFull, kc_n, 0);
                if (hot >= 0 && (kcx >> kKeyBits) - 1 == hot) pw_next = pow_span(dsc_h, logd_h, spx);
            }
            double cmul, cadd = 0.0;    // this step's normaliser = cmul * exp(cadd)
            float sf = 0.f;
            if (e >= 0) {
                // a = P_r (d~^span o (Pinv_r alpha_prev)); reference src/hmm.cpp:74-80
                __syncwarp();
                xd[lane] = (double)x;
                __syncwarp();
                double u, a, pw, lsc;
                if (e == hot) {
                    u = dot32_regs(pinv_r, xd);
                    pw = pw_hot;
                    lsc = logscale_h;
                } else {
                    u = dot32_gmem(m.PinvT + (size_t)e * 1024 + lane, xd);
                    pw = pow_span(m.dsc[e * 32 + lane], m.logd[e * 32 + lane], span);
                    lsc = m.logscale[e];
                }
                if (b >= s) w.uvec[(size_t)(g0 + b) * 32 + lane] = u;   // operand of the statistics pass (stats32.cu)
                __syncwarp();
                xd[lane] = pw * u;
                __syncwarp();
                a = (e == hot) ? dot32_regs(p_r, xd) : dot32_gmem(m.PT + (size_t)e * 1024 + lane, xd);
                __syncwarp();
                xd[lane] = a;
                __syncwarp();
                const double ssum = eigen_sum_f64(xd, M);     // a.sum() in the reference's order
                cmul = ssum;
                cadd = (double)span * lsc;
                x = (float)(a / ssum);
            } else {
                // float GEMV, k-sequential axpy order with the float-rounded matrix; reference src/hmm.cpp:85-89
                __syncwarp();
                xf[lane] = x;
                __syncwarp();
                const float4 *x4 = reinterpret_cast<const float4 *>(xf);
                const float *A = m.A32 + (size_t)k * 1024 + lane;
                float y = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 xv = x4[q];
                    y = __fadd_rn(y, __fmul_rn(xv.x, __ldg(A + (4 * q) * 32)));
                    y = __fadd_rn(y, __fmul_rn(xv.y, __ldg(A + (4 * q + 1) * 32)));
                    y = __fadd_rn(y, __fmul_rn(xv.z, __ldg(A + (4 * q + 2) * 32)));
                    y = __fadd_rn(y, __fmul_rn(xv.w, __ldg(A + (4 * q + 3) * 32)));
                }
                __syncwarp();
                xf[lane] = y;
                __syncwarp();
                sf = eigen_sum_f32(xf, M, (M & 3) ? (int)((4 - (((long)(b + 1) * M) & 3)) & 3) : 0);
                cmul = (double)sf;
                x = __fdiv_rn(y, sf);
            }
            if (lane < M && x < 1e-10f) x = 1e-10f;  // reference src/hmm.cpp:92-94
            if (b >= s) {
                acol[(size_t)(b - s + 1) * 32 + lane] = x;
                lprod *= cmul;
                llsum += cadd;
                if (++lcnt == 8 || !(lprod > 1e-200)) {   // also catches NaN
                    llsum += log(lprod);
                    lprod = 1.0;
                    lcnt = 0;
                }
                if (e < 0 && lane == 0) w.cnorm[g0 + b] = sf;
            } else if (b == s - 1) {
                acol[lane] = x;
                w.start_used[(size_t)c * 32 + lane] = x;
            }
            pw_hot = pw_next;
        }
    }
    w.end_alpha[(size_t)c * 32 + lane] = x;
    if (lane == 0) w.ll_chunk[c] = llsum + log(lprod);
}

__global__ void __launch_bounds__(kWarps32 * 32) k_backward32(Model m, Plan p, Work w, int pass)
{
    __shared__ __align__(16) double xd_s[kWarps32][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * kWarps32 + warp;
    if (c >= p.n_chunks) return;
    const int M = m.M;
    double *xd = xd_s[warp];

    const int t = p.ch_contig[c], s = p.ch_start[c], len = p.ch_len[c];
    const int64_t g0 = p.blk_off[t];
    const int L = (int)(p.blk_off[t + 1] - g0);
    const int bend = s + len;
    double beta;
    int b1;
    if (pass == 0) {
        b1 = bend + p.burn_in;
        if (b1 > L || bend == L) b1 = L;
        beta = lane < M ? 1.0 : 0.0;  // reference src/hmm.cpp:97
    } else {
        if (!w.bwd_flag[c]) return;
        b1 = bend;
        beta = w.beta_out_prev[(size_t)(c + 1) * 32 + lane];
    }
    // operands -> registers: row `lane` of Td; column `lane` of P_r and of Pinv_r for the hot eigen key
    const int hot = m.hot_eig;
    double td_r[32], pc_r[32], pic_r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) td_r[j] = m.TdT[j * 32 + lane];      // Td(lane, j)
    double dsc_h = 0.0, logd_h = 0.0;
    if (hot >= 0) {
        const double *P = m.P + (size_t)hot * 1024, *Pinv = m.Pinv + (size_t)hot * 1024;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            pc_r[i] = P[i * 32 + lane];      // P_r(i, lane)
            pic_r[i] = Pinv[i * 32 + lane];  // Pinv_r(i, lane)
        }
        dsc_h = m.dsc[hot * 32 + lane];
        logd_h = m.logd[hot * 32 + lane];
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) { pc_r[i] = 0.0; pic_r[i] = 0.0; }
    }
    // beta is kept only LOOSELY normalised inside the chunk: every statistic that consumes it is invariant to
    // its scale (k_stats divides by alpha.beta resp. by sum_a d~^s u_a w_a), so instead of the reference's
    // beta /= beta.sum() per step (src/hmm.cpp:142) we rescale by an exact power of two every 4 steps and
    // normalise properly where the value is compared or exported (chunk boundaries).
    int sp_n = 1, kc_n = 0;
    {
        const int idx = b1 - 1 - lane;
        if (idx >= s) { sp_n = p.span[g0 + idx]; kc_n = p.kcode[g0 + idx]; }
    }
    double pw_hot = 1.0;
    {
        const int sp0 = __shfl_sync(kFull, sp_n, 0), kc0 = __shfl_sync(kFull, kc_n, 0);
        if ((kc0 >> kKeyBits) - 1 == hot && hot >= 0) pw_hot = pow_span(dsc_h, logd_h, sp0);
    }
    int since = 0;
    for (int top = b1 - 1; top >= s; top -= 32) {
        const int sp_l = sp_n, kc_l = kc_n;
        sp_n = 1; kc_n = 0;
        {
            const int idx = top - 32 - lane;
            if (idx >= s) { sp_n = p.span[g0 + idx]; kc_n = p.kcode[g0 + idx]; }
        }
        const int cnt = min(32, top - s + 1);
        for (int tt = 0; tt < cnt; ++tt) {
            const int b = top - tt;
            if (b == bend - 1) {
                const double bs = warp_sum(beta);
                beta = beta / bs;
                w.bstart_used[(size_t)c * 32 + lane] = beta;
            }
            const bool storing = b < bend;
            const int span = __shfl_sync(kFull, sp_l, tt);
            const int kc = __shfl_sync(kFull, kc_l, tt);
            const int k = kc & kKeyMask, e = (kc >> kKeyBits) - 1;
            double pw_next = 1.0;
            {
                const int spx = tt + 1 < 32 ? __shfl_sync(kFull, sp_l, (tt + 1) & 31) : __shfl_sync(kFull, sp_n, 0);
                const int kcx = tt + 1 < 32 ? __shfl_sync(kFull, kc_l, (tt + 1) & 31) : __shfl_sync(kFull, kc_n, 0);
                if (hot >= 0 && (kcx >> kKeyBits) - 1 == hot) pw_next = pow_span(dsc_h, logd_h, spx);
            }
            double *bv = w.bvec + (size_t)(g0 + b) * 32;
            double nb;
            if (e >= 0) {
                // beta <- Pinv_r^T (d~^span o (P_r^T beta)); reference src/hmm.cpp:123-127
                __syncwarp();
                xd[lane] = beta;
                __syncwarp();
                double wv, pw;
                if (e == hot) {
                    wv = dot32_regs(pc_r, xd);
                    pw = pw_hot;
                } else {
                    wv = dot32_gmem(m.P + (size_t)e * 1024 + lane, xd);
                    pw = pow_span(m.dsc[e * 32 + lane], m.logd[e * 32 + lane], span);
                }
                if (storing) bv[lane] = wv;
                __syncwarp();
                xd[lane] = pw * wv;
                __syncwarp();
                nb = (e == hot) ? dot32_regs(pic_r, xd) : dot32_gmem(m.Pinv + (size_t)e * 1024 + lane, xd);
            } else {
                // beta <- Td (e_k o beta); reference src/hmm.cpp:139
                if (storing) bv[lane] = beta;
                const double ek = __ldg(m.E + (size_t)k * 32 + lane);
                __syncwarp();
                xd[lane] = ek * beta;
                __syncwarp();
                nb = dot32_regs(td_r, xd);
            }
            beta = nb;
            if (++since == 4 || __all_sync(kFull, !(fabs(nb) > 1e-100))) {
                beta *= pow2_rescale(warp_sum(nb));
                since = 0;
            }
            pw_hot = pw_next;
        }
    }
    {
        const double bs = warp_sum(beta);
        w.beta_out[(size_t)c * 32 + lane] = beta / bs;
    }
}

// chunks (= warps) of one recursion kernel that are resident at once on the whole GPU
int resident_warps32(int n_sm)
{
    int bf = 0, bb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bf, k_forward32, kWarps32 * 32, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bb, k_backward32, kWarps32 * 32, 0);
    int b = bf < bb ? bf : bb;
    if (b < 1) b = 1;
    return n_sm * b * kWarps32;
}

void launch_forward32(const Model &m, const Plan &p, const Work &w, int pass, cudaStream_t st)
{
    const int blocks = (p.n_chunks + kWarps32 - 1) / kWarps32;
    k_forward32<<<blocks, kWarps32 * 32, 0, st>>>(m, p, w, pass);
}

void launch_backward32(const Model &m, const Plan &p, const Work &w, int pass, cudaStream_t st)
{
    const int blocks = (p.n_chunks + kWarps32 - 1) / kWarps32;
    k_backward32<<<blocks, kWarps32 * 32, 0, st>>>(m, p, w, pass);
}

}  // namespace smcb

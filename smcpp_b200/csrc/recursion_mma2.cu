// smcpp_b200 -- tensor-path recursions with SEVERAL MMA row tiles per warp (M <= 32): a warp advances NM x 8 chunks.
//
// recursion_mma.cu gives a warp 8 chunks (one 8-row tile of the m8n8k4 DMMA).  ncu (profiles/r2a) shows both kernels
// bound by the LSU data pipe -- every DMMA needs its own 256-byte B fragment from shared memory, and every lane pulls
// 1 KB of the float step matrix per span-1 step -- with the FP64 tensor pipe at 27 % (forward) / 57 % (backward).
// Here lane (n, q) owns the chunks n, 8 + n, ... of NM tiles:
//   * a B fragment is loaded ONCE and feeds NM DMMAs (one per tile): half the fragment traffic per chunk step at NM = 2;
//   * the float step loads a row part of the step matrix once when the lane's chunks sit on the same key (the frequent
//     key covers ~80 % of the sites), a predicated second load otherwise;
//   * NM x 4 independent accumulator chains per GEMV instead of 4 keep the tensor pipe fed from one warp.
// Per chunk the arithmetic is the same instruction sequence as in recursion_mma.cu (same fragments, same order), so the
// results are bitwise identical to the one-tile kernels for the same chunking (tests/test_gpu_parity.py).
//
// Layout, round scheduling and numerics: see the header of recursion_mma.cu.
#include "device_utils.cuh"
#include "estep_kernels.cuh"

namespace smcb {

namespace mt {

constexpr int kMW = 4;             // warps per CTA
constexpr unsigned kAll = 0xffffffffu;
constexpr int MP = 32, NI = 8, NT = 4, MM = MP * MP, XS = MP + 4;

__device__ __forceinline__ int st_of(int q, int idx) { return 8 * (idx >> 1) + 2 * q + (idx & 1); }

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// y[t][.] = W x[t][.] for the 8 NM chunks of the warp; F = W in B-fragment order (see recursion_mma.cu: k_setup_frags),
// in shared memory (kShared) or global memory (read-only path).  One fragment load per (kt, nt), NM DMMAs on it.
template <int NM, bool kShared>
__device__ __forceinline__ void gemv_tiles(const double *F, const double (&v)[NM][NI], double (&y)[NM][NI], int lane)
{
    double c[NM][NT][2];
#pragma unroll
    for (int t = 0; t < NM; ++t)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) c[t][nt][0] = c[t][nt][1] = 0.0;
#pragma unroll
    for (int kt = 0; kt < NI; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const double b = kShared ? F[(kt * NT + nt) * 32 + lane] : __ldg(F + (kt * NT + nt) * 32 + lane);
#pragma unroll
            for (int t = 0; t < NM; ++t) dmma(c[t][nt][0], c[t][nt][1], v[t][kt], b);
        }
#pragma unroll
    for (int t = 0; t < NM; ++t)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { y[t][2 * nt] = c[t][nt][0]; y[t][2 * nt + 1] = c[t][nt][1]; }
}

// ---- packed float pairs: see recursion_mma.cu (fl(x a) = fma(x, a, -0), fl(p + y) = fma(p, 1, y), constants as kernel
// parameters so that ptxas cannot fuse the two roundings)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
struct f32x2x4 { f32x2 v[4]; };
__device__ __forceinline__ f32x2x4 ldg256p(const float *p)
{
    f32x2x4 r;
    asm("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]) : "l"(p));
    return r;
}
// the same load predicated IN PLACE: lanes whose predicate is false keep the value they had (and move no data)
__device__ __forceinline__ void ldg256p_if(f32x2x4 &r, const float *p, bool pred)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %5, 0;\n @p ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];\n}"
                 : "+l"(r.v[0]), "+l"(r.v[1]), "+l"(r.v[2]), "+l"(r.v[3])
                 : "l"(p), "r"((int)pred));
}

__device__ __forceinline__ double group_sum(double v)   // over the 4 lanes of a chunk
{
    v += __shfl_xor_sync(kAll, v, 1);
    v += __shfl_xor_sync(kAll, v, 2);
    return v;
}

__device__ __forceinline__ void ldg_if(int &dst, const int32_t *ptr, bool pred)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p ld.global.nc.b32 %0, [%1];\n}" : "+r"(dst) : "l"(ptr), "r"((int)pred));
}
__device__ __forceinline__ void ldg_if(int &dst, const kcode_t *ptr, bool pred)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p ld.global.nc.b32 %0, [%1];\n}" : "+r"(dst) : "l"(ptr), "r"((int)pred));
}

struct ObsBatch {          // (span, span id, code) of 8 consecutive blocks of a chunk: lane q holds blocks q and 4 + q
    int sp_lo, sp_hi, kc_lo, kc_hi, id_lo, id_hi;
};

// warp w of CTA bid owns the chunks ((bid kMW + w) NM + t) G + n,  t < NM, n < G <= 8
template <int NM>
__device__ __forceinline__ int chunk_of(int bid, int warp, int t, int n, int G, int n_chunks)
{
    return n < G ? ((bid * kMW + warp) * NM + t) * G + n : n_chunks;
}

// =============================================== forward ===================================================
template <int NM>
__device__ __forceinline__ void forward_tiles_body(const Model &m, const Plan &p, const Work &w, const int G, const int bid)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sF_Pinv = reinterpret_cast<double *>(smem_raw);       // [MM] + [MM]: B fragments of the hot eigen key
    double *sF_P = sF_Pinv + MM;
    float *s_x = reinterpret_cast<float *>(smem_raw + 2 * MM * sizeof(double));   // [kMW][NM][8][XS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = lane >> 2, q = lane & 3;
    const int hot = m.hot_eig;
    if (hot >= 0)
        for (int x = tid; x < MM; x += kMW * 32) {
            sF_Pinv[x] = m.F_Pinv[(size_t)hot * MM + x];
            sF_P[x] = m.F_P[(size_t)hot * MM + x];
        }
    __syncthreads();
    const int M = m.M;

    int c[NM], s[NM], bend[NM], cur[NM], base[NM], done[NM];
    bool active[NM];
    int64_t g0[NM];
    float *acol[NM], *xs[NM];
    float x[NM][NI];
    ObsBatch ob[NM], obn[NM];
    double llsum[NM], lprod[NM];
    int lcnt[NM];
    int span[NM], kc[NM], sid[NM];
    double2 pwv[NM][NT];

    auto store_col = [&](int t, float *dst) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<float2 *>(dst + 8 * nt + 2 * q) = make_float2(x[t][2 * nt], x[t][2 * nt + 1]);
    };
    // unconditional loads with a clamped index (see recursion_mma.cu); idle lanes read chunk 0's rows
    auto load_batch = [&](int t, int b) {
        ObsBatch o;
        const int64_t i0 = g0[t] + min(b + q, bend[t] - 1), i1 = g0[t] + min(b + 4 + q, bend[t] - 1);
        o.sp_lo = p.span[i0]; o.kc_lo = p.kcode[i0]; o.id_lo = p.span_id[i0];
        o.sp_hi = p.span[i1]; o.kc_hi = p.kcode[i1]; o.id_hi = p.span_id[i1];
        return o;
    };
    auto fetch_cur = [&](int t) {
        const int pos = cur[t] - base[t];
        const int src = (lane & ~3) | (pos & 3);
        span[t] = __shfl_sync(kAll, (pos & 4) ? ob[t].sp_hi : ob[t].sp_lo, src);
        kc[t] = __shfl_sync(kAll, (pos & 4) ? ob[t].kc_hi : ob[t].kc_lo, src);
        sid[t] = __shfl_sync(kAll, (pos & 4) ? ob[t].id_hi : ob[t].id_lo, src);
    };
    auto load_pw = [&](int t) {   // states st(q, 2nt), st(q, 2nt + 1) of the q-major table
        const double2 *pw = reinterpret_cast<const double2 *>(m.pwq + ((size_t)((kc[t] >> kKeyBits) - 1) * m.n_span + sid[t]) * MP + q * NI);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) pwv[t][nt] = __ldg(pw + nt);
    };

#pragma unroll
    for (int t = 0; t < NM; ++t) {
        c[t] = chunk_of<NM>(bid, warp, t, n, G, p.n_chunks);
        active[t] = c[t] < p.n_chunks;
        const int cc = active[t] ? c[t] : 0;
        const int ct = p.ch_contig[cc];
        s[t] = p.ch_start[cc];
        bend[t] = s[t] + p.ch_len[cc];
        g0[t] = p.blk_off[ct];
        acol[t] = w.alpha + (p.col_off[ct] + (int64_t)(cc - p.chunk_off[ct]) * (p.chunk_blocks + 1)) * MP;
        xs[t] = s_x + (((size_t)warp * NM + t) * 8 + n) * XS;
        cur[t] = s[t] - p.burn_in_fwd;
        if (cur[t] < 0) cur[t] = 0;
#pragma unroll
        for (int idx = 0; idx < NI; ++idx) x[t][idx] = (float)m.pi[st_of(q, idx)];   // reference src/hmm.cpp:59 (pads are 0)
        if (active[t] && cur[t] == s[t]) { store_col(t, acol[t]); store_col(t, w.start_used + (size_t)c[t] * MP); }
        base[t] = cur[t];
        ob[t] = load_batch(t, base[t]);
        obn[t] = load_batch(t, base[t] + 8);
        llsum[t] = 0.0; lprod[t] = 1.0; lcnt[t] = 0; done[t] = 0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) pwv[t][nt] = make_double2(0.0, 0.0);
        fetch_cur(t);
        if (active[t] && (kc[t] >> kKeyBits) > 0) load_pw(t);
    }
    int rounds = 0;

    for (;;) {
        bool any = false;
#pragma unroll
        for (int t = 0; t < NM; ++t) any = any || active[t];
        if (!__ballot_sync(kAll, any)) break;
        ++rounds;
        // the least advanced chunk of the warp picks the round's block type
        unsigned key = 0xffffffffu;
#pragma unroll
        for (int t = 0; t < NM; ++t)
            if (active[t]) key = min(key, ((unsigned)done[t] << 8) | ((unsigned)t << 5) | (unsigned)lane);
        const unsigned lead = __reduce_min_sync(kAll, key);
        const int lead_t = (lead >> 5) & 7;
        int lead_type = 0;
#pragma unroll
        for (int t = 0; t < NM; ++t)
            if (t == lead_t) lead_type = active[t] ? (kc[t] >> kKeyBits) : -1;
        const int T = __shfl_sync(kAll, lead_type, lead & 31);
        bool adv[NM];
#pragma unroll
        for (int t = 0; t < NM; ++t) adv[t] = active[t] && (kc[t] >> kKeyBits) == T;

        float xn[NM][NI];
        double cmul[NM], cadd[NM];
        float sf[NM];
        if (T > 0) {
            // a = P_r (d~^span o (Pinv_r alpha_prev)); reference src/hmm.cpp:74-80
            const int e = T - 1;
            double xd[NM][NI], u[NM][NI], a[NM][NI];
#pragma unroll
            for (int t = 0; t < NM; ++t)
#pragma unroll
                for (int idx = 0; idx < NI; ++idx) xd[t][idx] = (double)x[t][idx];
            if (e == hot) gemv_tiles<NM, true>(sF_Pinv, xd, u, lane);
            else gemv_tiles<NM, false>(m.F_Pinv + (size_t)e * MM, xd, u, lane);
#pragma unroll
            for (int t = 0; t < NM; ++t) {
                // u_l = Pinv_r alpha_hat_{l-1} is an operand of the statistics pass (stats32.cu), stored in eigen-index order
                if (adv[t] && cur[t] >= s[t]) {
                    double *ud = w.uvec + (size_t)(g0[t] + cur[t]) * MP + 2 * q;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<double2 *>(ud + 8 * nt) = make_double2(u[t][2 * nt], u[t][2 * nt + 1]);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) { u[t][2 * nt] *= pwv[t][nt].x; u[t][2 * nt + 1] *= pwv[t][nt].y; }
            }
            if (e == hot) gemv_tiles<NM, true>(sF_P, u, a, lane);
            else gemv_tiles<NM, false>(m.F_P + (size_t)e * MM, u, a, lane);
#pragma unroll
            for (int t = 0; t < NM; ++t) {
                double part = 0.0;
#pragma unroll
                for (int idx = 0; idx < NI; ++idx) part += a[t][idx];
                const double ssum = group_sum(part);
                const double rs = 1.0 / ssum;
#pragma unroll
                for (int idx = 0; idx < NI; ++idx) xn[t][idx] = (float)(a[t][idx] * rs);
                cmul[t] = ssum;
                cadd[t] = (double)(adv[t] ? span[t] : 1) * m.logscale[e];
                sf[t] = 0.f;
            }
        } else {
            // float GEMV, k-sequential axpy order with the float-rounded matrix; reference src/hmm.cpp:85-89
            __syncwarp();
#pragma unroll
            for (int t = 0; t < NM; ++t)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<float2 *>(xs[t] + 8 * nt + 2 * q) = make_float2(x[t][2 * nt], x[t][2 * nt + 1]);
            __syncwarp();
            const float *A[NM];
            bool fresh[NM];           // this tile's key differs from tile 0's: its own loads
#pragma unroll
            for (int t = 0; t < NM; ++t) {
                const int k = adv[t] ? (kc[t] & kKeyMask) : (adv[0] ? (kc[0] & kKeyMask) : 0);
                A[t] = m.A32q + ((size_t)k * MP * 4 + q) * NI;      // row i: + i * 4 * NI floats; this lane's NI columns are contiguous
                fresh[t] = t > 0 && A[t] != A[0];
            }
            f32x2 y2[NM][NI / 2];
#pragma unroll
            for (int t = 0; t < NM; ++t)
#pragma unroll
                for (int j = 0; j < NI / 2; ++j) y2[t][j] = 0ull;
            const f32x2 kNegZero2 = m.c_negzero2, kOne2 = m.c_one2;
#pragma unroll 4
            for (int i4 = 0; i4 < MP / 4; ++i4) {
                float4 xv[NM];
#pragma unroll
                for (int t = 0; t < NM; ++t) xv[t] = reinterpret_cast<const float4 *>(xs[t])[i4];
#pragma unroll
                for (int cidx = 0; cidx < 4; ++cidx) {
                    const size_t row = (size_t)(4 * i4 + cidx) * 4 * NI;
                    f32x2x4 av = ldg256p(A[0] + row);
#pragma unroll
                    for (int t = 0; t < NM; ++t) {
                        if (t > 0) ldg256p_if(av, A[t] + row, fresh[t]);   // tiles in ascending order: `av` ends up holding tile t's row
                        const float xi = cidx == 0 ? xv[t].x : cidx == 1 ? xv[t].y : cidx == 2 ? xv[t].z : xv[t].w;
                        const f32x2 xx = pack2(xi, xi);
#pragma unroll
                        for (int j2 = 0; j2 < 4; ++j2)   // two columns per instruction
                            y2[t][j2] = fma2(fma2(xx, av.v[j2], kNegZero2), kOne2, y2[t][j2]);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < NM; ++t) {
                float y[NI];
#pragma unroll
                for (int j = 0; j < NI / 2; ++j) unpack2(y2[t][j], y[2 * j], y[2 * j + 1]);
                if (M == MP) {
                    // Eigen's sum() order for 32 aligned floats (see recursion_mma.cu)
                    float ev = y[0], od = y[1];
#pragma unroll
                    for (int gq = 1; gq < NI / 2; ++gq) { ev = __fadd_rn(ev, y[2 * gq]); od = __fadd_rn(od, y[2 * gq + 1]); }
                    ev = __fadd_rn(ev, __shfl_xor_sync(kAll, ev, 2));
                    od = __fadd_rn(od, __shfl_xor_sync(kAll, od, 2));
                    ev = __fadd_rn(ev, __shfl_xor_sync(kAll, ev, 1));
                    od = __fadd_rn(od, __shfl_xor_sync(kAll, od, 1));
                    sf[t] = __fadd_rn(ev, od);
                } else {
                    __syncwarp();
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<float2 *>(xs[t] + 8 * nt + 2 * q) = make_float2(y[2 * nt], y[2 * nt + 1]);
                    __syncwarp();
                    sf[t] = eigen_sum_f32(xs[t], M, (M & 3) ? (int)((4 - (((long)(cur[t] + 1) * M) & 3)) & 3) : 0);
                }
#pragma unroll
                for (int idx = 0; idx < NI; ++idx) xn[t][idx] = __fdiv_rn(y[idx], sf[t]);
                cmul[t] = (double)sf[t];
                cadd[t] = 0.0;
            }
        }
#pragma unroll
        for (int t = 0; t < NM; ++t) {
            if (adv[t]) {
#pragma unroll
                for (int idx = 0; idx < NI; ++idx) {
                    float v = xn[t][idx];
                    if (st_of(q, idx) < M && v < 1e-10f) v = 1e-10f;     // reference src/hmm.cpp:92-94
                    x[t][idx] = v;
                }
                if (cur[t] >= s[t]) {
                    store_col(t, acol[t] + (size_t)(cur[t] - s[t] + 1) * MP);
                    lprod[t] *= cmul[t];
                    llsum[t] += cadd[t];
                    if (++lcnt[t] == 8 || !(lprod[t] > 1e-200)) { llsum[t] += log(lprod[t]); lprod[t] = 1.0; lcnt[t] = 0; }
                    if (T == 0 && q == 0) w.cnorm[g0[t] + cur[t]] = sf[t];
                } else if (cur[t] == s[t] - 1) {
                    store_col(t, acol[t]);
                    store_col(t, w.start_used + (size_t)c[t] * MP);
                }
                ++cur[t];
                ++done[t];
                if (cur[t] >= bend[t]) active[t] = false;
            }
            {
                const bool rot = adv[t] && active[t] && cur[t] - base[t] == 8;
                if (rot) { base[t] += 8; ob[t] = obn[t]; }
                const int64_t i0 = g0[t] + min(base[t] + 8 + q, bend[t] - 1), i1 = g0[t] + min(base[t] + 12 + q, bend[t] - 1);
                ldg_if(obn[t].sp_lo, p.span + i0, rot); ldg_if(obn[t].kc_lo, p.kcode + i0, rot); ldg_if(obn[t].id_lo, p.span_id + i0, rot);
                ldg_if(obn[t].sp_hi, p.span + i1, rot); ldg_if(obn[t].kc_hi, p.kcode + i1, rot); ldg_if(obn[t].id_hi, p.span_id + i1, rot);
            }
            fetch_cur(t);
            if (adv[t] && active[t] && (kc[t] >> kKeyBits) > 0) load_pw(t);
        }
    }
    int steps = 0;
#pragma unroll
    for (int t = 0; t < NM; ++t) {
        if (c[t] < p.n_chunks) {
            store_col(t, w.end_alpha + (size_t)c[t] * MP);
            if (q == 0) { w.ll_chunk[c[t]] = llsum[t] + log(lprod[t]); steps += done[t]; }
        }
    }
    if (lane == 0) atomicAdd(&w.counters[4], rounds);                 // diagnostics: lockstep efficiency
    if (steps) atomicAdd(&w.counters[5], steps);
}

// =============================================== backward ==================================================
template <int NM>
__device__ __forceinline__ void backward_tiles_body(const Model &m, const Plan &p, const Work &w, const int G, const int bid)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sF_Td = reinterpret_cast<double *>(smem_raw);        // [MM] Td fragments (span-1 rounds)
    double *sF_PT = sF_Td + MM, *sF_PinvT = sF_PT + MM;          // [MM] each: hot eigen key
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = lane >> 2, q = lane & 3;
    const int hot = m.hot_eig;
    for (int x = tid; x < MM; x += kMW * 32) sF_Td[x] = m.F_Td[x];
    if (hot >= 0)
        for (int x = tid; x < MM; x += kMW * 32) { sF_PT[x] = m.F_PT[(size_t)hot * MM + x]; sF_PinvT[x] = m.F_PinvT[(size_t)hot * MM + x]; }
    __syncthreads();
    const int M = m.M;

    int c[NM], s[NM], bend[NM], cur[NM], top[NM], done[NM], since[NM];
    bool active[NM];
    int64_t g0[NM];
    double beta[NM][NI];
    ObsBatch ob[NM], obn[NM];
    int kc[NM], sid[NM];
    double2 opv[NM][NT];

    auto store_vec = [&](double *dst, const double (&v)[NI], double mul) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<double2 *>(dst + 8 * nt + 2 * q) = make_double2(v[2 * nt] * mul, v[2 * nt + 1] * mul);
    };
    auto load_batch = [&](int t, int tp) {   // batch = blocks tp, tp-1, ..., tp-7; lane q holds tp-q and tp-4-q (clamped)
        ObsBatch o;
        const int64_t i0 = g0[t] + max(tp - q, s[t]), i1 = g0[t] + max(tp - 4 - q, s[t]);
        o.sp_lo = o.sp_hi = 1;
        o.kc_lo = p.kcode[i0]; o.id_lo = p.span_id[i0];
        o.kc_hi = p.kcode[i1]; o.id_hi = p.span_id[i1];
        return o;
    };
    auto fetch_cur = [&](int t) {
        const int pos = top[t] - cur[t];
        const int src = (lane & ~3) | (pos & 3);
        kc[t] = __shfl_sync(kAll, (pos & 4) ? ob[t].kc_hi : ob[t].kc_lo, src);
        sid[t] = __shfl_sync(kAll, (pos & 4) ? ob[t].id_hi : ob[t].id_lo, src);
    };
    auto load_op = [&](int t) {   // d~^span (span > 1) or e_k (span 1), both q-major
        const int ty = kc[t] >> kKeyBits;
        const double *row = ty > 0 ? m.pwq + ((size_t)(ty - 1) * m.n_span + sid[t]) * MP : m.Eq + (size_t)(kc[t] & kKeyMask) * MP;
        const double2 *src = reinterpret_cast<const double2 *>(row + q * NI);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) opv[t][nt] = __ldg(src + nt);
    };

#pragma unroll
    for (int t = 0; t < NM; ++t) {
        c[t] = chunk_of<NM>(bid, warp, t, n, G, p.n_chunks);
        active[t] = c[t] < p.n_chunks;
        const int cc = active[t] ? c[t] : 0;
        const int ct = p.ch_contig[cc];
        s[t] = p.ch_start[cc];
        bend[t] = s[t] + p.ch_len[cc];
        g0[t] = p.blk_off[ct];
        const int L = (int)(p.blk_off[ct + 1] - g0[t]);
        int b1 = bend[t] + p.burn_in;
        if (b1 > L || bend[t] == L) b1 = L;
        cur[t] = b1 - 1;                          // block processed next (descending)
#pragma unroll
        for (int idx = 0; idx < NI; ++idx) beta[t][idx] = st_of(q, idx) < M ? 1.0 : 0.0;   // reference src/hmm.cpp:97
        top[t] = cur[t];
        ob[t] = load_batch(t, top[t]);
        obn[t] = load_batch(t, top[t] - 8);
        since[t] = 0; done[t] = 0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) opv[t][nt] = make_double2(0.0, 0.0);
        fetch_cur(t);
        if (active[t]) load_op(t);
    }

    for (;;) {
        bool any = false;
#pragma unroll
        for (int t = 0; t < NM; ++t) any = any || active[t];
        if (!__ballot_sync(kAll, any)) break;
        unsigned key = 0xffffffffu;
#pragma unroll
        for (int t = 0; t < NM; ++t)
            if (active[t]) key = min(key, ((unsigned)done[t] << 8) | ((unsigned)t << 5) | (unsigned)lane);
        const unsigned lead = __reduce_min_sync(kAll, key);
        const int lead_t = (lead >> 5) & 7;
        int lead_type = 0;
#pragma unroll
        for (int t = 0; t < NM; ++t)
            if (t == lead_t) lead_type = active[t] ? (kc[t] >> kKeyBits) : -1;
        const int T = __shfl_sync(kAll, lead_type, lead & 31);
        bool adv[NM], storing[NM];
#pragma unroll
        for (int t = 0; t < NM; ++t) {
            adv[t] = active[t] && (kc[t] >> kKeyBits) == T;
            // the chunk's verified start value: normalised, recorded before the first stored step
            const bool rec = adv[t] && cur[t] == bend[t] - 1;
            if (__any_sync(kAll, rec)) {
                double part = 0.0;
#pragma unroll
                for (int idx = 0; idx < NI; ++idx) part += beta[t][idx];
                const double bs = group_sum(part);
                if (rec) {
                    const double rb = 1.0 / bs;
#pragma unroll
                    for (int idx = 0; idx < NI; ++idx) beta[t][idx] *= rb;
                    store_vec(w.bstart_used + (size_t)c[t] * MP, beta[t], 1.0);
                }
            }
            storing[t] = adv[t] && cur[t] < bend[t];
        }
        double nb[NM][NI];
        if (T > 0) {
            // beta <- Pinv_r^T (d~^span o (P_r^T beta)); reference src/hmm.cpp:123-127
            const int e = T - 1;
            double wv[NM][NI];
            if (e == hot) gemv_tiles<NM, true>(sF_PT, beta, wv, lane);
            else gemv_tiles<NM, false>(m.F_PT + (size_t)e * MM, beta, wv, lane);
#pragma unroll
            for (int t = 0; t < NM; ++t) {
                if (storing[t]) store_vec(w.bvec + (size_t)(g0[t] + cur[t]) * MP, wv[t], 1.0);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) { wv[t][2 * nt] *= opv[t][nt].x; wv[t][2 * nt + 1] *= opv[t][nt].y; }
            }
            if (e == hot) gemv_tiles<NM, true>(sF_PinvT, wv, nb, lane);
            else gemv_tiles<NM, false>(m.F_PinvT + (size_t)e * MM, wv, nb, lane);
        } else {
            // beta <- Td (e_k o beta); reference src/hmm.cpp:139
            double tv[NM][NI];
#pragma unroll
            for (int t = 0; t < NM; ++t) {
                if (storing[t]) store_vec(w.bvec + (size_t)(g0[t] + cur[t]) * MP, beta[t], 1.0);
#pragma unroll
                for (int h = 0; h < NI / 2; ++h) {
                    tv[t][2 * h] = opv[t][h].x * beta[t][2 * h];
                    tv[t][2 * h + 1] = opv[t][h].y * beta[t][2 * h + 1];
                }
            }
            gemv_tiles<NM, true>(sF_Td, tv, nb, lane);
        }
#pragma unroll
        for (int t = 0; t < NM; ++t) {
            // loose normalisation by an exact power of two (every statistic is invariant to beta's scale): at least every
            // fourth step, earlier when the vector has become tiny (exponent of the largest entry, integer compares)
            ++since[t];
            int hi = 0;
#pragma unroll
            for (int idx = 0; idx < NI; ++idx) hi = max(hi, __double2hiint(nb[t][idx]) & 0x7fffffff);
            hi = max(hi, __shfl_xor_sync(kAll, hi, 1));
            hi = max(hi, __shfl_xor_sync(kAll, hi, 2));
            const bool tiny = hi < 0x2b200000;       // largest |entry| < 2^-333 ~ 1e-100 (high word of the double)
            const bool need = adv[t] && (since[t] >= 4 || tiny);
            if (__any_sync(kAll, need)) {
                double part = 0.0;
#pragma unroll
                for (int idx = 0; idx < NI; ++idx) part += nb[t][idx];
                const double f = pow2_rescale(group_sum(part));
                if (need) {
#pragma unroll
                    for (int idx = 0; idx < NI; ++idx) nb[t][idx] *= f;
                    since[t] = 0;
                }
            }
            if (adv[t]) {
#pragma unroll
                for (int idx = 0; idx < NI; ++idx) beta[t][idx] = nb[t][idx];
                --cur[t];
                ++done[t];
                if (cur[t] < s[t]) active[t] = false;
            } else {
                --since[t];
            }
            {
                const bool rot = adv[t] && active[t] && top[t] - cur[t] == 8;
                if (rot) { top[t] -= 8; ob[t] = obn[t]; }
                const int64_t i0 = g0[t] + max(top[t] - 8 - q, s[t]), i1 = g0[t] + max(top[t] - 12 - q, s[t]);
                ldg_if(obn[t].kc_lo, p.kcode + i0, rot); ldg_if(obn[t].id_lo, p.span_id + i0, rot);
                ldg_if(obn[t].kc_hi, p.kcode + i1, rot); ldg_if(obn[t].id_hi, p.span_id + i1, rot);
            }
            fetch_cur(t);
            if (adv[t] && active[t]) load_op(t);
        }
    }
#pragma unroll
    for (int t = 0; t < NM; ++t) {
        double part = 0.0;
#pragma unroll
        for (int idx = 0; idx < NI; ++idx) part += beta[t][idx];
        const double bs = group_sum(c[t] < p.n_chunks ? part : 0.0);
        if (c[t] < p.n_chunks) store_vec(w.beta_out + (size_t)c[t] * MP, beta[t], 1.0 / bs);
    }
}

// backward CTAs first, forward CTAs last (see k_recursions_mma in recursion_mma.cu)
template <int NM>
__global__ void __launch_bounds__(kMW * 32, 2) k_recursions_tiles(Model m, Plan p, Work w, int G, int blocks)
{
    const int bid = blockIdx.x;
    if (bid < blocks) backward_tiles_body<NM>(m, p, w, G, bid);
    else forward_tiles_body<NM>(m, p, w, G, bid - blocks);
}
template <int NM>
__global__ void __launch_bounds__(kMW * 32, 2) k_forward_tiles(Model m, Plan p, Work w, int G) { forward_tiles_body<NM>(m, p, w, G, blockIdx.x); }
template <int NM>
__global__ void __launch_bounds__(kMW * 32, 2) k_backward_tiles(Model m, Plan p, Work w, int G) { backward_tiles_body<NM>(m, p, w, G, blockIdx.x); }

static size_t fwd_smem(int NM) { return 2 * MM * sizeof(double) + (size_t)kMW * NM * 8 * XS * sizeof(float); }
static size_t bwd_smem() { return 3 * MM * sizeof(double); }

}  // namespace mt

// ---- launch -------------------------------------------------------------------------------------------------
static int tiles_G(int n_chunks, int n_sm, int NM, const RecOpts &o)
{
    if (o.force_G == 1 || o.force_G == 2 || o.force_G == 4 || o.force_G == 8) return o.force_G;
    const int want = n_sm * 2;
    int G = 8;
    while (G > 1 && (n_chunks + G * NM - 1) / (G * NM) < want) G >>= 1;
    return G;
}

static void configure_tiles()
{
    static std::atomic<size_t> done[kMaxDevices];
    if (!needs_smem_config(done, 1)) return;
    auto mx = [](size_t a, size_t b) { return a > b ? a : b; };
    auto set = [](auto kernel, size_t smem) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
        if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    };
    set(mt::k_recursions_tiles<2>, mx(mt::fwd_smem(2), mt::bwd_smem()));
    set(mt::k_forward_tiles<2>, mt::fwd_smem(2));
    set(mt::k_backward_tiles<2>, mt::bwd_smem());
}

// chunks a CTA of the multi-tile kernels carries (the planner sizes the chunk count in whole CTA layers)
int tiles_chunks_per_cta(int NM) { return mt::kMW * 8 * NM; }

// Both recursions, 8 NM chunks per warp (Mp == 32, NM == 2).  fused: one launch (backward CTAs first); otherwise the
// caller provides two streams.  Returns false when the configuration is not covered (the caller takes recursion_mma.cu).
bool launch_recursions_tiles(const Model &m, const Plan &p, const Work &w, int n_sm, const RecOpts &o, cudaStream_t st_fwd, cudaStream_t st_bwd)
{
    if (m.Mp != 32 || o.tiles != 2) return false;
    constexpr int NM = 2;
    configure_tiles();
    const int G = tiles_G(p.n_chunks, n_sm, NM, o);
    const int warps = (p.n_chunks + G * NM - 1) / (G * NM), blocks = (warps + mt::kMW - 1) / mt::kMW;
    if (st_fwd == st_bwd) {
        const size_t smem = mt::fwd_smem(NM) > mt::bwd_smem() ? mt::fwd_smem(NM) : mt::bwd_smem();
        mt::k_recursions_tiles<NM><<<2 * blocks, mt::kMW * 32, smem, st_fwd>>>(m, p, w, G, blocks);
    } else {
        mt::k_backward_tiles<NM><<<blocks, mt::kMW * 32, mt::bwd_smem(), st_bwd>>>(m, p, w, G);
        mt::k_forward_tiles<NM><<<blocks, mt::kMW * 32, mt::fwd_smem(NM), st_fwd>>>(m, p, w, G);
    }
    return true;
}

}  // namespace smcb

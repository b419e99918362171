// smcpp_b200 -- sufficient statistics for M <= 32 on the FP64 tensor path (mma.sync.m8n8k4.f64, SASS DMMA).
//
// The per-block rank-1 updates of the E-step are GEMMs over the block index:
//   span-1 blocks:  X    += A_prev B^T          A_prev = [alpha_{l-1}],  B = [beta_l o e_k / (c_l p_l)]
//   span>1 blocks:  U     = Pinv_r A_prev       (one GEMM, blocks in the n dimension)
//                   R_e  += (U o pw) Y^T - U Z^T  Y = [C_l w_l], Z = Y o pw     (blocks in the k dimension)
// k_stats32 walks a slab's blocks in a precomputed order (plan-time permutation: span-1 blocks sorted by key,
// then the span>1 blocks of each eigen key), so every pass is dense -- no masking -- and the per-key gamma
// sums are accumulated in registers over key runs.  One warp owns a 32x32 FP64 accumulator in DMMA C-fragment
// layout; the 8 warps of a CTA reduce through shared memory in fixed order (bitwise reproducible).
//
// Fragment layout of mma.m8n8k4 (lane = 4*r + q, r = lane/4, q = lane%4):
//   A (8x4, row): a = A[r][q]        B (4x8, col): b = B[q][r]        C (8x8): c0 = C[r][2q], c1 = C[r][2q+1]
#include "device_utils.cuh"
#include "estep_kernels.cuh"

namespace smcb {

constexpr int kS32Warps = 8;
constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// The state index of an operand row is free to permute (it only permutes the rows / columns of the accumulator), so a
// lane's four (m-tile) values are the four CONSECUTIVE states 4r..4r+3 -- one 128/256-bit load covering whole lines across
// the warp -- instead of the textbook 8mt + r.  profiles/r1g: this kernel is bound by L1 wavefronts, not by latency.
struct dbl4 { double v[4]; };
__device__ __forceinline__ dbl4 ld4d(const double *p)   // 256-bit read-only load, 32-byte aligned
{
    dbl4 r;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
    return r;
}
struct flt8 { float v[8]; };
__device__ __forceinline__ flt8 ld8f(const float *p)
{
    flt8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
        : "l"(p));
    return r;
}

// sum over the 8 lanes that share q (= over r)
__device__ __forceinline__ double sum_over_r(double v)
{
    v += __shfl_xor_sync(kFullMask, v, 4);
    v += __shfl_xor_sync(kFullMask, v, 8);
    v += __shfl_xor_sync(kFullMask, v, 16);
    return v;
}
// sum over the 4 lanes that share r (= over q)
__device__ __forceinline__ double sum_over_q(double v)
{
    v += __shfl_xor_sync(kFullMask, v, 1);
    v += __shfl_xor_sync(kFullMask, v, 2);
    return v;
}

// accumulator tile -> shared [32][33] (padded): tile (mt, nt) row r is state 4r + mt, its column c is state 4c + nt
__device__ __forceinline__ void store_acc(double *sm, const double (&acc)[4][4][2], int r, int q)
{
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            sm[(4 * r + mt) * 33 + 4 * (2 * q) + nt] = acc[mt][nt][0];
            sm[(4 * r + mt) * 33 + 4 * (2 * q + 1) + nt] = acc[mt][nt][1];
        }
}

__global__ void __launch_bounds__(kS32Warps * 32, 2) k_stats32(Model m, Plan p, Work w)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = m.K, NE = m.n_eig;
    const int slab = blockIdx.x;
    const int t = p.sl_contig[slab], s0 = p.sl_start[slab];
    const uint32_t mask = p.sl_mask[slab];
    const int64_t g0 = p.blk_off[t];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    const int2 *rec = p.srec + g0 + s0;   // this slab's processing order: (block index in contig, key id | span id)
    const int32_t *seg = p.seg + (size_t)slab * (NE + 2);   // [dense | eig 0 | eig 1 | ... ] offsets into perm

    double *tiles = reinterpret_cast<double *>(smem_raw);                   // [kS32Warps][32*33]
    double *gs = tiles + (size_t)kS32Warps * 32 * 33;                       // [K][32]
    double *bnd = gs + (size_t)K * 32;                                      // [kS32Warps][2][32] boundary key sums
    double *dred = bnd + (size_t)kS32Warps * 2 * 32;                        // [kS32Warps][32]
    int *bkey = reinterpret_cast<int *>(dred + (size_t)kS32Warps * 32);     // [kS32Warps][2]
    double *pinv_s = reinterpret_cast<double *>(bkey + kS32Warps * 2);      // [4 mt][8 kt][32 lanes] A fragments of Pinv_r

    const int Lc = p.chunk_blocks;
    const int64_t colbase = p.col_off[t];
    auto alpha_col = [&](int b) -> const float * {   // alpha_hat column "before block b" (the one after it is + 32)
        const int cb = b / Lc;
        return w.alpha + (colbase + (int64_t)cb * (Lc + 1) + (b - cb * Lc)) * 32;
    };

    // ================= span-1 blocks: X and the per-key gamma sums =================
    if (mask & 1u) {
        for (int x = tid; x < K * 32; x += kS32Warps * 32) gs[x] = 0.0;
        if (lane < 2) bkey[warp * 2 + lane] = -1;
        __syncthreads();
        const int d0 = seg[0], d1 = seg[1];
        const int ngrp = (d1 - d0 + 3) >> 2;
        const int gbeg = (int)((long)ngrp * warp / kS32Warps), gend = (int)((long)ngrp * (warp + 1) / kS32Warps);
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        double gacc[4] = {0.0, 0.0, 0.0, 0.0};   // gamma sum of the current key run, states 8nt + r, partial over q
        int cur = -1;
        bool first_open = true;                   // the first key of this warp's range may be shared with the previous warp
        auto flush = [&](int key) {
            if (key < 0) return;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const double v = sum_over_q(gacc[nt]);
                if (q == 0) {
                    if (first_open) bnd[(warp * 2 + 0) * 32 + 4 * r + nt] = v;       // boundary slot: combined in warp order below
                    else gs[(size_t)key * 32 + 4 * r + nt] += v;                      // interior key: this warp is its only writer ... so far
                }
                gacc[nt] = 0.0;
            }
            if (first_open) {
                if (lane == 0) bkey[warp * 2 + 0] = key;
                first_open = false;
            }
        };
        // the (block index, key) record of the next group is fetched one iteration ahead: no dependent load chain
        int2 rnext = (gbeg < gend && d0 + 4 * gbeg + q < d1) ? __ldg(rec + d0 + 4 * gbeg + q) : make_int2(0, -1);
        for (int g = gbeg; g < gend; ++g) {
            const int2 rc = rnext;
            rnext = (g + 1 < gend && d0 + 4 * (g + 1) + q < d1) ? __ldg(rec + d0 + 4 * (g + 1) + q) : make_int2(0, -1);
            const bool valid = rc.y >= 0;
            const int b = rc.x;
            const int64_t gb = g0 + b;
            const int k = rc.y;
            double av[4], bvv[4], vv[4], be[4], ac[4], ek[4];
            double pp = 0.0, cn = 1.0;
            if (valid) {
                const float *ap = alpha_col(b);
                const float4 a4 = __ldg(reinterpret_cast<const float4 *>(ap) + r), c4 = __ldg(reinterpret_cast<const float4 *>(ap + 32) + r);
                const dbl4 b4 = ld4d(w.bvec + (size_t)gb * 32 + 4 * r), e4 = ld4d(m.E + (size_t)k * 32 + 4 * r);
                av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w;
                ac[0] = c4.x; ac[1] = c4.y; ac[2] = c4.z; ac[3] = c4.w;
#pragma unroll
                for (int i = 0; i < 4; ++i) { be[i] = b4.v[i]; ek[i] = e4.v[i]; }
                cn = (double)w.cnorm[gb];
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) { av[i] = 0.0; ac[i] = 0.0; be[i] = 0.0; ek[i] = 0.0; }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) pp = fma(ac[i], be[i], pp);
            pp = sum_over_r(pp);                                      // p = alpha_l . beta_l   (all lanes take part)
            const double inv_p = valid ? 1.0 / pp : 0.0;
            const double inv_cp = inv_p / cn;                         // 1 / (c_l p_l)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                bvv[i] = be[i] * ek[i] * inv_cp;
                vv[i] = ac[i] * be[i] * inv_p;
            }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], av[mt], bvv[nt]);
            // gamma sums by key run (the list is key-sorted, so runs are long)
            const int k0 = __shfl_sync(kFullMask, k, 0), k1 = __shfl_sync(kFullMask, k, 1), k2 = __shfl_sync(kFullMask, k, 2),
                      k3 = __shfl_sync(kFullMask, k, 3);
            if (k0 == cur && k1 == cur && k2 == cur && k3 == cur) {
#pragma unroll
                for (int i = 0; i < 4; ++i) gacc[i] += vv[i];
            } else {
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const int kq = qq == 0 ? k0 : qq == 1 ? k1 : qq == 2 ? k2 : k3;
                    if (kq < 0) continue;
                    if (kq != cur) { flush(cur); cur = kq; }
                    if (q == qq) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) gacc[i] += vv[i];
                    }
                }
            }
        }
        // the last key of the range may be shared with the next warp: boundary slot 1 (slot 0 if it is also the first)
        if (cur >= 0) {
            if (first_open) flush(cur);
            else {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const double v = sum_over_q(gacc[nt]);
                    if (q == 0) bnd[(warp * 2 + 1) * 32 + 4 * r + nt] = v;
                }
                if (lane == 0) bkey[warp * 2 + 1] = cur;
            }
        }
        store_acc(tiles + (size_t)warp * 32 * 33, acc, r, q);
        __syncthreads();
        // fixed-order combination: X tiles, then the boundary key sums in warp order
        double *Xp = w.Xpart + (size_t)slab * 1024;
        for (int x = tid; x < 1024; x += kS32Warps * 32) {
            const int i = x >> 5, j = x & 31;
            double s = 0.0;
#pragma unroll
            for (int ww = 0; ww < kS32Warps; ++ww) s += tiles[(size_t)ww * 32 * 33 + i * 33 + j];
            Xp[x] = s;
        }
        if (tid < 32) {
            for (int ww = 0; ww < kS32Warps; ++ww)
                for (int sl = 0; sl < 2; ++sl) {
                    const int key = bkey[ww * 2 + sl];
                    if (key >= 0) gs[(size_t)key * 32 + tid] += bnd[(ww * 2 + sl) * 32 + tid];
                }
        }
        __syncthreads();
        double *gp = w.gspart + (size_t)slab * K * 32;
        for (int x = tid; x < K * 32; x += kS32Warps * 32) gp[x] = gs[x];
        __syncthreads();
    }

    // ================= span>1 blocks, one pass per eigen key present in the slab =================
    for (int e = 0; e < NE; ++e) {
        if (!(mask & (2u << e))) continue;
        const int l0 = seg[1 + e], l1 = seg[2 + e];
        const int ngrp = (l1 - l0 + 7) >> 3;
        const int gbeg = (int)((long)ngrp * warp / kS32Warps), gend = (int)((long)ngrp * (warp + 1) / kS32Warps);
        // Pinv_r as A fragments in shared memory (permuted index order, see the header): A[a = 4r + mt][i = 8q + kt]
        {
            const double *Pinv = m.Pinv + (size_t)e * 1024;
            for (int x = tid; x < 1024; x += kS32Warps * 32) {
                const int ln = x & 31, kt = (x >> 5) & 7, mt = x >> 8;
                pinv_s[x] = Pinv[(4 * (ln >> 2) + mt) * 32 + 8 * (ln & 3) + kt];   // A[a = 4r + mt][i = 8q + kt]
            }
            __syncthreads();
        }
        double invd[4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const double dv = m.dsc[e * 32 + 4 * r + mt];
            invd[mt] = dv != 0.0 ? 1.0 / dv : 0.0;
        }
        const double sc = m.scale[e];
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        double dacc[4] = {0.0, 0.0, 0.0, 0.0};
        int2 rBn = (gbeg < gend && l0 + 8 * gbeg + r < l1) ? __ldg(rec + l0 + 8 * gbeg + r) : make_int2(-1, 0);
        int2 rHn0 = (gbeg < gend && l0 + 8 * gbeg + 2 * q < l1) ? __ldg(rec + l0 + 8 * gbeg + 2 * q) : make_int2(-1, 0);
        int2 rHn1 = (gbeg < gend && l0 + 8 * gbeg + 2 * q + 1 < l1) ? __ldg(rec + l0 + 8 * gbeg + 2 * q + 1) : make_int2(-1, 0);
        for (int g = gbeg; g < gend; ++g) {
            const int base = l0 + 8 * g;
            const int2 rB = rBn, rH0 = rHn0, rH1 = rHn1;
            {
                const int nb = base + 8;
                const bool more = g + 1 < gend;
                rBn = (more && nb + r < l1) ? __ldg(rec + nb + r) : make_int2(-1, 0);
                rHn0 = (more && nb + 2 * q < l1) ? __ldg(rec + nb + 2 * q) : make_int2(-1, 0);
                rHn1 = (more && nb + 2 * q + 1 < l1) ? __ldg(rec + nb + 2 * q + 1) : make_int2(-1, 0);
            }
            // U = Pinv_r [alpha_prev of 8 blocks]: B[i = 4kt + q][block r]
            double u[4][2];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) u[mt][0] = u[mt][1] = 0.0;
            {
                const bool vr = rB.x >= 0;
                const int br = vr ? rB.x : 0;
                const flt8 a8 = ld8f(alpha_col(br) + 8 * q);   // states 8q .. 8q+7 of block r
#pragma unroll
                for (int kt = 0; kt < 8; ++kt) {
                    const double bfr = vr ? (double)a8.v[kt] : 0.0;
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) dmma884(u[mt][0], u[mt][1], pinv_s[(mt * 8 + kt) * 32 + lane], bfr);
                }
            }
            // per lane: blocks 2q + h, states 8mt + r
            double xs[4][2], ys[4][2], zs[4][2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int2 rc = h == 0 ? rH0 : rH1;
                const bool vb = rc.x >= 0;
                const int64_t gb = g0 + (vb ? rc.x : 0);
                const int span = vb ? __ldg(m.span_list + rc.y) : 2;
                const double *bv = w.bvec + (size_t)gb * 32;
                const double *pwr = m.pwtab + ((size_t)e * m.n_span + (vb ? rc.y : 0)) * 32;
                double wv[4], pw[4], dot = 0.0;
                const dbl4 w4 = ld4d(bv + 4 * r), p4 = ld4d(pwr + 4 * r);   // eigen indices 4r .. 4r+3
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    wv[mt] = vb ? w4.v[mt] : 0.0;
                    pw[mt] = p4.v[mt];
                    dot = fma(pw[mt] * u[mt][h], wv[mt], dot);
                }
                dot = sum_over_r(dot);
                const double C = vb ? 1.0 / (sc * dot) : 0.0;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    const double y = C * wv[mt];
                    xs[mt][h] = u[mt][h] * pw[mt];
                    ys[mt][h] = y;
                    zs[mt][h] = y * pw[mt];
                    dacc[mt] = fma(y * u[mt][h] * (double)span, pw[mt] * invd[mt], dacc[mt]);
                }
            }
            // R += X Y^T - U Z^T over the 8 blocks (two k-tiles: blocks {2q} and {2q+1})
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        dmma884(acc[mt][nt][0], acc[mt][nt][1], xs[mt][h], ys[nt][h]);
                        dmma884(acc[mt][nt][0], acc[mt][nt][1], -u[mt][h], zs[nt][h]);
                    }
        }
        store_acc(tiles + (size_t)warp * 32 * 33, acc, r, q);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const double v = sum_over_q(dacc[mt]);
            if (q == 0) dred[warp * 32 + 4 * r + mt] = v;
        }
        __syncthreads();
        double *Rp = w.Rpart + ((size_t)slab * NE + e) * 1024;
        for (int x = tid; x < 1024; x += kS32Warps * 32) {
            const int i = x >> 5, j = x & 31;
            double s = 0.0;
#pragma unroll
            for (int ww = 0; ww < kS32Warps; ++ww) s += tiles[(size_t)ww * 32 * 33 + i * 33 + j];
            Rp[x] = s;
        }
        if (tid < 32) {
            double s = 0.0;
#pragma unroll
            for (int ww = 0; ww < kS32Warps; ++ww) s += dred[ww * 32 + tid];
            w.dpart[((size_t)slab * NE + e) * 32 + tid] = s;
        }
        __syncthreads();
    }
}

size_t stats32_smem_bytes(const Model &m)
{
    return ((size_t)kS32Warps * 32 * 33 + (size_t)m.K * 32 + (size_t)kS32Warps * 2 * 32 + (size_t)kS32Warps * 32 + 1024) * sizeof(double) +
           (size_t)kS32Warps * 2 * sizeof(int);
}

void launch_stats32(const Model &m, const Plan &p, const Work &w, cudaStream_t st)
{
    const size_t smem = stats32_smem_bytes(m);
    static size_t configured = 0;
    if (configured < smem) {
        cudaFuncSetAttribute(k_stats32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    k_stats32<<<p.n_slabs, kS32Warps * 32, smem, st>>>(m, p, w);
}

}  // namespace smcb

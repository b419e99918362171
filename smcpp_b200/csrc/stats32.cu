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

constexpr int kS32Warps = 4;
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

// gs_in_smem: the per-key gamma sums of the slab ([K][32] doubles) live in shared memory while K <= kS32MaxKeysSmem;
// data sets with more distinct keys (two-population full-SFS data reach ~10^3) accumulate them straight in the slab's
// output rows in global memory -- same code, same order, L2 instead of shared memory.
constexpr int kS32MaxKeysSmem = 288;

__global__ void __launch_bounds__(kS32Warps * 32, 2) k_stats32(Model m, Plan p, Work w, int gs_in_smem)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = m.K, NE = m.n_eig;
    const int slab = blockIdx.x;
    const int t = p.sl_contig[slab], s0 = p.sl_start[slab];
    const bool has_sites = mask_bit(p.sl_mask + (size_t)slab * p.mask_words, 0);
    const int64_t g0 = p.blk_off[t];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    const int2 *rec = p.srec + g0 + s0;   // this slab's processing order: (block index in contig, key id | span id)
    const int32_t *seg = p.seg + (size_t)slab * (NE + 2);   // [dense | eig 0 | eig 1 | ... ] offsets into perm

    double *tiles = reinterpret_cast<double *>(smem_raw);                   // [kS32Warps][32*33]
    double *gs_s = tiles + (size_t)kS32Warps * 32 * 33;                     // [K][32] (only when gs_in_smem)
    double *gs = gs_in_smem ? gs_s : w.gspart + (size_t)slab * K * 32;
    double *bnd = gs_s + (size_t)(gs_in_smem ? K : 0) * 32;                 // [kS32Warps][2][32] boundary key sums
    double *dred = bnd + (size_t)kS32Warps * 2 * 32;                        // [kS32Warps][32]
    int *bkey = reinterpret_cast<int *>(dred + (size_t)kS32Warps * 32);     // [kS32Warps][2]

    const int Lc = p.chunk_blocks;
    const int64_t colbase = p.col_off[t];
    auto alpha_col = [&](int b) -> const float * {   // alpha_hat column "before block b" (the one after it is + 32)
        const int cb = b / Lc;
        return w.alpha + (colbase + (int64_t)cb * (Lc + 1) + (b - cb * Lc)) * 32;
    };

    // ================= span-1 blocks: X and the per-key gamma sums =================
    if (has_sites) {
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
        // Software pipeline, two groups (8 blocks) per iteration: the records run two iterations ahead of the arithmetic and
        // the operands one, so ~4 KB per warp are in flight while the previous groups are on the tensor pipe.
        struct Ops { float4 a4, c4; dbl4 b4, e4; float cn; int k; };      // k < 0: padding entry
        auto load_rec = [&](int g) { return (g < gend && d0 + 4 * g + q < d1) ? __ldg(rec + d0 + 4 * g + q) : make_int2(0, -1); };
        auto load_ops = [&](int2 rc) {
            Ops o;
            o.k = rc.y;
            const int64_t gb = g0 + rc.x;
            const float *ap = alpha_col(rc.x);
            o.a4 = __ldg(reinterpret_cast<const float4 *>(ap) + r);
            o.c4 = __ldg(reinterpret_cast<const float4 *>(ap + 32) + r);
            o.b4 = ld4d(w.bvec + (size_t)gb * 32 + 4 * r);
            o.e4 = ld4d(m.E + (size_t)(rc.y < 0 ? 0 : rc.y) * 32 + 4 * r);
            o.cn = __ldg(w.cnorm + gb);
            return o;
        };
        auto process = [&](const Ops &o) {
            const bool valid = o.k >= 0;
            const int k = o.k;
            double av[4], ac[4], be[4], bvv[4], vv[4];
            av[0] = o.a4.x; av[1] = o.a4.y; av[2] = o.a4.z; av[3] = o.a4.w;
            ac[0] = o.c4.x; ac[1] = o.c4.y; ac[2] = o.c4.z; ac[3] = o.c4.w;
            double pp = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!valid) { av[i] = 0.0; ac[i] = 0.0; }
                be[i] = valid ? o.b4.v[i] : 0.0;
                pp = fma(ac[i], be[i], pp);
            }
            pp = sum_over_r(pp);                                      // p = alpha_l . beta_l   (all lanes take part)
            const double inv_cp = valid ? 1.0 / (pp * (double)o.cn) : 0.0;   // 1 / (c_l p_l): one division per block
            const double inv_p = inv_cp * (double)o.cn;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                bvv[i] = be[i] * o.e4.v[i] * inv_cp;
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
        };
        int2 rcA = load_rec(gbeg + 2), rcB = load_rec(gbeg + 3);
        Ops nA = load_ops(load_rec(gbeg)), nB = load_ops(load_rec(gbeg + 1));
        for (int g = gbeg; g < gend; g += 2) {
            const Ops cA = nA, cB = nB;
            nA = load_ops(rcA);
            nB = load_ops(rcB);
            rcA = load_rec(g + 4);
            rcB = load_rec(g + 5);
            process(cA);
            if (g + 1 < gend) process(cB);
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
        if (gs_in_smem) {
            double *gp = w.gspart + (size_t)slab * K * 32;
            for (int x = tid; x < K * 32; x += kS32Warps * 32) gp[x] = gs[x];
        }
        __syncthreads();
    }

}

// ================= span>1 blocks: R_e and D_e, one CTA per work item (Plan::erec) =================
// An item holds span>1 blocks of ONE (contig, eigen key) in span-id order.  Per block l (eigen coordinates a, b):
//     u = Pinv_r alpha_hat_{l-1}  (stored by the forward pass),   w = P_r^T beta_l  (stored by the backward pass),
//     pw = d~^span,   C = 1 / (scale sum_a pw_a u_a w_a),   y = C w
//     R_e(a, b) += u_a y_b (pw_a - pw_b),      D_e(a) += span d~_a^(span-1) u_a y_a
// Blocks of equal span share pw, so a run of them accumulates G += u y^T (ONE rank-1 DMMA stream, k = blocks) and is
// weighted once when the run ends: R_e += G o (pw_a - pw_b).  Groups of 4 blocks that straddle a run boundary (or
// data whose spans are all different) take the direct rank-2 form G += (u o pw) y^T - u (y o pw)^T and are added
// unweighted.  Each warp keeps its R_e accumulator in shared memory (touched once per run) and G in registers.
constexpr int kSEWarps = 4;

__global__ void __launch_bounds__(kSEWarps * 32, 2) k_stats32e(Model m, Plan p, Work w)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *tiles = reinterpret_cast<double *>(smem_raw);                 // [kSEWarps][32*33]
    double *dred = tiles + (size_t)kSEWarps * 32 * 33;                    // [kSEWarps][32]
    const int item = blockIdx.x;
    const int t = p.it_contig[item], e = p.it_eig[item], n = p.it_len[item];
    const int64_t g0 = p.blk_off[t];
    const int2 *rec = p.erec + p.it_start[item];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    double *tile = tiles + (size_t)warp * 32 * 33;
    for (int x = lane; x < 32 * 33; x += 32) tile[x] = 0.0;
    __syncwarp();

    const int ngrp = (n + 3) >> 2;
    const int gbeg = (int)((long)ngrp * warp / kSEWarps), gend = (int)((long)ngrp * (warp + 1) / kSEWarps);
    const double sc = m.scale[e];
    const double *pwbase = m.pwtab + (size_t)e * m.n_span * 32 + 4 * r;

    double G[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) G[i][j][0] = G[i][j][1] = 0.0;
    double dacc[4] = {0.0, 0.0, 0.0, 0.0};
    double pwrow[4] = {0.0, 0.0, 0.0, 0.0};   // pw[4r .. 4r+3] of the current uniform run
    int mode = 0, run_sid = -1;               // 0: G empty, 1: uniform run of span id run_sid, 2: rank-2 (already weighted)

    // G -> this warp's R_e tile; tile (mt, nt) of the C fragment: row a = 4r + mt, columns b = 4 (2q + h) + nt
    auto fold = [&]() {
        if (mode == 0) return;
        double pc[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) pc[h][nt] = __shfl_sync(kFullMask, pwrow[nt], 4 * (2 * q + h));   // pw[4 (2q + h) + nt]
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double wgt = mode == 1 ? pwrow[mt] - pc[h][nt] : 1.0;
                    double *dst = tile + (4 * r + mt) * 33 + 4 * (2 * q + h) + nt;
                    *dst = fma(G[mt][nt][h], wgt, *dst);
                    G[mt][nt][h] = 0.0;
                }
        mode = 0;
    };

    // software pipeline: records two groups ahead, operands one group ahead (lane (r, q): block q of the group, rows 4r..4r+3)
    struct Ops { dbl4 u, w, pw; double span; int sid; bool valid; };
    auto load_rec = [&](int g) { return (g < gend && 4 * g + q < n) ? __ldg(rec + 4 * g + q) : make_int2(-1, 0); };
    auto load_ops = [&](int2 rc) {
        Ops o;
        o.valid = rc.x >= 0;
        o.sid = rc.y;
        const int64_t gb = g0 + (o.valid ? rc.x : 0);
        o.u = ld4d(w.uvec + (size_t)gb * 32 + 4 * r);
        o.w = ld4d(w.bvec + (size_t)gb * 32 + 4 * r);
        o.pw = ld4d(pwbase + (size_t)rc.y * 32);
        o.span = (double)__ldg(m.span_list + rc.y);
        return o;
    };
    auto process = [&](const Ops &cur) {
        double uv[4], yv[4], pw[4], dot = 0.0;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            uv[mt] = cur.valid ? cur.u.v[mt] : 0.0;
            pw[mt] = cur.pw.v[mt];
            dot = fma(pw[mt] * uv[mt], cur.w.v[mt], dot);
        }
        dot = sum_over_r(dot);
        const double C = cur.valid ? 1.0 / (sc * dot) : 0.0;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            yv[mt] = cur.valid ? C * cur.w.v[mt] : 0.0;
            dacc[mt] = fma(yv[mt] * uv[mt], cur.span * pw[mt], dacc[mt]);   // 1 / d~_a is applied once, after the loop
        }
        const int s0 = __shfl_sync(kFullMask, cur.sid, 0);       // entry 0 of a group is always valid
        const bool uni = __all_sync(kFullMask, !cur.valid || cur.sid == s0);
        if (uni) {
            if (mode != 1 || run_sid != s0) {
                fold();
                mode = 1;
                run_sid = s0;
#pragma unroll
                for (int i = 0; i < 4; ++i) pwrow[i] = __shfl_sync(kFullMask, pw[i], lane & ~3);   // block 0's row
            }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(G[mt][nt][0], G[mt][nt][1], uv[mt], yv[nt]);
        } else {
            if (mode != 2) { fold(); mode = 2; }
            double xv[4], zv[4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) { xv[mt] = uv[mt] * pw[mt]; zv[mt] = yv[mt] * pw[mt]; }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    dmma884(G[mt][nt][0], G[mt][nt][1], xv[mt], yv[nt]);
                    dmma884(G[mt][nt][0], G[mt][nt][1], -uv[mt], zv[nt]);
                }
        }
    };
    int2 rcA = load_rec(gbeg + 2), rcB = load_rec(gbeg + 3);
    Ops nA = load_ops(load_rec(gbeg)), nB = load_ops(load_rec(gbeg + 1));
    for (int g = gbeg; g < gend; g += 2) {
        const Ops cA = nA, cB = nB;
        nA = load_ops(rcA);
        nB = load_ops(rcB);
        rcA = load_rec(g + 4);
        rcB = load_rec(g + 5);
        process(cA);
        if (g + 1 < gend) process(cB);
    }
    fold();
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const double dv = m.dsc[e * 32 + 4 * r + mt];
        const double v = sum_over_q(dacc[mt]) * (dv != 0.0 ? 1.0 / dv : 0.0);
        if (q == 0) dred[warp * 32 + 4 * r + mt] = v;
    }
    __syncthreads();
    double *Rp = w.Ritem + (size_t)item * 1024;
    for (int x = tid; x < 1024; x += kSEWarps * 32) {
        const int i = x >> 5, j = x & 31;
        double s = 0.0;
#pragma unroll
        for (int ww = 0; ww < kSEWarps; ++ww) s += tiles[(size_t)ww * 32 * 33 + i * 33 + j];
        Rp[x] = s;
    }
    if (tid < 32) {
        double s = 0.0;
#pragma unroll
        for (int ww = 0; ww < kSEWarps; ++ww) s += dred[ww * 32 + tid];
        w.ditem[(size_t)item * 32 + tid] = s;
    }
}

// =====================================================================================================================
// M in (32, 128] (NS = Mp / 32 in {2, 4}): the same two kernels with ONE WARP PER 32x32 TILE of the Mp x Mp accumulators.
// A CTA has 4 warps; warp w takes the tiles (ta, tb) = (t / NS, t % NS), t = w, w + 4, ... -- one tile per warp at NS = 2,
// four in sequence at NS = 4 -- and walks ALL groups of its slab / item for each: it reads the full Mp-long vectors for
// the per-block scalars (p = alpha . beta, resp. C) and feeds the tensor pipe with the ta-part of the row operand and the
// tb-part of the column operand; the redundant loads of the four warps hit L1 / L2.  Vector-shaped results (gamma sums,
// D_e) come from the tb = 0 tiles.
// =====================================================================================================================
constexpr int kSTWarps = 4;
constexpr int kSTMaxGsBytes = 64 * 1024;       // per-key gamma sums in shared memory up to this size, else in global memory

template <int NS, typename V>
__device__ __forceinline__ V pick(const V (&a)[NS], int i)      // a[i] without dynamic register indexing
{
    V r = a[0];
#pragma unroll
    for (int x = 1; x < NS; ++x)
        if (x == i) r = a[x];
    return r;
}

template <int NS>
__global__ void __launch_bounds__(kSTWarps * 32, 2) k_statsT(Model m, Plan p, Work w, int gs_in_smem)
{
    constexpr int MP = 32 * NS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = m.K, NE = m.n_eig;
    const int slab = blockIdx.x;
    if (!mask_bit(p.sl_mask + (size_t)slab * p.mask_words, 0)) return;
    const int t = p.sl_contig[slab], s0 = p.sl_start[slab];
    const int64_t g0 = p.blk_off[t];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    const int2 *rec = p.srec + g0 + s0;
    const int32_t *seg = p.seg + (size_t)slab * (NE + 2);
    double *tiles = reinterpret_cast<double *>(smem_raw);                   // [4][32*33]
    double *gs_s = tiles + (size_t)kSTWarps * 32 * 33;                      // [K][MP] when gs_in_smem
    double *gs = gs_in_smem ? gs_s : w.gspart + (size_t)slab * K * MP;
    const int Lc = p.chunk_blocks;
    const int64_t colbase = p.col_off[t];
    auto alpha_col = [&](int b) -> const float * {
        const int cb = b / Lc;
        return w.alpha + (colbase + (int64_t)cb * (Lc + 1) + (b - cb * Lc)) * MP;
    };
    for (int x = tid; x < K * MP; x += kSTWarps * 32) gs[x] = 0.0;
    __syncthreads();
    const int d0 = seg[0], d1 = seg[1];
    const int ngrp = (d1 - d0 + 3) >> 2;
    double *tile = tiles + (size_t)warp * 32 * 33;
    double *Xp = w.Xpart + (size_t)slab * MP * MP;

    for (int tix = warp; tix < NS * NS; tix += kSTWarps) {
        const int ta = tix / NS, tb = tix % NS;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        double gacc[4] = {0.0, 0.0, 0.0, 0.0};
        int cur = -1;
        auto flush = [&](int key) {          // tb == 0 tiles only: the only writer of (key, these 32 states)
            if (key < 0) return;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const double v = sum_over_q(gacc[nt]);
                if (q == 0) gs[(size_t)key * MP + 32 * ta + 4 * r + nt] += v;
                gacc[nt] = 0.0;
            }
        };
        struct Ops { float4 a4; float4 c[NS]; dbl4 b[NS]; dbl4 e4; float cn; int k; };
        auto load_rec = [&](int g) { return (g < ngrp && d0 + 4 * g + q < d1) ? __ldg(rec + d0 + 4 * g + q) : make_int2(0, -1); };
        auto load_ops = [&](int2 rc) {
            Ops o;
            o.k = rc.y;
            const int64_t gb = g0 + rc.x;
            const float4 *ap = reinterpret_cast<const float4 *>(alpha_col(rc.x));
            o.a4 = __ldg(ap + 8 * ta + r);
            const double *bv = w.bvec + (size_t)gb * MP;
#pragma unroll
            for (int x = 0; x < NS; ++x) {
                o.c[x] = __ldg(ap + MP / 4 + 8 * x + r);
                o.b[x] = ld4d(bv + 32 * x + 4 * r);
            }
            o.e4 = ld4d(m.E + (size_t)(rc.y < 0 ? 0 : rc.y) * MP + 32 * tb + 4 * r);
            o.cn = __ldg(w.cnorm + gb);
            return o;
        };
        auto process = [&](const Ops &o) {
            const bool valid = o.k >= 0;
            double pp = 0.0;
#pragma unroll
            for (int x = 0; x < NS; ++x) {
                pp = fma((double)o.c[x].x, o.b[x].v[0], pp);
                pp = fma((double)o.c[x].y, o.b[x].v[1], pp);
                pp = fma((double)o.c[x].z, o.b[x].v[2], pp);
                pp = fma((double)o.c[x].w, o.b[x].v[3], pp);
            }
            pp = sum_over_r(valid ? pp : 0.0);
            const double inv_cp = valid ? 1.0 / (pp * (double)o.cn) : 0.0;
            const double inv_p = inv_cp * (double)o.cn;
            const double av[4] = {valid ? o.a4.x : 0.0, valid ? o.a4.y : 0.0, valid ? o.a4.z : 0.0, valid ? o.a4.w : 0.0};
            const dbl4 btb = pick<NS>(o.b, tb);
            double bvv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) bvv[i] = btb.v[i] * o.e4.v[i] * inv_cp;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], av[mt], bvv[nt]);
            if (tb == 0) {
                const dbl4 bta = pick<NS>(o.b, ta);
                const float4 cta = pick<NS>(o.c, ta);
                const double vv[4] = {cta.x * bta.v[0] * inv_p, cta.y * bta.v[1] * inv_p, cta.z * bta.v[2] * inv_p, cta.w * bta.v[3] * inv_p};
                const int k = o.k;
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
        };
        int2 rcN = load_rec(1);
        Ops nxt = load_ops(load_rec(0));
        for (int g = 0; g < ngrp; ++g) {
            const Ops cu = nxt;
            nxt = load_ops(rcN);
            rcN = load_rec(g + 2);
            process(cu);
        }
        if (tb == 0) flush(cur);
        __syncwarp();
        store_acc(tile, acc, r, q);
        __syncwarp();
        for (int x = lane; x < 1024; x += 32) {
            const int i = x >> 5, j = x & 31;
            Xp[(size_t)(32 * ta + i) * MP + 32 * tb + j] = tile[i * 33 + j];
        }
    }
    __syncthreads();
    if (gs_in_smem) {
        double *gp = w.gspart + (size_t)slab * K * MP;
        for (int x = tid; x < K * MP; x += kSTWarps * 32) gp[x] = gs[x];
    }
}

template <int NS>
__global__ void __launch_bounds__(kSTWarps * 32, 2) k_statsTe(Model m, Plan p, Work w)
{
    constexpr int MP = 32 * NS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *tiles = reinterpret_cast<double *>(smem_raw);                 // [4][32*33]
    const int item = blockIdx.x;
    const int t = p.it_contig[item], e = p.it_eig[item], n = p.it_len[item];
    const int64_t g0 = p.blk_off[t];
    const int2 *rec = p.erec + p.it_start[item];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    double *tile = tiles + (size_t)warp * 32 * 33;
    const int ngrp = (n + 3) >> 2;
    const double sc = m.scale[e];
    const double *pwbase = m.pwtab + (size_t)e * m.n_span * MP + 4 * r;
    double *Rp = w.Ritem + (size_t)item * MP * MP;

    for (int tix = warp; tix < NS * NS; tix += kSTWarps) {
        const int ta = tix / NS, tb = tix % NS;
        __syncwarp();
        for (int x = lane; x < 32 * 33; x += 32) tile[x] = 0.0;
        __syncwarp();
        double G[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) G[i][j][0] = G[i][j][1] = 0.0;
        double dacc[4] = {0.0, 0.0, 0.0, 0.0};
        double pwa[4] = {0.0, 0.0, 0.0, 0.0}, pwb[4] = {0.0, 0.0, 0.0, 0.0};   // the run's pw on the tile's rows / columns
        int mode = 0, run_sid = -1;
        auto fold = [&]() {
            if (mode == 0) return;
            double pc[2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) pc[h][nt] = __shfl_sync(kFullMask, pwb[nt], 4 * (2 * q + h));
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double wgt = mode == 1 ? pwa[mt] - pc[h][nt] : 1.0;
                        double *dst = tile + (4 * r + mt) * 33 + 4 * (2 * q + h) + nt;
                        *dst = fma(G[mt][nt][h], wgt, *dst);
                        G[mt][nt][h] = 0.0;
                    }
            mode = 0;
        };
        struct Ops { dbl4 u[NS], w[NS], pw[NS]; double span; int sid; bool valid; };
        auto load_rec = [&](int g) { return (g < ngrp && 4 * g + q < n) ? __ldg(rec + 4 * g + q) : make_int2(-1, 0); };
        auto load_ops = [&](int2 rc) {
            Ops o;
            o.valid = rc.x >= 0;
            o.sid = rc.y;
            const int64_t gb = g0 + (o.valid ? rc.x : 0);
            const double *up = w.uvec + (size_t)gb * MP + 4 * r, *wp = w.bvec + (size_t)gb * MP + 4 * r, *pp = pwbase + (size_t)rc.y * MP;
#pragma unroll
            for (int x = 0; x < NS; ++x) { o.u[x] = ld4d(up + 32 * x); o.w[x] = ld4d(wp + 32 * x); o.pw[x] = ld4d(pp + 32 * x); }
            o.span = (double)__ldg(m.span_list + rc.y);
            return o;
        };
        auto process = [&](const Ops &c) {
            double dot = 0.0;
#pragma unroll
            for (int x = 0; x < NS; ++x)
#pragma unroll
                for (int i = 0; i < 4; ++i) dot = fma(c.pw[x].v[i] * c.u[x].v[i], c.w[x].v[i], dot);
            dot = sum_over_r(c.valid ? dot : 0.0);
            const double C = c.valid ? 1.0 / (sc * dot) : 0.0;
            const dbl4 ua = pick<NS>(c.u, ta), pa4 = pick<NS>(c.pw, ta), wb = pick<NS>(c.w, tb), pb4 = pick<NS>(c.pw, tb);
            double uv[4], yv[4], pa[4], pb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uv[i] = c.valid ? ua.v[i] : 0.0;
                pa[i] = pa4.v[i];
                yv[i] = c.valid ? C * wb.v[i] : 0.0;
                pb[i] = pb4.v[i];
            }
            if (tb == 0) {                       // D_e on this tile's rows: y_a = C w_a with a in the ta part
                const dbl4 wa = pick<NS>(c.w, ta);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double ya = c.valid ? C * wa.v[i] : 0.0;
                    dacc[i] = fma(ya * uv[i], c.span * pa[i], dacc[i]);
                }
            }
            const int s0 = __shfl_sync(kFullMask, c.sid, 0);
            const bool uni = __all_sync(kFullMask, !c.valid || c.sid == s0);
            if (uni) {
                if (mode != 1 || run_sid != s0) {
                    fold();
                    mode = 1;
                    run_sid = s0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        pwa[i] = __shfl_sync(kFullMask, pa[i], lane & ~3);    // block 0's rows
                        pwb[i] = __shfl_sync(kFullMask, pb[i], lane & ~3);
                    }
                }
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) dmma884(G[mt][nt][0], G[mt][nt][1], uv[mt], yv[nt]);
            } else {
                if (mode != 2) { fold(); mode = 2; }
                double xv[4], zv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { xv[i] = uv[i] * pa[i]; zv[i] = yv[i] * pb[i]; }
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        dmma884(G[mt][nt][0], G[mt][nt][1], xv[mt], yv[nt]);
                        dmma884(G[mt][nt][0], G[mt][nt][1], -uv[mt], zv[nt]);
                    }
            }
        };
        if constexpr (NS <= 2) {
            int2 rcN = load_rec(1);
            Ops nxt = load_ops(load_rec(0));
            for (int g = 0; g < ngrp; ++g) {
                const Ops cu = nxt;
                nxt = load_ops(rcN);
                rcN = load_rec(g + 2);
                process(cu);
            }
        } else {
            // 128 states: 3 x 4 x 256-bit operands per block do not fit twice next to the accumulators; only the records run ahead
            int2 rcN = load_rec(0);
            for (int g = 0; g < ngrp; ++g) {
                const int2 rc = rcN;
                rcN = load_rec(g + 1);
                process(load_ops(rc));
            }
        }
        fold();
        __syncwarp();
        for (int x = lane; x < 1024; x += 32) {
            const int i = x >> 5, j = x & 31;
            Rp[(size_t)(32 * ta + i) * MP + 32 * tb + j] = tile[i * 33 + j];
        }
        if (tb == 0) {
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                const double dv = m.dsc[e * MP + 32 * ta + 4 * r + mt];
                const double v = sum_over_q(dacc[mt]) * (dv != 0.0 ? 1.0 / dv : 0.0);
                if (q == 0) w.ditem[(size_t)item * MP + 32 * ta + 4 * r + mt] = v;
            }
        }
    }
}

template <int NS>
static void launch_statsT_impl(const Model &m, const Plan &p, const Work &w, cudaStream_t st, cudaStream_t st_runs)
{
    constexpr int MP = 32 * NS;
    const int in_smem = (size_t)m.K * MP * sizeof(double) <= (size_t)kSTMaxGsBytes ? 1 : 0;
    const size_t smem = ((size_t)kSTWarps * 32 * 33 + (in_smem ? (size_t)m.K * MP : 0)) * sizeof(double);
    static std::atomic<size_t> configured[kMaxDevices];
    if (needs_smem_config(configured, smem)) cudaFuncSetAttribute(k_statsT<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_statsT<NS><<<p.n_slabs, kSTWarps * 32, smem, st>>>(m, p, w, in_smem);
    if (p.n_items > 0) k_statsTe<NS><<<p.n_items, kSTWarps * 32, (size_t)kSTWarps * 32 * 33 * sizeof(double), st_runs>>>(m, p, w);
}

void launch_stats64(const Model &m, const Plan &p, const Work &w, cudaStream_t st, cudaStream_t st_runs)   // Mp in {64, 128}
{
    if (m.Mp == 64) launch_statsT_impl<2>(m, p, w, st, st_runs);
    else launch_statsT_impl<4>(m, p, w, st, st_runs);
}

size_t stats32_smem_bytes(const Model &m)
{
    const size_t kk = m.K <= kS32MaxKeysSmem ? m.K : 0;
    return ((size_t)kS32Warps * 32 * 33 + kk * 32 + (size_t)kS32Warps * 2 * 32 + (size_t)kS32Warps * 32) * sizeof(double) +
           (size_t)kS32Warps * 2 * sizeof(int);
}

void launch_stats32(const Model &m, const Plan &p, const Work &w, cudaStream_t st, cudaStream_t st_runs)
{
    const size_t smem = stats32_smem_bytes(m);
    static std::atomic<size_t> configured[kMaxDevices];
    if (needs_smem_config(configured, smem)) cudaFuncSetAttribute(k_stats32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_stats32<<<p.n_slabs, kS32Warps * 32, smem, st>>>(m, p, w, m.K <= kS32MaxKeysSmem ? 1 : 0);
    if (p.n_items > 0) {
        const size_t smem_e = ((size_t)kSEWarps * 32 * 33 + (size_t)kSEWarps * 32) * sizeof(double);
        static std::atomic<size_t> configured_e[kMaxDevices];
        if (needs_smem_config(configured_e, smem_e)) cudaFuncSetAttribute(k_stats32e, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e);
        k_stats32e<<<p.n_items, kSEWarps * 32, smem_e, st_runs>>>(m, p, w);
    }
}

}  // namespace smcb

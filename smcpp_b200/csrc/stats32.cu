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

// Timing ablations of k_stats32 (tools/ablate.sh; results are wrong by construction): 1 = rows in storage order instead of
// the key-sorted gather, 2 = 4 instead of 16 DMMAs per group, 3 = no operand loads.
#ifndef SMCB_STATS_ABLATE
#define SMCB_STATS_ABLATE 0
#endif

namespace smcb {

constexpr int kS32Warps = 4;
constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// volatile: keeps its place between the steps of a software pipeline
__device__ __forceinline__ void dmma884v(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
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

// 1 / x for a normal, positive x without the slow-path call of the compiler's division (a branch in the middle of the
// pipelined loop): MUFU.RCP64H seed (~2^-23) and three Newton steps, relative error ~1 ulp.
__device__ __forceinline__ double rcp_newton(double x)
{
    double r;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));   // volatile: keeps the compiler from branching around it
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const double e = fma(-x, r, 1.0);
        r = fma(r, e, r);
    }
    return r;
}

// ---- asynchronous global -> shared copies (LDGSTS): the operand pipeline of the statistics kernels.  Operands that wait
// in registers cost occupancy (18 registers per group in flight), and these kernels are bound by the number of resident warps;
// cp.async groups complete in order and need neither registers nor scoreboards.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)    // L2 only (streaming operands)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One ring stage of k_stats32 holds the operands of one group (4 blocks) of one warp; every lane copies exactly the bytes it
// reads back later (no cross-lane hazards).  16-byte chunks are XOR-swizzled so that the quarter-warp phases of LDS.128 (lanes
// 8p .. 8p+7 = two r values x four blocks q) hit eight different chunk banks.
constexpr int kRingDepth = 4;                                  // groups in flight per warp (9 KB)
constexpr int kStageB = 0, kStageA = 1024, kStageC = 1536, kStagePriv = 2048, kStageBytes = 2048 + 32 * 8;

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
constexpr int kS32MaxKeysSmem = 128;
constexpr size_t kS32RingBytes = (size_t)kS32Warps * kRingDepth * kStageBytes;      // 36 KB >= the 4 accumulator tiles (33 KB)
static_assert(kS32RingBytes >= (size_t)kS32Warps * 32 * 33 * sizeof(double), "the accumulator tiles alias the operand ring");

__global__ void __launch_bounds__(kS32Warps * 32, 4) k_stats32(Model m, Plan p, Work w, int gs_in_smem)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = m.K, NE = m.n_eig;
    const int slab = blockIdx.x;
    const int t = p.sl_contig[slab], s0 = p.sl_start[slab];
    const bool has_sites = mask_bit(p.sl_mask + (size_t)slab * p.mask_words, 0);
    const int64_t g0 = p.blk_off[t];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(kFullMask, tid >> 5, 0);   // broadcast: the compiler then knows the loop bounds are warp-uniform
    const int r = lane >> 2, q = lane & 3;
    const int2 *rec = p.srec + g0 + s0;   // this slab's processing order: (block index in contig, key id | span id)
    const int32_t *seg = p.seg + (size_t)slab * (NE + 2);   // [dense | eig 0 | eig 1 | ... ] offsets into perm

    unsigned char *ring = smem_raw;                                         // [kS32Warps][kRingDepth][kStageBytes] during the loop,
    double *tiles = reinterpret_cast<double *>(smem_raw);                   // [kS32Warps][32*33] after it
    double *gs_s = reinterpret_cast<double *>(smem_raw + kS32RingBytes);    // [K][32] (only when gs_in_smem)
    double *gs = gs_in_smem ? gs_s : w.gspart + (size_t)slab * K * 32;
    double *bnd = gs_s + (size_t)(gs_in_smem ? K : 0) * 32;                 // [kS32Warps][2][32] boundary key sums
    double *dred = bnd + (size_t)kS32Warps * 2 * 32;                        // [kS32Warps][32]
    int *bkey = reinterpret_cast<int *>(dred + (size_t)kS32Warps * 32);     // [kS32Warps][2]

    const int Lc = p.chunk_blocks;
    const int64_t colbase = p.col_off[t];
    const uint64_t lc_magic = Lc > 1 ? ~0ull / (uint64_t)Lc + 1 : 0;   // b / Lc = umul64hi(b, ceil(2^64 / Lc)) for 0 <= b < 2^31
    auto alpha_col = [&](int b) -> const float * {   // alpha_hat column "before block b" (the one after it is + 32)
        const int cb = Lc > 1 ? (int)__umul64hi((uint64_t)(uint32_t)b, lc_magic) : b;
        return w.alpha + (colbase + b + cb) * 32;      // chunk cb starts at column cb (Lc + 1): one extra column per chunk
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
        // One group = 4 blocks (lane (r, q): block q, states 4r..4r+3).  What bounds this loop (profiles/r2_summary.md, timing
        // ablations): a warp issues in order and is held while its DMMAs and its dependent scalar chain (dot -> butterfly ->
        // reciprocal -> scaling) run, so the time per group and warp is fixed (~1 000 cycles even with no loads at all) and the
        // only lever is the number of resident warps.  Software pipelining inside the warp, 2 vs 4 operand sets in registers,
        // an L2 prefetch, 40 % fewer instructions: none of them moved the 1.7 ms of round 1 at 8 warps per SM.  Hence this
        // shape: the operands wait in a shared-memory ring filled by cp.async (no registers, no scoreboards), the loop body is
        // lean (no validity masks -- only the slab's last group can be partial and takes the masked path below; emission row
        // and key run are per-lane caches; the alpha_hat column index needs no division; branch-free reciprocal), and the
        // kernel fits 128 registers: 16 warps per SM.  1.70 -> 1.47 ms alone on the benchmark.
        struct Raw { float4 a4, c4; dbl4 b4; float cn; int k; };
        struct Prep { float4 a4; double bv[4]; };                         // operands of the 16 DMMAs of one group
        double ecur[4] = {0.0, 0.0, 0.0, 0.0};                            // emission row of this lane's current key
        int ekey = -1;
        const int nfull = (d1 - d0) >> 2;                                 // groups [0, nfull) have four valid entries
        const int gfull = gend < nfull ? gend : nfull;
        auto load_rec = [&](int g) {
            int2 rc = g < gfull ? __ldg(rec + d0 + 4 * g + q) : make_int2(0, 0);
            if (SMCB_STATS_ABLATE == 1) rc.x = s0 + (g < gfull ? 4 * g + q : 0);
            return rc;
        };
        auto load_raw = [&](int2 rc) {                                    // straight into registers (first and last group)
            Raw o;
            o.k = rc.y;
            const int64_t gb = g0 + rc.x;
            const float *ap = alpha_col(rc.x);
            o.a4 = __ldg(reinterpret_cast<const float4 *>(ap) + r);
            o.c4 = __ldg(reinterpret_cast<const float4 *>(ap + 32) + r);
            o.b4 = ld4d(w.bvec + (size_t)gb * 32 + 4 * r);
            o.cn = __ldg(w.cnorm + gb);
            return o;
        };
        // this lane's slots in a ring stage (byte offsets from the stage base)
        const uint32_t ring0 = smem_addr(ring) + (uint32_t)warp * kRingDepth * kStageBytes;
        const int fq = (q & 1) | ((q & 2) << 1);                          // b rows: chunk c of block q lives at chunk c ^ fq
        const uint32_t off_b0 = kStageB + q * 256 + (((2 * r) ^ fq) << 4), off_b1 = kStageB + q * 256 + (((2 * r + 1) ^ fq) << 4);
        const uint32_t off_a = kStageA + q * 128 + ((r ^ (2 * q)) << 4), off_c = off_a + (kStageC - kStageA);
        const uint32_t off_p = kStagePriv + lane * 8;
        auto issue = [&](int2 rc, int stage, bool live) {                 // the group's operands -> ring stage (asynchronous)
            if (live && SMCB_STATS_ABLATE != 3) {
                const uint32_t sb = ring0 + (uint32_t)stage * kStageBytes;
                const int64_t gb = g0 + rc.x;
                const float *ap = alpha_col(rc.x);
                const double *bp = w.bvec + (size_t)gb * 32 + 4 * r;
                cp_async16(sb + off_b0, bp);
                cp_async16(sb + off_b1, bp + 2);
                cp_async16(sb + off_a, ap + 4 * r);
                cp_async16(sb + off_c, ap + 32 + 4 * r);
                cp_async4(sb + off_p, w.cnorm + gb);
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(sb + off_p + 4), "r"(rc.y) : "memory");
            }
            cp_async_commit();
        };
        auto read_stage = [&](int stage) {
            Raw o;
            const uint32_t sb = ring0 + (uint32_t)stage * kStageBytes;
            uint32_t cnb;
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(o.b4.v[0]), "=d"(o.b4.v[1]) : "r"(sb + off_b0) : "memory");
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(o.b4.v[2]), "=d"(o.b4.v[3]) : "r"(sb + off_b1) : "memory");
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o.a4.x), "=f"(o.a4.y), "=f"(o.a4.z), "=f"(o.a4.w) : "r"(sb + off_a) : "memory");
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o.c4.x), "=f"(o.c4.y), "=f"(o.c4.z), "=f"(o.c4.w) : "r"(sb + off_c) : "memory");
            asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(cnb), "=r"(o.k) : "r"(sb + off_p) : "memory");
            o.cn = __uint_as_float(cnb);
            return o;
        };
        struct Scal { double bee[4], acbe[4], pp; };                  // beta o e_k and alpha_l o beta_l of this lane's 4 states
        auto scal_dot = [&](const Raw &o, Scal &c) {                  // local part of p = alpha_l . beta_l
            if (o.k != ekey) {
                const dbl4 e4 = ld4d(m.E + (size_t)o.k * 32 + 4 * r);
#pragma unroll
                for (int i = 0; i < 4; ++i) ecur[i] = e4.v[i];
                ekey = o.k;
            }
            const double ac[4] = {o.c4.x, o.c4.y, o.c4.z, o.c4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                c.bee[i] = o.b4.v[i] * ecur[i];
                c.acbe[i] = ac[i] * o.b4.v[i];
            }
            c.pp = (c.acbe[0] + c.acbe[1]) + (c.acbe[2] + c.acbe[3]);
        };
        auto scal_finish = [&](const Raw &o, const Scal &c, Prep &P, double (&vv)[4]) {   // 1 / (c_l p_l), scaled operands
            const double cn = (double)o.cn;
            const double inv_cp = rcp_newton(c.pp * cn);              // one reciprocal per block
            const double inv_p = inv_cp * cn;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                P.bv[i] = c.bee[i] * inv_cp;
                vv[i] = c.acbe[i] * inv_p;
            }
            P.a4 = o.a4;
        };
        auto mma_row = [&](const Prep &P, int mt) {                   // 4 of a group's 16 DMMAs
            const double av = mt == 0 ? P.a4.x : mt == 1 ? P.a4.y : mt == 2 ? P.a4.z : P.a4.w;
#pragma unroll
            for (int nt = 0; nt < (SMCB_STATS_ABLATE == 2 ? 1 : 4); ++nt) dmma884v(acc[mt][nt][0], acc[mt][nt][1], av, P.bv[nt]);
        };
        // gamma sums by key run (the list is key-sorted, so runs are long); k < 0: padding entry
        auto gamma_add = [&](int k, const double (&vv)[4]) {
            if (__all_sync(kFullMask, k == cur)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) gacc[i] += vv[i];
                return;
            }
            const int k0 = __shfl_sync(kFullMask, k, 0), k1 = __shfl_sync(kFullMask, k, 1), k2 = __shfl_sync(kFullMask, k, 2),
                      k3 = __shfl_sync(kFullMask, k, 3);
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
        };
        auto mma_all = [&](const Prep &P) {
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) mma_row(P, mt);
        };
        if (gbeg < gfull) {
            // ring: group j lives in stage (j - gbeg) % kRingDepth; kRingDepth groups are in flight ahead of the arithmetic
#pragma unroll
            for (int d = 0; d < kRingDepth; ++d) issue(load_rec(gbeg + d), d, gbeg + d < gfull);
            int2 rcN = load_rec(gbeg + kRingDepth);               // record of the next group to issue
            int stage = 0;
            for (int g = gbeg; g < gfull; ++g) {
                cp_async_wait<kRingDepth - 1>();
                const Raw c = read_stage(stage);
                // The freed stage takes group g + kRingDepth right away.  Same-thread write-after-read: the LDS above and this
                // LDGSTS go through the LSU in program order and the copy lands hundreds of cycles later (racecheck clean).
                // Refilling only after the values were used is 8 % slower (statistics 2.67 -> 2.89 ms): shorter lead.
                issue(rcN, stage, g + kRingDepth < gfull);
                rcN = load_rec(g + kRingDepth + 1);
                stage = (stage + 1) & (kRingDepth - 1);
                Scal sc;
                Prep P;
                double vv[4];
                scal_dot(c, sc);
                sc.pp = sum_over_r(sc.pp);
                scal_finish(c, sc, P, vv);
                mma_all(P);
                gamma_add(c.k, vv);
            }
            cp_async_wait<0>();
        }
        __syncthreads();                              // the ring is reused for the accumulator tiles below
        if (gend > nfull) {
            // the slab's last, partial group (this warp owns it): masked, not pipelined
            const bool valid = d0 + 4 * nfull + q < d1;
            const int2 rc = valid ? __ldg(rec + d0 + 4 * nfull + q) : make_int2(0, 0);
            const Raw o = load_raw(rc);
            Scal sc;
            Prep P;
            double vv[4];
            scal_dot(o, sc);
            sc.pp = sum_over_r(valid ? sc.pp : 0.0);
            scal_finish(o, sc, P, vv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                P.bv[i] = valid ? P.bv[i] : 0.0;
                vv[i] = valid ? vv[i] : 0.0;
            }
            if (!valid) P.a4 = make_float4(0.f, 0.f, 0.f, 0.f);
            mma_all(P);
            gamma_add(valid ? o.k : -1, vv);
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

#ifndef SMCB_SE_MINB
#define SMCB_SE_MINB 3          // CTAs per SM the register allocation leaves room for (168 registers)
#endif
#ifndef SMCB_SE_DEPTH
#define SMCB_SE_DEPTH 4         // groups in flight per warp
#endif
constexpr int kSERingDepth = SMCB_SE_DEPTH;
constexpr int kSEStageU = 0, kSEStageW = 1024, kSEStagePriv = 2048, kSEStageBytes = 2048 + 32 * 4;
constexpr size_t kSERingBytes = (size_t)kSEWarps * kSERingDepth * kSEStageBytes;

__global__ void __launch_bounds__(kSEWarps * 32, SMCB_SE_MINB) k_stats32e(Model m, Plan p, Work w)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *tiles = reinterpret_cast<double *>(smem_raw);                 // [kSEWarps][32*33]
    double *dred = tiles + (size_t)kSEWarps * 32 * 33;                    // [kSEWarps][32]
    unsigned char *ring = reinterpret_cast<unsigned char *>(dred + (size_t)kSEWarps * 32);   // [kSEWarps][kSERingDepth][kSEStageBytes]
    const int item = blockIdx.x;
    const int t = p.it_contig[item], e = p.it_eig[item], n = p.it_len[item];
    const int64_t g0 = p.blk_off[t];
    const int2 *rec = p.erec + p.it_start[item];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(kFullMask, tid >> 5, 0);   // broadcast: warp-uniform loop bounds for the compiler
    const int r = lane >> 2, q = lane & 3;
    double *tile = tiles + (size_t)warp * 32 * 33;
    for (int x = lane; x < 32 * 33; x += 32) tile[x] = 0.0;
    __syncwarp();

    const int ngrp = (n + 3) >> 2;
    const int gbeg = (int)((long)ngrp * warp / kSEWarps), gend = (int)((long)ngrp * (warp + 1) / kSEWarps);
    const double sc = m.scale[e];
    const double *pwtab_e = m.pwtab + (size_t)e * m.n_span * 32;
    const double *pwbase = pwtab_e + 4 * r;

    double G[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) G[i][j][0] = G[i][j][1] = 0.0;
    double dacc[4] = {0.0, 0.0, 0.0, 0.0};
    int mode = 0, run_sid = -1;               // 0: G empty, 1: uniform run of span id run_sid, 2: rank-2 (already weighted)

    // G -> this warp's R_e tile; tile (mt, nt) of the C fragment: row a = 4r + mt, columns b = 4 (2q + h) + nt.  The run's
    // d~^span row comes from the table (L1) -- once per run, no registers held for it.
    auto fold = [&]() {
        if (mode == 0) return;
        dbl4 pa, pc[2];
        if (mode == 1) {
            const double *row = pwtab_e + (size_t)run_sid * 32;
            pa = ld4d(row + 4 * r);
            pc[0] = ld4d(row + 4 * (2 * q));
            pc[1] = ld4d(row + 4 * (2 * q + 1));
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double wgt = mode == 1 ? pa.v[mt] - pc[h].v[nt] : 1.0;
                    double *dst = tile + (4 * r + mt) * 33 + 4 * (2 * q + h) + nt;
                    *dst = fma(G[mt][nt][h], wgt, *dst);
                    G[mt][nt][h] = 0.0;
                }
        mode = 0;
    };

    // Lane (r, q): block q of the group, rows 4r..4r+3.  Operand ring as in k_stats32 (u and w rows through cp.async, the
    // span id next to them); the d~^span row and the span of a group are fetched from their tables (L1) one group ahead.
    struct Raw { dbl4 u, w; int sid; };
    const int nfull = n >> 2;                                         // groups [0, nfull) have four valid entries
    const int gfull = gend < nfull ? gend : nfull;
    auto load_rec = [&](int g) { return g < gfull ? __ldg(rec + 4 * g + q) : make_int2(0, 0); };
    const uint32_t ring0 = smem_addr(ring) + (uint32_t)warp * kSERingDepth * kSEStageBytes;
    const int fq = (q & 1) | ((q & 2) << 1);
    const uint32_t off_u0 = kSEStageU + q * 256 + (((2 * r) ^ fq) << 4), off_u1 = kSEStageU + q * 256 + (((2 * r + 1) ^ fq) << 4);
    const uint32_t off_w0 = off_u0 + kSEStageW, off_w1 = off_u1 + kSEStageW, off_p = kSEStagePriv + lane * 4;
    auto issue = [&](int2 rc, int stage, bool live) {
        if (live) {
            const uint32_t sb = ring0 + (uint32_t)stage * kSEStageBytes;
            const size_t gb = (size_t)(g0 + rc.x) * 32 + 4 * r;
            cp_async16(sb + off_u0, w.uvec + gb);
            cp_async16(sb + off_u1, w.uvec + gb + 2);
            cp_async16(sb + off_w0, w.bvec + gb);
            cp_async16(sb + off_w1, w.bvec + gb + 2);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(sb + off_p), "r"(rc.y) : "memory");
        }
        cp_async_commit();
    };
    auto read_stage = [&](int stage) {
        Raw o;
        const uint32_t sb = ring0 + (uint32_t)stage * kSEStageBytes;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(o.u.v[0]), "=d"(o.u.v[1]) : "r"(sb + off_u0) : "memory");
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(o.u.v[2]), "=d"(o.u.v[3]) : "r"(sb + off_u1) : "memory");
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(o.w.v[0]), "=d"(o.w.v[1]) : "r"(sb + off_w0) : "memory");
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(o.w.v[2]), "=d"(o.w.v[3]) : "r"(sb + off_w1) : "memory");
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(o.sid) : "r"(sb + off_p) : "memory");
        return o;
    };
    auto read_sid = [&](int stage) {
        int sid;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(sid) : "r"(ring0 + (uint32_t)stage * kSEStageBytes + off_p) : "memory");
        return sid;
    };
    // one group: C = 1 / (scale sum_a pw_a u_a w_a), y = C w, D_e, then G += u y^T (one run) or the weighted rank-2 form
    auto process = [&](const Raw &c, const dbl4 &pw, double span, bool valid) {
        double uv[4], yv[4], dot = 0.0;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            uv[mt] = valid ? c.u.v[mt] : 0.0;
            dot = fma(pw.v[mt] * uv[mt], c.w.v[mt], dot);
        }
        dot = sum_over_r(dot);
        const double rr = rcp_newton(valid ? sc * dot : 1.0);
        const double C = valid ? rr : 0.0;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            yv[mt] = C * c.w.v[mt];
            dacc[mt] = fma(yv[mt] * uv[mt], span * pw.v[mt], dacc[mt]);   // 1 / d~_a is applied once, after the loop
        }
        const int s0 = __shfl_sync(kFullMask, c.sid, 0);         // entry 0 of a group is always valid
        const bool uni = __all_sync(kFullMask, !valid || c.sid == s0);
        if (uni) {
            if (mode != 1 || run_sid != s0) {
                fold();
                mode = 1;
                run_sid = s0;
            }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884v(G[mt][nt][0], G[mt][nt][1], uv[mt], yv[nt]);
        } else {
            if (mode != 2) { fold(); mode = 2; }
            double xv[4], zv[4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) { xv[mt] = uv[mt] * pw.v[mt]; zv[mt] = yv[mt] * pw.v[mt]; }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    dmma884v(G[mt][nt][0], G[mt][nt][1], xv[mt], yv[nt]);
                    dmma884v(G[mt][nt][0], G[mt][nt][1], -uv[mt], zv[nt]);
                }
        }
    };
    if (gbeg < gfull) {
#pragma unroll
        for (int d = 0; d < kSERingDepth; ++d) issue(load_rec(gbeg + d), d, gbeg + d < gfull);
        int2 rcN = load_rec(gbeg + kSERingDepth);
        int stage = 0;
        int sidn = read_sid(0);                                       // (st.shared at issue time: readable at once)
        dbl4 pwn = ld4d(pwbase + (size_t)sidn * 32);
        double spn = (double)__ldg(m.span_list + sidn);
        for (int g = gbeg; g < gfull; ++g) {
            cp_async_wait<kSERingDepth - 1>();
            const Raw c = read_stage(stage);
            const dbl4 pw = pwn;
            const double span = spn;
            issue(rcN, stage, g + kSERingDepth < gfull);
            rcN = load_rec(g + kSERingDepth + 1);
            stage = (stage + 1) & (kSERingDepth - 1);
            if (g + 1 < gfull) {                                      // table rows of the next group
                sidn = read_sid(stage);
                pwn = ld4d(pwbase + (size_t)sidn * 32);
                spn = (double)__ldg(m.span_list + sidn);
            }
            process(c, pw, span, true);
        }
        cp_async_wait<0>();
    }
    if (gend > nfull) {
        // the item's last, partial group (this warp owns it): masked, straight from global memory
        const bool valid = 4 * nfull + q < n;
        const int2 rc = valid ? __ldg(rec + 4 * nfull + q) : make_int2(0, 0);
        Raw c;
        c.sid = rc.y;
        c.u = ld4d(w.uvec + (size_t)(g0 + rc.x) * 32 + 4 * r);
        c.w = ld4d(w.bvec + (size_t)(g0 + rc.x) * 32 + 4 * r);
        process(c, ld4d(pwbase + (size_t)rc.y * 32), (double)__ldg(m.span_list + rc.y), valid);
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
    return kS32RingBytes + (kk * 32 + (size_t)kS32Warps * 2 * 32 + (size_t)kS32Warps * 32) * sizeof(double) +
           (size_t)kS32Warps * 2 * sizeof(int);
}

void launch_stats32(const Model &m, const Plan &p, const Work &w, cudaStream_t st, cudaStream_t st_runs)
{
    const size_t smem = stats32_smem_bytes(m);
    static std::atomic<size_t> configured[kMaxDevices];
    if (needs_smem_config(configured, smem)) cudaFuncSetAttribute(k_stats32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_stats32<<<p.n_slabs, kS32Warps * 32, smem, st>>>(m, p, w, m.K <= kS32MaxKeysSmem ? 1 : 0);
    if (p.n_items > 0) {
        const size_t smem_e = ((size_t)kSEWarps * 32 * 33 + (size_t)kSEWarps * 32) * sizeof(double) + kSERingBytes;
        static std::atomic<size_t> configured_e[kMaxDevices];
        if (needs_smem_config(configured_e, smem_e)) cudaFuncSetAttribute(k_stats32e, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e);
        k_stats32e<<<p.n_items, kSEWarps * 32, smem_e, st_runs>>>(m, p, w);
    }
}

}  // namespace smcb

// smcpp_b200 -- forward / backward recursions on the FP64 tensor path, 8 chunks per warp (M <= 128; the text
// below describes M <= 32, i.e. NS = 1; NS = Mp / 32 scales the tile counts).
//
// A GEMV of the recursion (y = W x, 32x32 double) for EIGHT independent chunks at once is one small GEMM,
// Y[8 chunks][32] = X[8][32] W^T, i.e. 32 mma.sync.m8n8k4.f64 (SASS: DMMA) per warp instead of 8 x 32 DFMA + 8 x 16
// broadcast loads: the chunk index is the m dimension, the operand matrix W sits in shared memory in
// B-fragment order (one LDS.64 per DMMA, shared by every warp of the CTA) and the 8 chunks give the
// instruction-level parallelism that one dependent chain per warp cannot (profiles/r1b: the one-chunk-per-warp
// kernels issue 10-30 % of the time).
//
// Lane = 4n + q: n = chunk slot (row of A and C fragments), q = k-slot.  A lane holds, for its chunk, the 8 states
//     st(q, idx) = 8 (idx / 2) + 2 q + (idx % 2),   idx = 0..7
// which is exactly where mma puts C[n][8 nt + 2q + h] (idx = 2 nt + h).  Feeding those registers back as
// A-fragment k-tile idx works because the k index of an MMA may be permuted freely as long as A and B agree: the
// B fragments are built (k_setup_frags) for the state order st(q, idx).  So chained GEMVs need no shuffles.
//
// The 8 chunks of a warp advance in lockstep "rounds": every round has one block type (span-1, or span>1 with
// eigen key e) -- the type of the first unfinished chunk -- and only chunks whose current block has that type
// commit.  With the alternating run/site structure of real and synthetic contigs all 8 commit every round;
// arbitrary data only loses efficiency, never correctness.
//
// Numerics: same reference semantics as recursion32.cu (float alpha_hat, fl32 step matrix, sequential axpy
// order and Eigen's sum() order for the float normaliser, 1e-10f floor); the fp64 normalisation multiplies by
// 1/sum instead of dividing and sums the 32 doubles in a different order (a 1e-16 relative effect; the
// one-chunk kernels, used for the sequential mode and for repair sweeps, keep the reference's exact order).
#include "device_utils.cuh"
#include "estep_kernels.cuh"

namespace smcb {

constexpr int kMW = 4;             // warps per CTA
constexpr unsigned kAll = 0xffffffffu;
#ifndef SMCB_STREAM_PW
#define SMCB_STREAM_PW 0
#endif
constexpr bool kStreamPw = SMCB_STREAM_PW != 0;   // d~^span rows: ld.global.nc.L1::no_allocate
// Timing ablations for tools/ablate.sh (WRONG results, never in the product build): 1 = the float step reads one row of its
// step matrix instead of 32, 2 = every DMMA of a GEMV reuses the first B fragment, 3 = the float step runs 4 rows instead of 32
#ifndef SMCB_ABLATE
#define SMCB_ABLATE 0
#endif

__device__ __forceinline__ int st_of(int q, int idx) { return 8 * (idx >> 1) + 2 * q + (idx & 1); }

// fragment sources
constexpr int kFragReg = 0, kFragShared = 1, kFragGlobal = 2;

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// y[chunk][.] = W x[chunk][.] for the 8 chunks of the warp; F = W in B-fragment order (shared or global):
// F[(kt * 4NS + nt) * 32 + lane] = W(8 nt + r, st(q, kt)), 8NS k-tiles x 4NS n-tiles
template <int NS, bool kShared>
__device__ __forceinline__ void gemv8(const double *F, const double (&v)[8 * NS], double (&y)[8 * NS], int lane)
{
    double c[4 * NS][2];
#pragma unroll
    for (int nt = 0; nt < 4 * NS; ++nt) c[nt][0] = c[nt][1] = 0.0;
#pragma unroll
    for (int kt = 0; kt < 8 * NS; ++kt)
#pragma unroll
        for (int nt = 0; nt < 4 * NS; ++nt) {
            const int fi = SMCB_ABLATE == 2 ? 0 : (kt * 4 * NS + nt);
            const double b = kShared ? F[fi * 32 + lane] : __ldg(F + fi * 32 + lane);
            dmma(c[nt][0], c[nt][1], v[kt], b);
        }
#pragma unroll
    for (int nt = 0; nt < 4 * NS; ++nt) { y[2 * nt] = c[nt][0]; y[2 * nt + 1] = c[nt][1]; }
}

// same with the B fragments of W held in registers (NS = 1, hot eigen key): no shared-memory traffic at all
template <int NF>
__device__ __forceinline__ void gemv8_reg(const double (&F)[NF], const double (&v)[8], double (&y)[8])
{
    static_assert(NF == 32 || NF == 1, "fragment array");
    double c[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) c[nt][0] = c[nt][1] = 0.0;
    if (NF == 32) {
#pragma unroll
        for (int kt = 0; kt < 8; ++kt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma(c[nt][0], c[nt][1], v[kt], F[(kt * 4 + nt) % NF]);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { y[2 * nt] = c[nt][0]; y[2 * nt + 1] = c[nt][1]; }
}

template <int NS, int FRAG, int NF>
__device__ __forceinline__ void gemv_hot(const double (&R)[NF], const double *S, const double *Gm, const double (&v)[8 * NS],
                                         double (&y)[8 * NS], int lane)
{
    if constexpr (FRAG == kFragReg) gemv8_reg(R, v, y);
    else if constexpr (FRAG == kFragShared) gemv8<NS, true>(S, v, y, lane);
    else gemv8<NS, false>(Gm, v, y, lane);
}

// ---- packed float pairs (sm_100 FFMA2): the span-1 float GEMV must round the product and the sum separately (the
// reference's Eigen GEMV is compiled without FMA), which costs an FMUL and an FADD per element.  Two exact identities
// let one packed instruction do each for TWO elements without ever fusing them:
//     fl(x a)   = fma.rn(x, a, -0.0)      (adding -0 changes neither the value nor the sign of a zero product)
//     fl(p + y) = fma.rn(p, 1.0, y)
// ptxas (12.9) contracts a mul.rn.f32x2 feeding an add.rn.f32x2 into ONE fused FFMA2 even under --fmad=false, and it
// does the same to the two fma forms above when -0.0 and 1.0 are literals (it first simplifies them to mul / add).  The
// two constants therefore arrive as kernel parameters (Model::c_negzero2 / c_one2), which ptxas cannot see through; the
// SASS then holds exactly two FFMA2 per element pair (checked with cuobjdump, and bit for bit by the parity tests).
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
// one 256-bit read-only load of 8 consecutive floats as 4 packed pairs, 32-byte aligned
__device__ __forceinline__ f32x2x4 ldg256p(const float *p)
{
    f32x2x4 r;
    asm("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]) : "l"(p));
    return r;
}

__device__ __forceinline__ void lds2p(const float *p, f32x2 &a, f32x2 &b)   // 128-bit shared load of 4 floats as 2 packed pairs
{
    asm("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"((unsigned)__cvta_generic_to_shared(p)));
}

// y[j] = fl(y[j] + fl(x_i A(i, j))) for i = 0..MP-1 in order, for this lane's 8 NS columns (packed in pairs):
// the reference's k-sequential axpy order with the float-rounded step matrix (src/hmm.cpp:85-89).
// A points at this lane's columns of row 0 (row stride 4 NI floats).  MODE says where the matrix lives:
//   kAGlobal   global memory through the read-only path (one 256-bit load per row part)
//   kAShared   shared memory (every chunk of the warp has one of the resident keys)
//   kAGeneric  per lane either of the two (generic 128-bit loads): the frequent keys' matrices are resident in shared
//              memory, the rare ones (60+ full-SFS keys, 4 KB each) come through L1 / L2
// ncu r2a: with all matrices in global memory the L1 hit rate is 54 % and 30 % of the kernel's samples wait on these loads.
constexpr int kAGlobal = 0, kAShared = 1, kAMixed = 2;
// One row part (8 floats) for a warp whose lanes are split between resident and non-resident keys: two PREDICATED loads
// into the same registers -- lanes with a resident key read shared memory (4 distinct 16-byte pieces per warp instruction:
// one LSU wavefront), the others the read-only path (their quads only) -- instead of one generic load whose slowest lane
// and widest form set the cost for everybody.
__device__ __forceinline__ void ld_row_mixed(f32x2x4 &r, const float *ps, const float *pg, bool in_smem)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %6, 0;\n"
                 " @p ld.shared.v2.b64 {%0,%1}, [%4];\n @p ld.shared.v2.b64 {%2,%3}, [%4+16];\n"
                 " @!p ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%5];\n}"
                 : "+l"(r.v[0]), "+l"(r.v[1]), "+l"(r.v[2]), "+l"(r.v[3])
                 : "r"((unsigned)__cvta_generic_to_shared(ps)), "l"(pg), "r"((int)in_smem));
}
template <int NS, int MODE>
__device__ __forceinline__ void float_gemv(const float *A, const float *Ag, bool in_smem, const float4 *xr, f32x2 (&y2)[4 * NS],
                                           f32x2 kNegZero2, f32x2 kOne2)
{
    constexpr int MP = 32 * NS, NI = 8 * NS;
#pragma unroll(NS == 1 ? 8 : 2)
    for (int i4 = 0; i4 < (SMCB_ABLATE == 3 ? 1 : MP / 4); ++i4) {
        const float4 xv = xr[i4];
#pragma unroll
        for (int cidx = 0; cidx < 4; ++cidx) {
            const float xi = cidx == 0 ? xv.x : cidx == 1 ? xv.y : cidx == 2 ? xv.z : xv.w;
            const f32x2 xx = pack2(xi, xi);
            const size_t row = SMCB_ABLATE == 1 ? 0 : (size_t)(4 * i4 + cidx) * 4 * NI;
#pragma unroll
            for (int h = 0; h < NS; ++h) {
                f32x2x4 av;
                if (MODE == kAShared) { lds2p(A + row + 8 * h, av.v[0], av.v[1]); lds2p(A + row + 8 * h + 4, av.v[2], av.v[3]); }
                else if (MODE == kAMixed) { av.v[0] = av.v[1] = av.v[2] = av.v[3] = 0ull; ld_row_mixed(av, A + row + 8 * h, Ag + row + 8 * h, in_smem); }
                else av = ldg256p(Ag + row + 8 * h);
#pragma unroll
                for (int j2 = 0; j2 < 4; ++j2)   // two columns per instruction
                    y2[4 * h + j2] = fma2(fma2(xx, av.v[j2], kNegZero2), kOne2, y2[4 * h + j2]);
            }
        }
    }
}

__device__ __forceinline__ double group_sum(double v)   // over the 4 lanes of a chunk
{
    v += __shfl_xor_sync(kAll, v, 1);
    v += __shfl_xor_sync(kAll, v, 2);
    return v;
}

// ---- fragment tables (built once per E-step) ------------------------------------------------------------
__global__ void k_setup_frags(Model m)
{
    const int NE = m.n_eig, K = m.K, Mp = m.Mp, NS = Mp / 32, MM = Mp * Mp, NT = 4 * NS, NI = 8 * NS;
    const long tid = blockIdx.x * (long)blockDim.x + threadIdx.x, nth = (long)gridDim.x * blockDim.x;
    double *F_Td = const_cast<double *>(m.F_Td), *F_P = const_cast<double *>(m.F_P), *F_PT = const_cast<double *>(m.F_PT),
           *F_Pinv = const_cast<double *>(m.F_Pinv), *F_PinvT = const_cast<double *>(m.F_PinvT);
    // F_W[(kt*NT + nt)*32 + lane] = W(8 nt + r, st(q, kt)),  lane = 4 r + q
    for (long x = tid; x < (long)(1 + NE) * MM; x += nth) {
        const int mat = (int)(x / MM), y = (int)(x % MM);
        const int lane = y & 31, f = y >> 5, nt = f % NT, kt = f / NT, r = lane >> 2, q = lane & 3;
        const int j = 8 * nt + r, i = st_of(q, kt);
        if (mat == 0) {
            F_Td[y] = m.Td[(size_t)j * Mp + i];
        } else {
            const int e = mat - 1;
            const double *P = m.P + (size_t)e * MM, *Pinv = m.Pinv + (size_t)e * MM;
            F_P[(size_t)e * MM + y] = P[(size_t)j * Mp + i];
            F_PT[(size_t)e * MM + y] = P[(size_t)i * Mp + j];
            F_Pinv[(size_t)e * MM + y] = Pinv[(size_t)j * Mp + i];
            F_PinvT[(size_t)e * MM + y] = Pinv[(size_t)i * Mp + j];
        }
    }
    // q-major permuted emission vectors: Eq[k][q*NI + idx] = E[k][st(q, idx)]
    double *Eq = const_cast<double *>(m.Eq);
    for (long x = tid; x < (long)K * Mp; x += nth) {
        const int k = (int)(x / Mp), y = (int)(x % Mp);
        Eq[x] = m.E[(size_t)k * Mp + st_of(y / NI, y % NI)];
    }
    // step matrices: A32q[((k*Mp + i)*4 + q)*NI + idx] = fl32(e_k(j) Td(i,j)), j = st(q, idx)
    float *A32q = const_cast<float *>(m.A32q);
    for (long x = tid; x < (long)K * MM; x += nth) {
        const int idx = (int)(x % NI);
        long rest = x / NI;
        const int q = (int)(rest & 3);
        rest >>= 2;
        const int i = (int)(rest % Mp), k = (int)(rest / Mp);
        A32q[x] = m.A32[(size_t)k * MM + (size_t)i * Mp + st_of(q, idx)];
    }
}

void launch_setup_frags(const Model &m, cudaStream_t st)
{
    long work = (long)m.K * m.Mp * m.Mp;
    int blocks = (int)((work + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    k_setup_frags<<<blocks, 256, 0, st>>>(m);
}

// Predicated loads that write their destination IN PLACE (it keeps its value when the predicate is false).  A plain
// `if (p) x = load` on a loop-carried register compiles to a load into a scratch register plus a move, and the move
// waits for the load right where it was issued (ncu r1h: 4-5 % of each recursion kernel at the batch rotation).
__device__ __forceinline__ void ldg_if(int &dst, const int32_t *ptr, bool pred)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p ld.global.nc.b32 %0, [%1];\n}" : "+r"(dst) : "l"(ptr), "r"((int)pred));
}
__device__ __forceinline__ void ldg_if(int &dst, const kcode_t *ptr, bool pred)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p ld.global.nc.b32 %0, [%1];\n}" : "+r"(dst) : "l"(ptr), "r"((int)pred));
}
__device__ __forceinline__ double2 ldg_stream2(const double2 *p)   // read-only, no L1 allocation (single-use table rows)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void prefetch_l1(const void *ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

// ---- shared per-chunk bookkeeping -------------------------------------------------------------------------
struct ObsBatch {          // (span, span id, code) of 8 consecutive blocks of the lane's chunk: lane q holds blocks q and 4 + q
    int sp_lo, sp_hi, kc_lo, kc_hi, id_lo, id_hi;
};

// =============================================== forward ===================================================
// FRAG: where the B fragments of the hot eigen key live -- registers (NS = 1 only: 2 x 64 registers, at most 8 warps/SM,
// forward and backward then run one after the other), shared memory (all warps of both kernels resident at once) or
// global memory through the read-only path (M = 128: a fragment table is 128 KB).  The launcher picks.
template <int NS, int FRAG>
__device__ __forceinline__ void forward_mma_body(const Model &m, const Plan &p, const Work &w, const int Gw, const int nkc, const int bid)
{
    constexpr int MP = 32 * NS, NI = 8 * NS, NT = 4 * NS, MM = MP * MP, XS = MP + 4;
    constexpr int NF = FRAG == kFragReg ? 32 : 1;
    const int G = Gw & 255, wpc = (Gw >> 8) ? (Gw >> 8) : kMW;       // chunks per warp, warps of a CTA that carry chunks (pack_gw)
    static_assert(FRAG != kFragReg || NS == 1, "register fragments need M <= 32");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sF_Pinv = reinterpret_cast<double *>(smem_raw);       // [MM] + [MM], used when FRAG == shared
    double *sF_P = sF_Pinv + MM;
    float *s_x = reinterpret_cast<float *>(smem_raw + (FRAG == kFragShared ? 2 * MM * sizeof(double) : 0));   // [kMW][8][XS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = lane >> 2, q = lane & 3;
    const int hot = m.hot_eig;
    double rF_Pinv[NF], rF_P[NF];
    if constexpr (FRAG == kFragReg) {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            rF_Pinv[f] = hot >= 0 ? m.F_Pinv[(size_t)hot * MM + f * 32 + lane] : 0.0;
            rF_P[f] = hot >= 0 ? m.F_P[(size_t)hot * MM + f * 32 + lane] : 0.0;
        }
    } else {
        rF_Pinv[0] = rF_P[0] = 0.0;
        if constexpr (FRAG == kFragShared) {
            if (hot >= 0)
                for (int x = tid; x < MM; x += kMW * 32) {
                    sF_Pinv[x] = m.F_Pinv[(size_t)hot * MM + x];
                    sF_P[x] = m.F_P[(size_t)hot * MM + x];
                }
            __syncthreads();
        }
    }
    float *xs = s_x + ((size_t)warp * 8 + n) * XS;   // this chunk's row (stride XS floats: fewer bank conflicts)
    // float step matrices of the nkc most frequent span-1 keys, [nkc][MM] in A32q order
    float *s_A = s_x + (size_t)kMW * 8 * XS;
    for (int sl = 0; sl < nkc; ++sl) {
        const float4 *src = reinterpret_cast<const float4 *>(m.A32q + (size_t)m.hot_keys[sl] * MM);
        float4 *dst = reinterpret_cast<float4 *>(s_A + (size_t)sl * MM);
        for (int x4 = tid; x4 < MM / 4; x4 += kMW * 32) dst[x4] = __ldg(src + x4);
    }
    if (nkc > 0) __syncthreads();
    const int M = m.M;

    // G (<= 8) chunks per warp: small inputs spread over more warps (rows n >= G of the MMA stay idle)
    const int c = (n < G && warp < wpc) ? (bid * wpc + warp) * G + n : p.n_chunks;
    bool active = c < p.n_chunks;
    const int cc = active ? c : 0;
    const int t = p.ch_contig[cc], s = p.ch_start[cc], len = p.ch_len[cc];
    const int64_t g0 = p.blk_off[t];
    const int bend = s + len;
    float *acol = w.alpha + (p.col_off[t] + (int64_t)(cc - p.chunk_off[t]) * (p.chunk_blocks + 1)) * MP;
    int cur = s - p.burn_in_fwd;
    if (cur < 0) cur = 0;

    float x[NI];
#pragma unroll
    for (int idx = 0; idx < NI; ++idx) x[idx] = (float)m.pi[st_of(q, idx)];   // reference src/hmm.cpp:59 (pads are 0)
    auto store_col = [&](float *dst) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<float2 *>(dst + 8 * nt + 2 * q) = make_float2(x[2 * nt], x[2 * nt + 1]);
    };
    if (active && cur == s) { store_col(acol); store_col(w.start_used + (size_t)c * MP); }

    int base = cur;
    ObsBatch ob, obn;                         // current batch of 8 blocks and the one after it (fetched a batch ahead)
    // Loads are unconditional with a clamped index: a guarded load with a default value costs a select right behind the
    // load, which parks the warp on the scoreboard where the batch is fetched (ncu r1g: 3-5 % of the kernel).  Entries
    // past the chunk's end are never used (the chunk goes inactive first); idle lanes read chunk 0's rows.
    auto load_batch = [&](int b) {
        ObsBatch o;
        const int64_t i0 = g0 + min(b + q, bend - 1), i1 = g0 + min(b + 4 + q, bend - 1);
        o.sp_lo = p.span[i0]; o.kc_lo = p.kcode[i0]; o.id_lo = p.span_id[i0];
        o.sp_hi = p.span[i1]; o.kc_hi = p.kcode[i1]; o.id_hi = p.span_id[i1];
        return o;
    };
    ob = load_batch(base);
    obn = load_batch(base + 8);
    double llsum = 0.0, lprod = 1.0;
    int lcnt = 0, done = 0, rounds = 0;
    // (span, code, span id) of the chunk's current block and, for a span>1 block, its d~^span vector: fetched at the END
    // of the previous round, so the table row's L2 latency hides behind the round bookkeeping and the first GEMV
    // (ptxas sinks a load issued inside the round down to its first use).
    int span, kc, sid;
    auto fetch_cur = [&]() {
        const int pos = cur - base;
        const int src = (lane & ~3) | (pos & 3);
        span = __shfl_sync(kAll, (pos & 4) ? ob.sp_hi : ob.sp_lo, src);
        kc = __shfl_sync(kAll, (pos & 4) ? ob.kc_hi : ob.kc_lo, src);
        sid = __shfl_sync(kAll, (pos & 4) ? ob.id_hi : ob.id_lo, src);
    };
    double2 pwv[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) pwv[nt] = make_double2(0.0, 0.0);
    auto load_pw = [&]() {   // states st(q, 2nt), st(q, 2nt + 1) of the q-major table
        const double2 *pw = reinterpret_cast<const double2 *>(m.pwq + ((size_t)((kc >> kKeyBits) - 1) * m.n_span + sid) * MP + q * NI);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) pwv[nt] = kStreamPw ? ldg_stream2(pw + nt) : __ldg(pw + nt);
    };
    fetch_cur();
    if (active && (kc >> kKeyBits) > 0) load_pw();

    for (;;) {
        const unsigned am = __ballot_sync(kAll, active);
        if (!am) break;
        ++rounds;
        const int type = active ? (kc >> kKeyBits) : -1;
        // the least advanced chunk picks the round's block type (no chunk can starve, phases re-align by themselves)
        const unsigned lead = __reduce_min_sync(kAll, active ? (((unsigned)done << 5) | (unsigned)lane) : 0xffffffffu);
        const int T = __shfl_sync(kAll, type, lead & 31);
        const bool adv = active && type == T;
        const int k = adv ? (kc & kKeyMask) : (nkc > 0 ? m.hot_keys[0] : 0);
        float xn[NI];
        double cmul = 1.0, cadd = 0.0;
        float sf = 0.f;
        if (T > 0) {
            // a = P_r (d~^span o (Pinv_r alpha_prev)); reference src/hmm.cpp:74-80
            const int e = T - 1;
            double xd[NI], u[NI], a[NI];
#pragma unroll
            for (int idx = 0; idx < NI; ++idx) xd[idx] = (double)x[idx];
            if (e == hot) gemv_hot<NS, FRAG>(rF_Pinv, sF_Pinv, m.F_Pinv + (size_t)e * MM, xd, u, lane);
            else gemv8<NS, false>(m.F_Pinv + (size_t)e * MM, xd, u, lane);
            const int sp = adv ? span : 1;
            {
                // u_l = Pinv_r alpha_hat_{l-1} is an operand of the statistics pass (stats32.cu), stored in eigen-index order
                if (adv && cur >= s) {
                    double *ud = w.uvec + (size_t)(g0 + cur) * MP + 2 * q;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<double2 *>(ud + 8 * nt) = make_double2(u[2 * nt], u[2 * nt + 1]);
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) { u[2 * nt] *= pwv[nt].x; u[2 * nt + 1] *= pwv[nt].y; }
            if (e == hot) gemv_hot<NS, FRAG>(rF_P, sF_P, m.F_P + (size_t)e * MM, u, a, lane);
            else gemv8<NS, false>(m.F_P + (size_t)e * MM, u, a, lane);
            double part = 0.0;
#pragma unroll
            for (int idx = 0; idx < NI; ++idx) part += a[idx];
            const double ssum = group_sum(part);
            const double rs = 1.0 / ssum;
#pragma unroll
            for (int idx = 0; idx < NI; ++idx) xn[idx] = (float)(a[idx] * rs);
            cmul = ssum;
            cadd = (double)sp * m.logscale[e];
        } else {
            // float GEMV, k-sequential axpy order with the float-rounded matrix; reference src/hmm.cpp:85-89
            __syncwarp();
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<float2 *>(xs + 8 * nt + 2 * q) = make_float2(x[2 * nt], x[2 * nt + 1]);
            __syncwarp();
            const float4 *xr = reinterpret_cast<const float4 *>(xs);
            f32x2 y2[NI / 2];
#pragma unroll
            for (int idx = 0; idx < NI / 2; ++idx) y2[idx] = 0ull;
            int slot = -1;
            for (int sl = 0; sl < nkc; ++sl)
                if (k == m.hot_keys[sl]) slot = sl;
            // row i: + i * 4 * NI floats; this lane's NI columns are contiguous
            const float *Ag = m.A32q + ((size_t)k * MP * 4 + q) * NI;
            const float *As = s_A + (size_t)(slot >= 0 ? slot : 0) * MM + q * NI;
            if (nkc == 0) {
                float_gemv<NS, kAGlobal>(As, Ag, false, xr, y2, m.c_negzero2, m.c_one2);
            } else if (__all_sync(kAll, slot >= 0)) {
                float_gemv<NS, kAShared>(As, Ag, true, xr, y2, m.c_negzero2, m.c_one2);
            } else {
                float_gemv<NS, kAMixed>(As, Ag, slot >= 0, xr, y2, m.c_negzero2, m.c_one2);
            }
            float y[NI];
#pragma unroll
            for (int idx = 0; idx < NI / 2; ++idx) unpack2(y2[idx], y[2 * idx], y[2 * idx + 1]);
            if (M == MP) {
                // Eigen's sum() order for 32 NS aligned floats (device_utils.cuh: eigen_sum_f32; packets p0 / p1 accumulate the
                // coefficients = c and = 4 + c (mod 8), then p0 + p1, then (x + z) + (y + w)) without a trip through shared
                // memory: lane q holds exactly the coefficients = 2q, 2q + 1 (mod 8), in ascending order, so its two
                // chains are p0.x/p0.y (q = 0), p0.z/p0.w (q = 1), p1.x/p1.y (q = 2), p1.z/p1.w (q = 3); the rest is
                // two butterfly steps (float addition commutes exactly).
                float ev = y[0], od = y[1];
#pragma unroll
                for (int gq = 1; gq < NI / 2; ++gq) { ev = __fadd_rn(ev, y[2 * gq]); od = __fadd_rn(od, y[2 * gq + 1]); }
                ev = __fadd_rn(ev, __shfl_xor_sync(kAll, ev, 2));     // q in {0, 2}: t.x, others: t.z
                od = __fadd_rn(od, __shfl_xor_sync(kAll, od, 2));     //              t.y          t.w
                ev = __fadd_rn(ev, __shfl_xor_sync(kAll, ev, 1));     // t.x + t.z
                od = __fadd_rn(od, __shfl_xor_sync(kAll, od, 1));     // t.y + t.w
                sf = __fadd_rn(ev, od);
            } else {
                __syncwarp();
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<float2 *>(xs + 8 * nt + 2 * q) = make_float2(y[2 * nt], y[2 * nt + 1]);
                __syncwarp();
                sf = eigen_sum_f32(xs, M, (M & 3) ? (int)((4 - (((long)(cur + 1) * M) & 3)) & 3) : 0);
            }
#pragma unroll
            for (int idx = 0; idx < NI; ++idx) xn[idx] = __fdiv_rn(y[idx], sf);
            cmul = (double)sf;
        }
        if (adv) {
#pragma unroll
            for (int idx = 0; idx < NI; ++idx) {
                float v = xn[idx];
                if (st_of(q, idx) < M && v < 1e-10f) v = 1e-10f;     // reference src/hmm.cpp:92-94
                x[idx] = v;
            }
            if (cur >= s) {
                store_col(acol + (size_t)(cur - s + 1) * MP);
                lprod *= cmul;
                llsum += cadd;
                if (++lcnt == 8 || !(lprod > 1e-200)) { llsum += log(lprod); lprod = 1.0; lcnt = 0; }
                if (T == 0 && q == 0) w.cnorm[g0 + cur] = sf;
            } else if (cur == s - 1) {
                store_col(acol);
                store_col(w.start_used + (size_t)c * MP);
            }
            ++cur;
            ++done;
            if (cur >= bend) active = false;
        }
        {
            const bool rot = adv && active && cur - base == 8;
            if (rot) { base += 8; ob = obn; }
            const int64_t i0 = g0 + min(base + 8 + q, bend - 1), i1 = g0 + min(base + 12 + q, bend - 1);
            ldg_if(obn.sp_lo, p.span + i0, rot); ldg_if(obn.kc_lo, p.kcode + i0, rot); ldg_if(obn.id_lo, p.span_id + i0, rot);
            ldg_if(obn.sp_hi, p.span + i1, rot); ldg_if(obn.kc_hi, p.kcode + i1, rot); ldg_if(obn.id_hi, p.span_id + i1, rot);
        }
        fetch_cur();
        if (adv && active && (kc >> kKeyBits) > 0) load_pw();
        if constexpr (NS == 1) {
            // float step matrix of the block AFTER the next one -> L1 (one full round ahead).  The matrices of the rare keys
            // (60+ full-SFS keys, 4 KB each) do not stay in L1; without this every 4-row group of their GEMV waits on L2.
            const int pos2 = cur + 1 - base;
            const int v2 = pos2 < 8 ? ((pos2 & 4) ? ob.kc_hi : ob.kc_lo) : obn.kc_lo;
            const int kc2 = __shfl_sync(kAll, v2, (lane & ~3) | (pos2 & 3));
            bool resident = (kc2 & kKeyMask) == m.hot_keys[0];
            for (int sl = 1; sl < nkc; ++sl) resident = resident || (kc2 & kKeyMask) == m.hot_keys[sl];
            if (adv && active && cur + 1 < bend && (kc2 >> kKeyBits) == 0 && !resident) {
                const float *row = m.A32q + ((size_t)(kc2 & kKeyMask) * MP + 8 * q) * 4 * NI;   // rows 8q .. 8q + 7, 128 B each
#pragma unroll
                for (int j = 0; j < 8; ++j) prefetch_l1(row + (size_t)j * 4 * NI);
            }
        }
    }
    if (c < p.n_chunks) {
        store_col(w.end_alpha + (size_t)c * MP);
        if (q == 0) w.ll_chunk[c] = llsum + log(lprod);
    }
    if (lane == 0) atomicAdd(&w.counters[4], rounds);                 // diagnostics: lockstep efficiency
    if (q == 0 && c < p.n_chunks) atomicAdd(&w.counters[5], done);
}

// =============================================== backward ==================================================
template <int NS, int FRAG>
__device__ __forceinline__ void backward_mma_body(const Model &m, const Plan &p, const Work &w, const int Gw, const int bid)
{
    constexpr int MP = 32 * NS, NI = 8 * NS, NT = 4 * NS, MM = MP * MP;
    constexpr int NF = FRAG == kFragReg ? 32 : 1;
    const int G = Gw & 255, wpc = (Gw >> 8) ? (Gw >> 8) : kMW;
    static_assert(FRAG != kFragReg || NS == 1, "register fragments need M <= 32");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // Td fragments (span-1 rounds) are shared-resident unless everything is read from global memory (M = 128)
    double *sF_Td = reinterpret_cast<double *>(smem_raw);        // [MM]
    double *sF_PT = sF_Td + MM, *sF_PinvT = sF_PT + MM;          // [MM] each, used when FRAG == shared
    const int tid = threadIdx.x, lane = tid & 31;
    const int n = lane >> 2, q = lane & 3;
    const int hot = m.hot_eig;
    double rF_PT[NF], rF_PinvT[NF];
    if constexpr (FRAG != kFragGlobal) {
        for (int x = tid; x < MM; x += kMW * 32) sF_Td[x] = m.F_Td[x];
    }
    if constexpr (FRAG == kFragReg) {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            rF_PT[f] = hot >= 0 ? m.F_PT[(size_t)hot * MM + f * 32 + lane] : 0.0;
            rF_PinvT[f] = hot >= 0 ? m.F_PinvT[(size_t)hot * MM + f * 32 + lane] : 0.0;
        }
    } else {
        rF_PT[0] = rF_PinvT[0] = 0.0;
        if constexpr (FRAG == kFragShared) {
            if (hot >= 0)
                for (int x = tid; x < MM; x += kMW * 32) { sF_PT[x] = m.F_PT[(size_t)hot * MM + x]; sF_PinvT[x] = m.F_PinvT[(size_t)hot * MM + x]; }
        }
    }
    __syncthreads();
    const int M = m.M;

    const int c = (n < G && (tid >> 5) < wpc) ? (bid * wpc + (tid >> 5)) * G + n : p.n_chunks;
    bool active = c < p.n_chunks;
    const int cc = active ? c : 0;
    const int t = p.ch_contig[cc], s = p.ch_start[cc], len = p.ch_len[cc];
    const int64_t g0 = p.blk_off[t];
    const int L = (int)(p.blk_off[t + 1] - g0);
    const int bend = s + len;
    int b1 = bend + p.burn_in;
    if (b1 > L || bend == L) b1 = L;
    int cur = b1 - 1;                         // block processed next (descending)

    double beta[NI];
#pragma unroll
    for (int idx = 0; idx < NI; ++idx) beta[idx] = st_of(q, idx) < M ? 1.0 : 0.0;   // reference src/hmm.cpp:97
    auto store_vec = [&](double *dst, const double (&v)[NI], double mul) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<double2 *>(dst + 8 * nt + 2 * q) = make_double2(v[2 * nt] * mul, v[2 * nt + 1] * mul);
    };

    int top = cur;                            // batch = blocks top, top-1, ..., top-7; lane q holds top-q and top-4-q
    ObsBatch ob, obn;
    auto load_batch = [&](int tp) {   // unconditional, clamped (see the forward kernel)
        ObsBatch o;
        const int64_t i0 = g0 + max(tp - q, s), i1 = g0 + max(tp - 4 - q, s);
        o.sp_lo = o.sp_hi = 1;
        o.kc_lo = p.kcode[i0]; o.id_lo = p.span_id[i0];
        o.kc_hi = p.kcode[i1]; o.id_hi = p.span_id[i1];
        return o;
    };
    ob = load_batch(top);
    obn = load_batch(top - 8);
    int since = 0, done = 0;
    // code and span id of the chunk's current block and the vector that multiplies inside its step -- d~^span (span > 1)
    // or e_k (span 1), both q-major -- fetched at the END of the previous round (see the forward kernel)
    int kc, sid;
    auto fetch_cur = [&]() {
        const int pos = top - cur;
        const int src = (lane & ~3) | (pos & 3);
        kc = __shfl_sync(kAll, (pos & 4) ? ob.kc_hi : ob.kc_lo, src);
        sid = __shfl_sync(kAll, (pos & 4) ? ob.id_hi : ob.id_lo, src);
    };
    double2 opv[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) opv[nt] = make_double2(0.0, 0.0);
    auto load_op = [&]() {
        const int ty = kc >> kKeyBits;
        const double *row = ty > 0 ? m.pwq + ((size_t)(ty - 1) * m.n_span + sid) * MP : m.Eq + (size_t)(kc & kKeyMask) * MP;
        const double2 *src = reinterpret_cast<const double2 *>(row + q * NI);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) opv[nt] = (kStreamPw && ty > 0) ? ldg_stream2(src + nt) : __ldg(src + nt);
    };
    fetch_cur();
    if (active) load_op();

    for (;;) {
        const unsigned am = __ballot_sync(kAll, active);
        if (!am) break;
        const int type = active ? (kc >> kKeyBits) : -1;
        const unsigned lead = __reduce_min_sync(kAll, active ? (((unsigned)done << 5) | (unsigned)lane) : 0xffffffffu);
        const int T = __shfl_sync(kAll, type, lead & 31);
        const bool adv = active && type == T;
        // the chunk's verified start value: normalised, recorded before the first stored step
        const bool rec = adv && cur == bend - 1;
        if (__any_sync(kAll, rec)) {
            double part = 0.0;
#pragma unroll
            for (int idx = 0; idx < NI; ++idx) part += beta[idx];
            const double bs = group_sum(part);
            if (rec) {
                const double rb = 1.0 / bs;
#pragma unroll
                for (int idx = 0; idx < NI; ++idx) beta[idx] *= rb;
                store_vec(w.bstart_used + (size_t)c * MP, beta, 1.0);
            }
        }
        const bool storing = adv && cur < bend;
        double *bv = w.bvec + (size_t)(g0 + (adv ? cur : 0)) * MP;
        double nb[NI];
        if (T > 0) {
            // beta <- Pinv_r^T (d~^span o (P_r^T beta)); reference src/hmm.cpp:123-127
            const int e = T - 1;
            double wv[NI];
            if (e == hot) gemv_hot<NS, FRAG>(rF_PT, sF_PT, m.F_PT + (size_t)e * MM, beta, wv, lane);
            else gemv8<NS, false>(m.F_PT + (size_t)e * MM, beta, wv, lane);
            if (storing) store_vec(bv, wv, 1.0);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) { wv[2 * nt] *= opv[nt].x; wv[2 * nt + 1] *= opv[nt].y; }
            if (e == hot) gemv_hot<NS, FRAG>(rF_PinvT, sF_PinvT, m.F_PinvT + (size_t)e * MM, wv, nb, lane);
            else gemv8<NS, false>(m.F_PinvT + (size_t)e * MM, wv, nb, lane);
        } else {
            // beta <- Td (e_k o beta); reference src/hmm.cpp:139
            if (storing) store_vec(bv, beta, 1.0);
            double tv[NI];
#pragma unroll
            for (int h = 0; h < NI / 2; ++h) {
                tv[2 * h] = opv[h].x * beta[2 * h];
                tv[2 * h + 1] = opv[h].y * beta[2 * h + 1];
            }
            if constexpr (FRAG == kFragGlobal) gemv8<NS, false>(m.F_Td, tv, nb, lane);
            else gemv8<NS, true>(sF_Td, tv, nb, lane);
        }
        // loose normalisation by an exact power of two (every statistic is invariant to beta's scale)
        ++since;
        // "the vector has become tiny": exponent of its largest entry, integer compares on the high words (eight DSETP per
        // step on the FP64 pipe were 8 % of this kernel's samples, ncu r2a)
        int hi = 0;
#pragma unroll
        for (int idx = 0; idx < NI; ++idx) hi = max(hi, __double2hiint(nb[idx]) & 0x7fffffff);
        hi = max(hi, __shfl_xor_sync(kAll, hi, 1));
        hi = max(hi, __shfl_xor_sync(kAll, hi, 2));
        const bool need = adv && (since >= 4 || hi < 0x2b200000);      // largest |entry| < 2^-333 ~ 1e-100
        if (__any_sync(kAll, need)) {
            double part = 0.0;
#pragma unroll
            for (int idx = 0; idx < NI; ++idx) part += nb[idx];
            const double f = pow2_rescale(group_sum(part));
            if (need) {
#pragma unroll
                for (int idx = 0; idx < NI; ++idx) nb[idx] *= f;
                since = 0;
            }
        }
        if (adv) {
#pragma unroll
            for (int idx = 0; idx < NI; ++idx) beta[idx] = nb[idx];
            --cur;
            ++done;
            if (cur < s) active = false;
        } else {
            --since;
        }
        {
            const bool rot = adv && active && top - cur == 8;
            if (rot) { top -= 8; ob = obn; }
            const int64_t i0 = g0 + max(top - 8 - q, s), i1 = g0 + max(top - 12 - q, s);
            ldg_if(obn.kc_lo, p.kcode + i0, rot); ldg_if(obn.id_lo, p.span_id + i0, rot);
            ldg_if(obn.kc_hi, p.kcode + i1, rot); ldg_if(obn.id_hi, p.span_id + i1, rot);
        }
        fetch_cur();
        if (adv && active) load_op();
    }
    if (c < p.n_chunks) {
        double part = 0.0;
#pragma unroll
        for (int idx = 0; idx < NI; ++idx) part += beta[idx];
        const double bs = group_sum(part);
        store_vec(w.beta_out + (size_t)c * MP, beta, 1.0 / bs);
    } else {
        group_sum(0.0);
    }
}

template <int NS, int FRAG>
__global__ void __launch_bounds__(kMW * 32) k_forward_mma(Model m, Plan p, Work w, int G, int nkc)
{
    forward_mma_body<NS, FRAG>(m, p, w, G, nkc, blockIdx.x);
}

template <int NS, int FRAG>
__global__ void __launch_bounds__(kMW * 32) k_backward_mma(Model m, Plan p, Work w, int G)
{
    backward_mma_body<NS, FRAG>(m, p, w, G, blockIdx.x);
}

// Both recursions in ONE launch: the forward and the backward pass are independent (the statistics need both), and they
// complement each other on an SM -- the forward pass spends half of its rounds on the FP32 pipe (float step), the backward
// pass lives on the FP64 tensor pipe.  A CTA is either a forward or a backward CTA: the first `blocks` CTAs are the
// backward pass, the rest the forward pass.  The hardware hands CTAs to the SMs in block-index order, so every SM gets
// its share of both, and the forward warps -- the longer pass -- land in the higher warp slots, which the issue
// arbiter favours (interleaving the roles by block index gave the backward pass the priority: 8.2 ms instead of 7.65 ms
// for the recursion phase of the benchmark).  This replaces two streams + an event + a host synchronisation between setup
// and recursions (two kernels enqueued behind running setup kernels started one after the other instead of side by side).
template <int NS, int FRAG>
__global__ void __launch_bounds__(kMW * 32) k_recursions_mma(Model m, Plan p, Work w, int G, int nkc, int blocks)
{
    const int bid = blockIdx.x;
    if (bid < blocks) backward_mma_body<NS, FRAG>(m, p, w, G, bid);
    else forward_mma_body<NS, FRAG>(m, p, w, G, nkc, bid - blocks);
}

// ---- launch -------------------------------------------------------------------------------------------------
// RecOpts::cached_keys: number of span-1 keys whose float step matrix the forward kernel keeps in shared memory
// (M <= 32 only; option "fwd_cached_keys"), 4 KB each.  Lanes whose key is resident read shared memory, the others the
// read-only path (float_gemv<kAGeneric>).
static int cached_keys(const Model &m, const RecOpts &o)
{
    if (m.Mp != 32) return 0;
    int n = 0;
    while (n < o.cached_keys && n < 4 && m.hot_keys[n] >= 0) ++n;
    return n;
}
static size_t fwd_smem(int NS, int frag, int nkc)
{
    const size_t MP = 32 * NS;
    return (frag == kFragShared ? 2 * MP * MP * sizeof(double) : 0) + (size_t)kMW * 8 * (MP + 4) * sizeof(float) +
           (size_t)nkc * MP * MP * sizeof(float);
}
static size_t bwd_smem(int NS, int frag)
{
    const size_t MP = 32 * NS;
    return frag == kFragGlobal ? 16 : (frag == kFragShared ? 3 : 1) * MP * MP * sizeof(double);
}

int resident_warps_mma(int n_sm, int Mp, const RecOpts &o)
{
    int bf = 0, bb = 0;
    if (Mp == 32) {
        const int nkc = o.cached_keys < 0 ? 0 : (o.cached_keys > 4 ? 4 : o.cached_keys);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bf, k_forward_mma<1, kFragShared>, kMW * 32, fwd_smem(1, kFragShared, nkc));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bb, k_backward_mma<1, kFragShared>, kMW * 32, bwd_smem(1, kFragShared));
    } else if (Mp == 64) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bf, k_forward_mma<2, kFragShared>, kMW * 32, fwd_smem(2, kFragShared, 0));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bb, k_backward_mma<2, kFragShared>, kMW * 32, bwd_smem(2, kFragShared));
    } else {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bf, k_forward_mma<4, kFragGlobal>, kMW * 32, fwd_smem(4, kFragGlobal, 0));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bb, k_backward_mma<4, kFragGlobal>, kMW * 32, bwd_smem(4, kFragGlobal));
    }
    int b = bf < bb ? bf : bb;
    if (b < 1) b = 1;
    return n_sm * b * kMW;
}

// chunks per warp: 8 when there are enough chunks to give every SM `want` warps, fewer otherwise
// (RecOpts::force_G, option "chunks_per_warp", pins it: small inputs otherwise always run with G = 1)
static int chunks_per_warp(int n_chunks, int n_sm, int Mp, const RecOpts &o)
{
    if (o.force_G == 1 || o.force_G == 2 || o.force_G == 4 || o.force_G == 8) return o.force_G;
    // 128 states: a chunk step is 64 x the work of a 32-state step and the burn-in is long, so a single contig offers few
    // chunks; four of them per warp and one warp per SM beat two per warp (r2, 977 chunks: forward 47 ms with G = 4, 50 with
    // G = 8, 77 with G = 2 -- no better than the one-chunk-per-warp kernel)
    const int want = Mp == 128 ? n_sm : n_sm * 2, floor_G = Mp == 128 ? 4 : 1;
    int G = 8;
    while (G > floor_G && (n_chunks + G - 1) / G < want) G >>= 1;
    return G;
}

// Warps of a CTA that carry chunks (packed into the kernels' G argument).  Spreading few warps over more, emptier CTAs was
// measured and lost: at 128 states the B fragments come from global memory, and the warps of one CTA share them in L1
// (977 chunks, 4 per warp: 245 one-warp CTAs 71.7 ms forward, 62 full CTAs 47.2 ms).  Kept as a mechanism, always kMW.
static int warps_per_cta(int, int) { return kMW; }
static int pack_gw(int G, int wpc) { return G | (wpc << 8); }

// The tensor-path forward kernel gives each chunk only 4 lanes for the float GEMV of the span-1 step; above 32 states
// that only pays off when every warp has its full 8 chunks, otherwise the one-chunk-per-warp kernel is used
// (the backward pass has no float step and always takes the tensor path).
bool mma_forward_pays(int n_chunks, int n_sm, int Mp, const RecOpts &o)
{
    const int G = chunks_per_warp(n_chunks, n_sm, Mp, o);
    return Mp == 32 || G == 8 || (Mp == 128 && G >= 4 && n_chunks >= n_sm);
}

// register-resident fragments pay off while forward + backward (250 registers each) still fit on the GPU together
static bool use_reg_frags(int warps, int n_sm) { return warps <= n_sm * 6; }

// Percent of the unified L1 / shared-memory array requested as shared memory.  The forward pass streams its float step
// matrices through L1 (54 % hit rate), so the recursion kernels ask for no more shared memory than their resident CTAs
// need (measured on C3: 50 % -> 40 %: forward 7.59 -> 7.49 ms; a 3-contig shard 50 % -> 25 %: 1.89 -> 1.84 ms).  Both
// kernels must get the SAME carve-out, otherwise the second one waits for the SMs to drain.
static int g_carveout = -1;     // -1 = automatic; option "carveout" pins it (process-wide, tools)
void set_recursion_carveout(int pct) { g_carveout = pct < 0 ? -1 : (pct > 100 ? 100 : pct); }
static int g_carveout_now = 50;

template <typename KF>
static void set_attrs(KF kernel, size_t smem)
{
    // forward and backward kernels must be able to share an SM: the same shared-memory carve-out for every variant,
    // otherwise the second kernel waits for the SMs to drain and the two passes run back to back
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, g_carveout_now);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// blocks: CTAs per recursion kernel; smem_pair: dynamic shared memory of one forward + one backward CTA
static void configure_once(int blocks, int n_sm, size_t smem_pair)
{
    int want = g_carveout;
    if (want < 0) {
        const int per_sm = (blocks + n_sm - 1) / n_sm;                       // resident CTAs of each kernel per SM
        const size_t need = (size_t)per_sm * (smem_pair + 2 * 1024) + 1024;    // + the 1 KB the driver reserves per CTA
        want = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024)) + 1;
        if (want > 100) want = 100;
    }
    static std::atomic<int> configured[kMaxDevices];
    std::atomic<int> &c = configured[current_device_slot()];
    if (c.load(std::memory_order_relaxed) == want + 1) return;
    c.store(want + 1, std::memory_order_relaxed);
    g_carveout_now = want;
    set_attrs(k_forward_mma<1, kFragReg>, fwd_smem(1, kFragReg, 4));
    set_attrs(k_forward_mma<1, kFragShared>, fwd_smem(1, kFragShared, 4));
    set_attrs(k_backward_mma<1, kFragReg>, bwd_smem(1, kFragReg));
    set_attrs(k_backward_mma<1, kFragShared>, bwd_smem(1, kFragShared));
    set_attrs(k_forward_mma<2, kFragShared>, fwd_smem(2, kFragShared, 0));
    set_attrs(k_backward_mma<2, kFragShared>, bwd_smem(2, kFragShared));
    set_attrs(k_forward_mma<4, kFragGlobal>, fwd_smem(4, kFragGlobal, 0));
    set_attrs(k_backward_mma<4, kFragGlobal>, bwd_smem(4, kFragGlobal));
    auto mx = [](size_t a, size_t b) { return a > b ? a : b; };
    set_attrs(k_recursions_mma<1, kFragReg>, mx(fwd_smem(1, kFragReg, 4), bwd_smem(1, kFragReg)));
    set_attrs(k_recursions_mma<1, kFragShared>, mx(fwd_smem(1, kFragShared, 4), bwd_smem(1, kFragShared)));
    set_attrs(k_recursions_mma<2, kFragShared>, mx(fwd_smem(2, kFragShared, 0), bwd_smem(2, kFragShared)));
    set_attrs(k_recursions_mma<4, kFragGlobal>, mx(fwd_smem(4, kFragGlobal, 0), bwd_smem(4, kFragGlobal)));
}

static size_t smem_pair(const Model &m, const RecOpts &o, int warps, int n_sm)
{
    if (m.Mp == 32) {
        const int frag = use_reg_frags(warps, n_sm) ? kFragReg : kFragShared;
        return fwd_smem(1, frag, cached_keys(m, o)) + bwd_smem(1, frag);
    }
    if (m.Mp == 64) return fwd_smem(2, kFragShared, 0) + bwd_smem(2, kFragShared);
    return fwd_smem(4, kFragGlobal, 0) + bwd_smem(4, kFragGlobal);
}

void launch_forward_mma(const Model &m, const Plan &p, const Work &w, int n_sm, const RecOpts &o, cudaStream_t st)
{
    const int G = chunks_per_warp(p.n_chunks, n_sm, m.Mp, o);
    const int warps = (p.n_chunks + G - 1) / G, wpc = warps_per_cta(warps, n_sm), blocks = (warps + wpc - 1) / wpc, Gw = pack_gw(G, wpc);
    configure_once(blocks, n_sm, smem_pair(m, o, warps, n_sm));
    if (m.Mp == 32) {
        const int nkc = cached_keys(m, o);
        if (use_reg_frags(warps, n_sm)) k_forward_mma<1, kFragReg><<<blocks, kMW * 32, fwd_smem(1, kFragReg, nkc), st>>>(m, p, w, Gw, nkc);
        else k_forward_mma<1, kFragShared><<<blocks, kMW * 32, fwd_smem(1, kFragShared, nkc), st>>>(m, p, w, Gw, nkc);
    } else if (m.Mp == 64) {
        k_forward_mma<2, kFragShared><<<blocks, kMW * 32, fwd_smem(2, kFragShared, 0), st>>>(m, p, w, Gw, 0);
    } else {
        k_forward_mma<4, kFragGlobal><<<blocks, kMW * 32, fwd_smem(4, kFragGlobal, 0), st>>>(m, p, w, Gw, 0);
    }
}

void launch_backward_mma(const Model &m, const Plan &p, const Work &w, int n_sm, const RecOpts &o, cudaStream_t st)
{
    const int G = chunks_per_warp(p.n_chunks, n_sm, m.Mp, o);
    const int warps = (p.n_chunks + G - 1) / G, wpc = warps_per_cta(warps, n_sm), blocks = (warps + wpc - 1) / wpc, Gw = pack_gw(G, wpc);
    configure_once(blocks, n_sm, smem_pair(m, o, warps, n_sm));
    if (m.Mp == 32) {
        if (use_reg_frags(warps, n_sm)) k_backward_mma<1, kFragReg><<<blocks, kMW * 32, bwd_smem(1, kFragReg), st>>>(m, p, w, Gw);
        else k_backward_mma<1, kFragShared><<<blocks, kMW * 32, bwd_smem(1, kFragShared), st>>>(m, p, w, Gw);
    } else if (m.Mp == 64) {
        k_backward_mma<2, kFragShared><<<blocks, kMW * 32, bwd_smem(2, kFragShared), st>>>(m, p, w, Gw);
    } else {
        k_backward_mma<4, kFragGlobal><<<blocks, kMW * 32, bwd_smem(4, kFragGlobal), st>>>(m, p, w, Gw);
    }
}

// forward + backward in one launch (pass 0); returns false when this input takes separate kernels
bool launch_recursions_mma(const Model &m, const Plan &p, const Work &w, int n_sm, const RecOpts &o, cudaStream_t st)
{
    if (o.fused == 0 || !mma_forward_pays(p.n_chunks, n_sm, m.Mp, o)) return false;
    const int G = chunks_per_warp(p.n_chunks, n_sm, m.Mp, o);
    const int warps = (p.n_chunks + G - 1) / G, wpc = warps_per_cta(warps, n_sm), blocks = (warps + wpc - 1) / wpc, Gw = pack_gw(G, wpc);
    configure_once(blocks, n_sm, smem_pair(m, o, warps, n_sm));
    const int grid = 2 * blocks;
    auto mx = [](size_t a, size_t b) { return a > b ? a : b; };
    if (m.Mp == 32) {
        const int nkc = cached_keys(m, o);
        if (use_reg_frags(warps, n_sm))
            k_recursions_mma<1, kFragReg><<<grid, kMW * 32, mx(fwd_smem(1, kFragReg, nkc), bwd_smem(1, kFragReg)), st>>>(m, p, w, Gw, nkc, blocks);
        else
            k_recursions_mma<1, kFragShared><<<grid, kMW * 32, mx(fwd_smem(1, kFragShared, nkc), bwd_smem(1, kFragShared)), st>>>(m, p, w, Gw, nkc, blocks);
    } else if (m.Mp == 64) {
        k_recursions_mma<2, kFragShared><<<grid, kMW * 32, mx(fwd_smem(2, kFragShared, 0), bwd_smem(2, kFragShared)), st>>>(m, p, w, Gw, 0, blocks);
    } else {
        k_recursions_mma<4, kFragGlobal><<<grid, kMW * 32, mx(fwd_smem(4, kFragGlobal, 0), bwd_smem(4, kFragGlobal)), st>>>(m, p, w, Gw, 0, blocks);
    }
    return true;
}

}  // namespace smcb

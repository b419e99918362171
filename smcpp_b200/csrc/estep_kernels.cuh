// smcpp_b200 -- device side of the E-step (sm_100a).
//
// Pipeline per E-step:
//   k_setup      : padded / transposed operand tables, fl32(e_k(j) * Td(i,j)) step matrices
//   k_forward    : one warp per chunk, scaled forward recursion with the reference's float semantics
//                  (reference src/hmm.cpp:58-96); chunk starts come from a burn-in over the preceding
//                  blocks and are verified against the previous chunk's end (k_check_forward)
//   k_backward   : one warp per chunk, beta recursion (reference src/hmm.cpp:97-149, recursion part)
//   k_stats      : block-parallel accumulation of the xi / gamma sufficient statistics
//   k_finalize   : per-contig reduction, eigenbasis -> state basis, "o Td", floors (src/hmm.cpp:150-152)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smcb {

constexpr int kMaxMp = 128;       // padded state count limit (M <= 128)
// per-block code (Plan::kcode): bits 0-15 key id, bits 16-31 1 + eigen index if span > 1 (else 0)
typedef uint32_t kcode_t;
constexpr int kKeyBits = 16;
constexpr int kKeyMask = 0xffff;
constexpr int kMaxKeys = 65535;   // distinct observation keys; also bounds the keys that occur with span > 1

// Per-E-step operands on the device.  All matrices are padded to Mp = 32*ceil(M/32) in the fast
// (lane) dimension; pads are zero.
struct Model {
    int M, Mp, K, n_eig;
    int hot_eig;             // eigen index with the most span>1 blocks (kept in registers by the M<=32 kernels), or -1
    const double *pi;        // [Mp]
    const double *Td;        // [Mp][Mp]  Td[i*Mp + j] = Td(i,j)   (all tables: Mp rows, zero padded)
    const double *TdT;       // [Mp][Mp]  TdT[j*Mp + i] = Td(i,j)
    const float *A32;        // [K][Mp][Mp] A32[(k*Mp + i)*Mp + j] = fl32(e_k(j) * Td(i,j))
    const double *E;         // [K][Mp]
    const int *eig_of_key;   // [K]  -1 or eigen index
    const int *key_of_eig;   // [n_eig]
    const double *P;         // [n_eig][Mp][Mp]  P[(e*Mp + i)*Mp + a] = P_r(i,a)
    const double *PT;        // [n_eig][Mp][Mp]  PT[(e*Mp + a)*Mp + j] = P_r(j,a)
    const double *Pinv;      // [n_eig][Mp][Mp]  Pinv[(e*Mp + a)*Mp + i] = Pinv_r(a,i)
    const double *PinvT;     // [n_eig][Mp][Mp]  PinvT[(e*Mp + i)*Mp + a] = Pinv_r(a,i)
    const double *dsc;       // [n_eig][Mp] d_r / scale
    const double *logd;      // [n_eig][Mp] log|dsc|
    const double *dr;        // [n_eig][Mp] d_r
    const double *scale;     // [n_eig]
    const double *logscale;  // [n_eig]
    // tensor-path (Mp in {32, 64, 128}) operand tables, see recursion_mma.cu: matrices in mma B-fragment order,
    // vectors and the float step matrices permuted to the per-lane state order st(q, idx)
    const double *F_Td;      // [Mp*Mp]
    const double *F_P;       // [n_eig][Mp*Mp]  W = P_r
    const double *F_PT;      // [n_eig][Mp*Mp]  W = P_r^T
    const double *F_Pinv;    // [n_eig][Mp*Mp]  W = Pinv_r
    const double *F_PinvT;   // [n_eig][Mp*Mp]  W = Pinv_r^T
    const double *Eq;        // [K][Mp]
    const float *A32q;       // [K][Mp][4][Mp/4]
    // d~^span tables: one row per (eigen key, distinct span) -- the reference tabulates per (span, key) too
    // (span_Qs, src/transition_bundle.cpp:29-58); rows are built per E-step by k_setup_pwtab
    const double *pwtab;     // [n_eig][n_span][Mp]
    const double *pwq;       // the same rows in the tensor path's per-lane state order: pwq[q*(Mp/4) + idx] = pw[st(q, idx)]
    int hot_keys[4];         // most frequent span-1 keys of the data set (descending), -1 = none
    unsigned long long c_negzero2, c_one2;   // packed float pairs {-0.0f, -0.0f} and {1.0f, 1.0f}, see recursion_mma.cu (pack2)
    const int32_t *span_list;// [n_span] distinct spans > 1 of the data set
    int n_span;
    const double *invdiff;   // [n_eig][Mp][Mp] 1/(d~_a - d~_b), built only for posterior decoding
    // Irregular spectra (the reference keeps only the real parts of a complex eigensystem, include/transition_bundle.h:19-24,
    // so P_r Pinv_r != I; a negative eigenvalue makes its span tables NaN, src/transition_bundle.cpp:46-50): such eigen
    // keys take the reference's literal per-block formulas (k_stats_literal) instead of the O(M^2) displacement form.
    const uint8_t *irregular;  // [n_eig] or nullptr
    int literal;               // 1: some key is irregular -- sequential chains, generic kernels, NaN semantics of src/hmm.cpp:123-127
};

// Static per-dataset layout (set_contigs) + per-plan chunking.
struct Plan {
    int n_contigs, n_chunks, n_slabs;
    int chunk_blocks, burn_in, slab_blocks;   // burn_in: backward recursion (fp64 beta, verified to 1e-10)
    int burn_in_fwd;                          // forward recursion (float alpha_hat, verified to the float noise floor): shorter
    int64_t total_blocks;
    // per block (concatenated over contigs)
    const int32_t *span;     // [total]
    const kcode_t *kcode;    // [total]  bits 0-15: key id; bits 16-31: 1 + eigen index if span > 1, else 0
    const int32_t *span_id;  // [total]  index into Model::span_list (0 for span-1 blocks)
    // per contig
    const int64_t *blk_off;  // [C+1] first global block of contig
    const int64_t *col_off;  // [C]   first alpha column of contig (chunk c at col_off + c*(chunk_blocks+1))
    const int32_t *chunk_off;// [C+1] first chunk of contig
    const int32_t *slab_off; // [C+1]
    // per chunk
    const int32_t *ch_contig;// [n_chunks]
    const int32_t *ch_start; // [n_chunks] first block (within contig)
    const int32_t *ch_len;   // [n_chunks]
    // per slab
    const int32_t *sl_contig;// [n_slabs]
    const int32_t *sl_start; // [n_slabs]
    const int32_t *sl_len;   // [n_slabs]
    const uint32_t *sl_mask; // [n_slabs][mask_words] bit set: bit 0 = has span-1 blocks; bit 1+e = has span>1 blocks of eigen key e
    const uint32_t *ct_mask; // [n_contigs][mask_words] the same per contig (union over its slabs)
    int mask_words;          // ceil((1 + n_eig) / 32)
    // processing order of the statistics kernel: per slab [span-1 blocks sorted by key | eigen key 0 | eigen key 1 ...]
    const int2 *srec;        // [total]  (block index within the contig, key id for span-1 blocks / span id otherwise), at the slab's own offset
    const int32_t *seg;      // [n_slabs][n_eig + 2] segment boundaries (relative to the slab start)
    // M <= 32: work items of the span>1 statistics kernel (stats32.cu: k_stats32e).  The span>1 blocks of each
    // (contig, eigen key) are sorted by span id -- blocks of equal span share d~^span, so a whole run needs ONE
    // rank-1 accumulation plus one weighting per run instead of a rank-2 update per block -- and cut into items.
    int n_items;
    const int2 *erec;        // [#span>1 blocks]  (block index within the contig, span id), item after item
    const int64_t *it_start; // [n_items] first entry of the item in erec
    const int32_t *it_len;   // [n_items]
    const int32_t *it_contig;// [n_items]
    const int32_t *it_eig;   // [n_items]
    const int32_t *it_off;   // [C * n_eig + 1] first item of (contig, eigen key)
};

__host__ __device__ inline bool mask_bit(const uint32_t *mask, int bit) { return (mask[bit >> 5] >> (bit & 31)) & 1u; }

// Work buffers.
struct Work {
    float *alpha;            // [n_cols][Mp]  chunk-local alpha_hat columns (column 0 of a chunk = its start)
    float *cnorm;            // [total]       float forward normaliser of span-1 blocks (hmm.cpp:87)
    double *bvec;            // [total][Mp]   beta_l (span 1) or w_l = P_r^T beta_l (span > 1)
    double *uvec;            // [total][Mp]   M <= 64: u_l = Pinv_r alpha_hat_{l-1} of span>1 blocks, written by the forward pass
    double *Ritem;           // [n_items][Mp*Mp]  per-item partials of R_e (M <= 64)
    double *ditem;           // [n_items][Mp]     per-item partials of D_e
    float *start_used;       // [n_chunks][Mp]
    float *end_alpha;        // [n_chunks][Mp]
    float *end_alpha_prev;   // [n_chunks][Mp] snapshot of the previous sweep
    double *ll_chunk;        // [n_chunks]
    double *bstart_used;     // [n_chunks][Mp] beta the chunk started from (at its right end)
    double *beta_out;        // [n_chunks][Mp] beta at the chunk's left end
    double *beta_out_prev;   // [n_chunks][Mp]
    uint8_t *fwd_flag;       // [n_chunks] 1 = must be (re)run in the next sweep
    uint8_t *bwd_flag;
    uint8_t *fwd_rerun;      // [n_chunks] 1 = the chunk has been re-run from its neighbour's end value in this E-step       // [n_chunks]
    int *counters;           // [8]: 0 fwd flagged, 1 bwd flagged, 2 fwd max mismatch (float bits), 3 bwd max mismatch bits(hi) ..
    double *Xpart;           // [n_slabs][Mp*Mp]
    double *Rpart;           // [n_slabs][n_eig][Mp*Mp]
    double *dpart;           // [n_slabs][n_eig][Mp]
    double *gspart;          // [n_slabs][K][Mp]
    double *scratch;         // [C][2][Mp*Mp]  temporaries of k_finalize
    double *sums;            // [C][sum_stride] slab partials reduced per contig: X | R_e | D_e | gs
    double *sums_part;       // [C][reduce_parts()][sum_stride] first level of that reduction
    double *Xlit;            // [n_slabs][Mp*Mp]       literal-path partials of X (irregular eigen keys only; else nullptr)
    double *gslit;           // [n_slabs][n_eig][Mp]   literal-path partials of the gamma sums of the eigen keys
    double *lit_scratch;     // [n_slabs][2][Mp*Mp]
    int *nanpos;             // [C] literal path: last block (highest index) whose normalising constant is NaN, -1 = none.  The
                             // reference carries that constant into beta (log_C, src/hmm.cpp:117-127): every block before it is NaN too
    uint8_t *poison;         // [C][K] keys that occur at or before nanpos: their gamma sums are NaN in the reference
    // outputs (device)
    double *ll;              // [C]
    double *xisum;           // [C][M][M]
    double *gamma0;          // [C][M]
    double *gamma_sums;      // [C][K][M]
    double *reduced;         // [1 + M + M*M + K*M]
};

// (launch_setup fills the tables the Model's const pointers refer to)
void launch_setup(const Model &m, const double *pi_in, const double *T_in, const double *E_in, const double *P_in,
                  const double *Pinv_in, const double *d_in, const double *dsc_in, const double *scale_in, cudaStream_t st);
void launch_forward(const Model &m, const Plan &p, const Work &w, int pass, cudaStream_t st);
void launch_forward32(const Model &m, const Plan &p, const Work &w, int pass, cudaStream_t st);   // Mp == 32
void launch_backward32(const Model &m, const Plan &p, const Work &w, int pass, cudaStream_t st);  // Mp == 32
size_t sums_stride(const Model &m);
int reduce_parts();
int resident_warps32(int n_sm);
void launch_stats32(const Model &m, const Plan &p, const Work &w, cudaStream_t st, cudaStream_t st_runs);          // Mp == 32
void launch_stats64(const Model &m, const Plan &p, const Work &w, cudaStream_t st, cudaStream_t st_runs);          // Mp == 64
constexpr int kItemBlocks = 4096;   // span>1 blocks per work item of k_stats32e
void launch_setup_pwtab(const Model &m, int n_sm, cudaStream_t st);
void launch_setup_frags(const Model &m, cudaStream_t st);                                      // Mp == 32
struct RecOpts {             // per-context tuning of the tensor-path recursions (set_option)
    int cached_keys = 0;     // span-1 keys whose float step matrix is resident in shared memory (M <= 32), 0..4 (measured: no gain, the
                             // LSU data pipe carries the same bytes into the registers either way)
    int force_G = 0;         // chunks per warp pinned to 1 / 2 / 4 / 8 (0 = automatic)
    int fused = 0;           // forward and backward recursion in one launch (measured slower than two streams: 8.2 vs 7.65 ms on C3)
    int tiles = 1;           // MMA row tiles per warp at M <= 32: 1 = recursion_mma.cu (8 chunks per warp); 2 = recursion_mma2.cu (16 chunks per
                             // warp: half the LSU traffic per chunk step, but 255 registers leave one warp per scheduler -- measured slower, 9.9 vs 7.6 ms)
};
// M <= 32, several row tiles per warp (recursion_mma2.cu); st_fwd == st_bwd selects the one-launch form
bool launch_recursions_tiles(const Model &m, const Plan &p, const Work &w, int n_sm, const RecOpts &o, cudaStream_t st_fwd, cudaStream_t st_bwd);
int tiles_chunks_per_cta(int NM);
bool launch_recursions_mma(const Model &m, const Plan &p, const Work &w, int n_sm, const RecOpts &o, cudaStream_t st);
void launch_forward_mma(const Model &m, const Plan &p, const Work &w, int n_sm, const RecOpts &o, cudaStream_t st);   // Mp in {32, 64, 128}, pass 0, <= 8 chunks / warp
void launch_backward_mma(const Model &m, const Plan &p, const Work &w, int n_sm, const RecOpts &o, cudaStream_t st);
int resident_warps_mma(int n_sm, int Mp, const RecOpts &o);
bool mma_forward_pays(int n_chunks, int n_sm, int Mp, const RecOpts &o);
void set_recursion_carveout(int pct);
void launch_check_forward(const Model &m, const Plan &p, const Work &w, float tol0, float tol, cudaStream_t st);
void launch_backward(const Model &m, const Plan &p, const Work &w, int pass, cudaStream_t st);
void launch_check_backward(const Model &m, const Plan &p, const Work &w, double tol, cudaStream_t st);
// st_runs: stream of the span>1 statistics kernel (independent of the span-1 kernel; == st runs them back to back)
void launch_stats(const Model &m, const Plan &p, const Work &w, cudaStream_t st, cudaStream_t st_runs);
void launch_finalize(const Model &m, const Plan &p, const Work &w, cudaStream_t st);
void launch_stats_literal(const Model &m, const Plan &p, const Work &w, cudaStream_t st);
void launch_posterior(const Model &m, const Plan &p, const Work &w, double *gamma, const int64_t *gcol_off, int normalise, int n_sm, cudaStream_t st);
void launch_gather_alpha(const Model &m, const Plan &p, const Work &w, int contig, float *out, int n_sm, cudaStream_t st);
int stats_smem_bytes(const Model &m);
// M-step objective from the resident statistics (qfunc.cu); q[4 * (1 + D)]: per term the value and D derivatives
void launch_q(int C, int M, int K, int D, const double *pi, const double *T, const double *E, const double *dpi, const double *dT,
              const double *dE, const uint8_t *present, const int32_t *key_nb, const double *gamma0, const double *xisum,
              const double *gamma_sums, double *terms, double *q, cudaStream_t st);
void launch_fp64_peak(double *sink, int iters, int n_sm, cudaStream_t st);

}  // namespace smcb

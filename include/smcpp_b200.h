/*
 * smcpp_b200 -- C ABI of the B200-native E-step of SMC++ (libsmcpp_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point names
 * the piece of the reference (popgenmethods/smcpp @ 6779faec) that it replaces; INTEGRATION.md shows the
 * host-side binding a maintainer adds to the reference's InferenceManager to call it.
 *
 * One context drives ONE GPU (one process per GPU, or several contexts in one process).  Contigs are
 * independent HMMs (reference src/inference_manager.cpp:89-94), so a multi-GPU run gives every context
 * its shard of the contigs and sums the packed statistics (`reduced`) with one all-reduce.
 *
 * All matrices cross this boundary ROW-major, X[i*M + j] = X(i, j) (what NumPy and the reference's
 * store_matrix(), src/common.cpp:8-11, produce).  All functions return 0 on success; on failure they
 * return non-zero and smcpp_b200_last_error() describes it.  Nothing here throws or calls back into the
 * host language, and everything may be called with the Python GIL released (the reference calls Estep
 * under `with nogil`, smcpp/_smcpp.pyx:189-190, declared without `except +`, smcpp/_smcpp.pxd:49).
 * There is no CPU fallback: without a usable CUDA device create() fails.
 */
#ifndef SMCPP_B200_H
#define SMCPP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smcpp_b200_ctx smcpp_b200_ctx;

/* ABI version of this header (bumped on any signature change). */
int smcpp_b200_abi_version(void);

/* Create / destroy a context bound to CUDA device `device`.
 * Replaces: nothing in the reference (it has no device); lifetime = the InferenceManager's. */
int smcpp_b200_create(smcpp_b200_ctx **out, int device);
void smcpp_b200_destroy(smcpp_b200_ctx *ctx);
const char *smcpp_b200_last_error(const smcpp_b200_ctx *ctx); /* ctx may be NULL: error of the last failed create() */

/* Tuning knobs (optional).  name in {"chunk_blocks", "burn_in_blocks" (both passes), "burn_in_blocks_forward", "target_warps", "slab_blocks", "fwd_tol",
 * "fwd_tol_burn_in", "bwd_tol", "max_sweeps", "max_restarts", "force_sequential", "mma_min_chunks", "force_mma_forward", "chunks_per_warp",
 * "fwd_cached_keys", "fused_recursions"} (all per context; the last three exist for tests / experiments).
 * Note on "fwd_tol_burn_in" (default 1e-6): the forward pass of a chunk starts from a burn-in over the preceding blocks and
 * is accepted when its float alpha_hat start agrees with the neighbour chunk's end to this relative tolerance -- the noise
 * floor of two float trajectories with different histories.  This is an approximation the reference does not make; it is
 * what keeps the log-likelihood within ~1e-11 relative instead of bit-identical ("force_sequential" = 1 gives the literal chain).
 * Returns non-zero for an unknown name. */
int smcpp_b200_set_option(smcpp_b200_ctx *ctx, const char *name, double value);

/*
 * One-time upload of the observations of this context's contigs.
 * Replaces: InferenceManager::InferenceManager / map_obs / fill_targets / populate_emission_probs
 *           (reference src/inference_manager.cpp:21-54, 180-188, 232-254, 190-211) and HMM::HMM's
 *           bookkeeping (src/hmm.cpp:8-28).
 *   obs[c]    int32 row-major [lengths[c]][1 + 3*npop], rows [span, a_1, b_1, nb_1 (, a_2, b_2, nb_2)];
 *             only read during this call (the reference keeps the pointers; we copy to the device).
 *   keys      optional global key table [n_keys][3*npop] in the reference's std::map order
 *             (lexicographic, include/block_key.h:51-60); NULL = derive it from these contigs.  A key
 *             of the data that is missing from an explicit table is an error.  Multi-GPU runs pass
 *             the union over all shards so every context packs gamma_sums identically.
 * Errors: span <= 0 -> "data are malformed: span <= 0" (reference src/inference_manager.cpp:243-244).
 */
int smcpp_b200_set_contigs(smcpp_b200_ctx *ctx, int n_contigs, const int32_t *const *obs,
                           const int32_t *lengths, int npop, const int32_t *keys, int n_keys);

/* Key universe after set_contigs(): K, the K x 3*npop table, which keys occur with span > 1
 * (the reference's eigensystem targets, src/inference_manager.cpp:245-246) and, per contig, which
 * keys occur at all (the reference's gamma_sums maps hold exactly those, src/hmm.cpp:51-53,69). */
int smcpp_b200_num_keys(const smcpp_b200_ctx *ctx);
int smcpp_b200_get_keys(const smcpp_b200_ctx *ctx, int32_t *keys /* K*3P */);
int smcpp_b200_num_eig_keys(const smcpp_b200_ctx *ctx);
int smcpp_b200_get_eig_keys(const smcpp_b200_ctx *ctx, int32_t *key_idx /* n_eig */);
int smcpp_b200_get_key_present(const smcpp_b200_ctx *ctx, uint8_t *present /* C*K */);
int64_t smcpp_b200_total_blocks(const smcpp_b200_ctx *ctx);

/*
 * Eigensystems of diag(e_key) * Td^T for the keys returned by get_eig_keys(), computed by the library
 * (real non-symmetric QR algorithm, real parts kept like the reference).
 * Replaces: the EigenSolver loop of TransitionBundle::update (reference src/transition_bundle.cpp:14-25)
 *           and struct eigensystem (include/transition_bundle.h:9-30).
 * Outputs [n_eig][M][M] / [n_eig][M] / [n_eig]; cplx[e] != 0 when an eigenvalue had an imaginary part.
 */
int smcpp_b200_eigensystems(smcpp_b200_ctx *ctx, int M, const double *T, const double *E /* K*M */,
                            double *P, double *Pinv, double *d, double *d_scaled, double *scale, int32_t *cplx);

/* Context-free forms of the same host routine (usable without a GPU; the unit tests call these):
 * one general real matrix A[n][n] -> P, Pinv (real parts), Re(d), Im(d); and the per-key batch. */
int smcpp_b200_host_eig(int n, const double *A, double *P, double *Pinv, double *d_re, double *d_im);
int smcpp_b200_host_eigensystems(int M, int K, int n_eig, const int32_t *eig_key_idx, const double *T,
                                 const double *E, double *P, double *Pinv, double *d, double *d_scaled,
                                 double *scale, int32_t *cplx);

/*
 * Host builders of the per-E-step HMM inputs from the model (value parts; context-free, no GPU needed).
 *   eta: n_pieces piecewise-constant sizes a[] with lengths s[] (what SMCModel.stepwise_values() hands to
 *        setParams, reference smcpp/_smcpp.pyx:205-221); hidden_states: M+1 boundaries, last = inf.
 * Replace: recompute_initial_distribution (reference src/inference_manager.cpp:56-69);
 *          compute_transition / HJTransition (src/transition.cpp:133-262; 113-bit instead of 256-bit products);
 *          recompute_emission_probs + incorporate_theta + construct_bins (src/inference_manager.cpp:329-482,
 *          src/conditioned_sfs.cpp:100-148) with the conditioned SFS supplied by the caller as
 *          sfs[M][na[0]+1][sfs_dim] (the reference's DummySFS contract, include/conditioned_sfs.h:46-67).
 * host_emission fails with the reference's messages ("probability vector not in [0, 1]", "s<=0", ...).
 */
int smcpp_b200_host_initial_distribution(int M, const double *hidden_states, int n_pieces, const double *a,
                                         const double *s, double *pi);
int smcpp_b200_host_average_coal_times(int M, const double *hidden_states, int n_pieces, const double *a,
                                       const double *s, double *out);
int smcpp_b200_host_transition(int M, const double *hidden_states, int n_pieces, const double *a, const double *s,
                               double rho, double *T);
int smcpp_b200_host_emission(int npop, const int32_t *n, const int32_t *na, int M, const double *hidden_states,
                             int n_pieces, const double *a, const double *s, double theta, double alpha,
                             double pol_err, const double *sfs, int K, const int32_t *keys, double *E,
                             char *errbuf, int errbuf_len);

/*
 * One E-step over all contigs of this context.
 * Replaces: InferenceManager::Estep -> TransitionBundle::update(T, true) -> parallel_do(HMM::Estep)
 *           (reference src/inference_manager.cpp:108-114, src/transition_bundle.cpp:3-61,
 *           src/hmm.cpp:45-153), with pi / T / emission table as do_dirty_work() leaves them (:213-229).
 * Inputs (host pointers):
 *   pi[M], T[M][M] (value part of the transition), E[K][M] (emission_probs in key order),
 *   n_eig eigensystems in the order of get_eig_keys(): P, Pinv [n_eig][M][M], d, d_scaled [n_eig][M],
 *   scale [n_eig]; pass P == NULL to have the library compute them (smcpp_b200_eigensystems).
 * Outputs (host pointers, caller-owned; any may be NULL):
 *   ll[C]                 HMM::ll
 *   xisum[C][M][M]        HMM::xisum   (after the final "o Td" and 1e-20 floor, src/hmm.cpp:151-152)
 *   gamma0[C][M]          HMM::gamma.col(0)
 *   gamma_sums[C][K][M]   HMM::gamma_sums, zero rows for keys absent from the contig
 *   reduced[1+M+M*M+K*M]  sum over this context's contigs of [ll | gamma0 | xisum | gamma_sums]
 *                         (the all-reduce payload of a multi-GPU run; reference Q() is linear in it)
 */
int smcpp_b200_estep(smcpp_b200_ctx *ctx, int M, const double *pi, const double *T, const double *E,
                     int n_eig, const double *P, const double *Pinv, const double *d, const double *d_scaled,
                     const double *scale, double *ll, double *xisum, double *gamma0, double *gamma_sums,
                     double *reduced);

/* Device pointer (on this context's GPU) to the packed `reduced` vector of the last estep(), length
 * 1 + M + M*M + K*M doubles, valid until the next estep(); lets the caller all-reduce it in place. */
int smcpp_b200_reduced_device_ptr(smcpp_b200_ctx *ctx, void **ptr, int64_t *count);

/* Copies that vector into caller-owned DEVICE memory on the same GPU (e.g. a torch tensor that is then
 * all-reduced over NCCL); synchronises the context's stream before returning. */
int smcpp_b200_copy_reduced_to_device(smcpp_b200_ctx *ctx, void *dst_device, int64_t count);

/* Same E-step, but nothing is copied back to the host: results stay in the device buffers (per-contig
 * outputs and `reduced`).  Used to time the kernels alone; follow with smcpp_b200_fetch() for values. */
int smcpp_b200_estep_device(smcpp_b200_ctx *ctx, int M, const double *pi, const double *T, const double *E,
                            int n_eig, const double *P, const double *Pinv, const double *d,
                            const double *d_scaled, const double *scale, int upload_inputs);
int smcpp_b200_fetch(smcpp_b200_ctx *ctx, double *ll, double *xisum, double *gamma0, double *gamma_sums,
                     double *reduced);

/*
 * The M-step objective from the statistics that are resident on the device.
 * Replaces: HMM::Q (reference src/hmm.cpp:155-193) summed over contigs by InferenceManager::Q
 *           (src/inference_manager.cpp:116-126): per contig each term is accumulated in the reference's element order with
 *           its doubly compensated summation (include/common.h:27-46), then the contigs are added in order.
 *   q[4]  = { sum log(pi) gamma0,  sum_{keys with nb == 0} log(e_key) gamma_sums_key,  the same for nb > 0,  sum log(T) xisum },
 *           keys present in the contig only (src/hmm.cpp:166-172); a present key with a non-positive emission entry makes
 *           its class -infinity and ends the key loop (the reference warns there, :172-178).
 *   n_deriv > 0: the reference evaluates Q on autodiff scalars; hand in the derivative arrays of the inputs,
 *           dpi[n_deriv][M], dT[n_deriv][M][M], dE[n_deriv][K][M], and get dq[4][n_deriv] (d log x = dx / x contracted with the
 *           same statistics in the same order).  pi, T, E are the CURRENT model's (the M-step evaluates Q many times per E-step).
 * set_statistics loads per-contig statistics instead of computing them -- the reference's HMM constructor pre-fill
 * (gamma_sums = span * pi per key, xisum = gamma0 = 0: src/hmm.cpp:16-27, Q() before the first E-step) or a checkpoint.
 */
int smcpp_b200_q(smcpp_b200_ctx *ctx, int M, const double *pi, const double *T, const double *E, int n_deriv, const double *dpi,
                 const double *dT, const double *dE, double *q /* 4 */, double *dq /* 4 * n_deriv, or NULL */);
int smcpp_b200_set_statistics(smcpp_b200_ctx *ctx, int M, const double *xisum /* C*M*M */, const double *gamma0 /* C*M */,
                              const double *gamma_sums /* C*K*M */);

/*
 * Full posterior decoding, the `smc++ posterior` variant of the E-step.
 * Replaces: InferenceManager::saveGamma + the per-block gamma columns of HMM::Estep (reference
 *           include/inference_manager.h:40, src/hmm.cpp:48-49, 116-121, 134-136, 147-150) and getGammas()
 *           (src/inference_manager.cpp:136-142).
 * set_save_gamma(1) makes every following estep() also compute gamma[l][m] for l = 0..L of every contig (column 0 is
 * alpha_0 o beta_0; a column sums to the block's span); fetch_gamma copies one contig's [(L+1)][M] doubles to the host
 * (the reference's matrix is M x (L+1); the Python mirror returns the transposed view).
 * set_save_gamma(2): every column is also divided by its sum on the device -- the normalisation `smc++ posterior` applies on
 * the host before writing (reference smcpp/commands/posterior.py:104-106: g /= g.sum(axis=0)).
 */
int smcpp_b200_set_save_gamma(smcpp_b200_ctx *ctx, int on);
int smcpp_b200_fetch_gamma(smcpp_b200_ctx *ctx, int contig, double *out /* (L+1)*M */);

/*
 * Observation pre-processing on the device: the input side of the path.  One handle holds ONE contig's
 * rows [span, a_1, b_1, nb_1 (, a_2, b_2, nb_2)] on the GPU; every step replaces them, so the `smc++ estimate` chain
 * (reference smcpp/analysis/analysis.py:60-63) runs device-resident: upload -> thin -> bin -> recode_monomorphic ->
 * compress -> download (or hand the rows to set_contigs).  Results are bit-identical to the reference's functions.
 * Replaces:
 *   obs_thin               thin_data(data, thinning, offset)   reference smcpp/_estimation_tools.pyx:8-84 (including its
 *                          treatment of full-SFS bases whose a's sum to 2, :60-68)
 *   obs_bin                bin_observations(contig, w) + process_bin   smcpp/_estimation_tools.pyx:110-172;
 *                          a[npop] = distinguished lineages per population (contig.a)
 *   obs_recode_monomorphic RecodeMonomorphic._recode            smcpp/data_filter.py:331-336
 *   obs_compress           compress_repeated_obs(dataset)       smcpp/estimation_tools.py:51-61
 * obs_last_ms: device time (CUDA events) of the last step.  No CPU fallback: create() fails without a CUDA device.
 */
typedef struct smcpp_b200_obs smcpp_b200_obs;
int smcpp_b200_obs_create(smcpp_b200_obs **out, int device);
void smcpp_b200_obs_destroy(smcpp_b200_obs *o);
const char *smcpp_b200_obs_last_error(const smcpp_b200_obs *o); /* o may be NULL: error of the last failed create() */
int smcpp_b200_obs_upload(smcpp_b200_obs *o, const int32_t *rows, int64_t n_rows, int npop);
int smcpp_b200_obs_thin(smcpp_b200_obs *o, int thinning, int offset);
int smcpp_b200_obs_bin(smcpp_b200_obs *o, const int64_t *a, int64_t w);
int smcpp_b200_obs_recode_monomorphic(smcpp_b200_obs *o, const int64_t *a);
int smcpp_b200_obs_compress(smcpp_b200_obs *o);
/* The front of the chain (reference smcpp/analysis/base.py:50-52): RecodeNonseg(cutoff) -> Compress -> BreakLongSpans.
 *   obs_recode_nonseg      recode_nonseg(contig, cutoff) with a cutoff given   smcpp/estimation_tools.py:88-114
 *   obs_break_long_spans   break_long_spans(contig, span_cutoff)              smcpp/estimation_tools.py:117-167: cuts the contig
 *                          at long missing rows; the pieces (each led by a one-base missing row) stay on the device one
 *                          after the other; piece_offsets returns n_pieces + 1 row offsets, select_piece makes the rows
 *                          [row_begin, row_end) of the broken array (normally one piece's range) the handle's current rows. */
int smcpp_b200_obs_recode_nonseg(smcpp_b200_obs *o, int64_t cutoff);
int smcpp_b200_obs_break_long_spans(smcpp_b200_obs *o, int64_t span_cutoff, int64_t *n_pieces);
int smcpp_b200_obs_piece_offsets(smcpp_b200_obs *o, int64_t *offsets /* n_pieces + 1 */);
int smcpp_b200_obs_select_piece(smcpp_b200_obs *o, int64_t piece, int64_t row_begin, int64_t row_end);
int64_t smcpp_b200_obs_rows(const smcpp_b200_obs *o);
float smcpp_b200_obs_last_ms(const smcpp_b200_obs *o);
int smcpp_b200_obs_download(smcpp_b200_obs *o, int32_t *rows /* rows() x (1 + 3 npop) */);

/*
 * Several GPUs in ONE process (the drop-in case: `smc++ estimate` is a single Python process).  A multi handle owns one
 * context per device, shards the contigs over them by longest processing time (contigs are independent HMMs, reference
 * src/inference_manager.cpp:89-94: `#pragma omp parallel for` over hmms), drives each device from its own host thread and
 * sums the packed statistics with ONE ncclAllReduce(ncclDouble, ncclSum) per E-step, in place in device memory (NCCL is
 * bound at run time; a handle with one device never loads it).
 * Replaces: InferenceManager::parallel_do (src/inference_manager.cpp:89-94) + the sums over HMMs of InferenceManager::Q /
 *           loglik (:116-126, :174-177).
 *   multi_set_contigs  derives the global key table (union over all contigs, std::map order) and uploads every shard;
 *   multi_estep        pi / T / E as in smcpp_b200_estep (the library computes the eigensystems); per-contig outputs in the
 *                      CALLER's contig order, key_present[C][K], reduced[1 + M + M*M + K*M] = the all-reduced statistics;
 *   multi_context      the i-th device's context (options, statistics, smcpp_b200_q on that device's contigs).
 */
typedef struct smcpp_b200_multi smcpp_b200_multi;
int smcpp_b200_multi_create(smcpp_b200_multi **out, const int *devices, int n_devices);
void smcpp_b200_multi_destroy(smcpp_b200_multi *m);
const char *smcpp_b200_multi_last_error(const smcpp_b200_multi *m); /* m may be NULL: error of the last failed create() */
int smcpp_b200_multi_num_devices(const smcpp_b200_multi *m);
int smcpp_b200_multi_context(smcpp_b200_multi *m, int i, smcpp_b200_ctx **ctx);
int smcpp_b200_multi_set_contigs(smcpp_b200_multi *m, int n_contigs, const int32_t *const *obs, const int32_t *lengths, int npop);
int smcpp_b200_multi_num_keys(const smcpp_b200_multi *m);
int smcpp_b200_multi_get_keys(const smcpp_b200_multi *m, int32_t *keys /* K*3P */);
int smcpp_b200_multi_get_shard(const smcpp_b200_multi *m, int device_index, int32_t *contigs /* or NULL */, int32_t *n);
int smcpp_b200_multi_estep(smcpp_b200_multi *m, int M, const double *pi, const double *T, const double *E, double *ll,
                           double *xisum, double *gamma0, double *gamma_sums, uint8_t *key_present, double *reduced);

/* Diagnostics of the last estep(): see smcpp_b200_stats_t. */
typedef struct smcpp_b200_stats_t {
    int32_t n_chunks;        /* chunks the contigs were split into */
    int32_t chunk_blocks;    /* blocks per chunk */
    int32_t burn_in_blocks;  /* burn-in length used for chunk boundaries */
    int32_t fwd_sweeps;      /* forward passes executed (1 = burn-in accepted everywhere) */
    int32_t bwd_sweeps;
    int32_t fwd_redone;      /* chunks re-run because the boundary check failed */
    int32_t bwd_redone;
    int32_t kernel_launches; /* CUDA kernels launched by the last estep() */
    float ms_setup, ms_forward, ms_backward, ms_stats, ms_finalize, ms_total; /* CUDA-event times */
    double fwd_max_mismatch, bwd_max_mismatch;
    float ms_forward_only;         /* forward pass 0 alone; ms_forward = both recursions incl. sweeps, ms_backward = backward pass 0 alone */
    int32_t mma_rounds, mma_steps; /* tensor-path forward kernel: warp rounds and committed chunk-steps (efficiency = steps / (8 rounds)) */
    int32_t literal_keys;          /* eigen keys with an irregular spectrum (complex pair / negative eigenvalue) in the last estep():
                                      they -- and the whole E-step's chain, run sequentially -- follow the reference's literal formulas */
    int32_t restarts;              /* pass 0 was run again with a doubled burn-in this many times (many boundaries failed: slowly mixing model) */
    int32_t converged;             /* 0: the repair sweeps hit max_sweeps with chunk boundaries still failing -- estep() returned an error */
} smcpp_b200_stats_t;
int smcpp_b200_get_stats(const smcpp_b200_ctx *ctx, smcpp_b200_stats_t *out);

/* Measured FP64 FMA throughput of this GPU (TFLOP/s, 2 flop per FMA): a register-resident DFMA loop on
 * every SM, timed with CUDA events.  bench.py uses it as the compute roofline denominator (the driver's
 * MEASURED_PEAKS.json holds HBM and bf16 figures only). */
int smcpp_b200_fp64_peak(smcpp_b200_ctx *ctx, double *tflops);

/* CUDA stream handle (cudaStream_t) the context launches on, for event timing by the caller. */
int smcpp_b200_stream(smcpp_b200_ctx *ctx, void **stream);

/* Debug/verification taps (used by tests): copy out the stored float alpha_hat columns of one contig,
 * [L+1][M], as the reference's HMM::alpha_hat (include/hmm.h:35) would hold them. */
int smcpp_b200_debug_alpha_hat(smcpp_b200_ctx *ctx, int contig, float *out /* (L+1)*M */);

#ifdef __cplusplus
}
#endif
#endif /* SMCPP_B200_H */

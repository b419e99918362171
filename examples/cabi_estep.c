/*
 * Plain-C consumer of the C ABI (include/smcpp_b200.h): what a host written in any language binds.
 *
 *   gcc -std=c99 -Wall -Iinclude examples/cabi_estep.c -Lsmcpp_b200 -lsmcpp_b200 -Wl,-rpath,$PWD/smcpp_b200 -lm -o cabi_estep
 *   ./cabi_estep            # needs a CUDA device; prints the log-likelihood and a few statistics of a toy data set
 *
 * The toy: one population, M = 4 hidden states, 2 contigs of run-length-encoded rows [span, a, b, nb] (reference
 * README.rst:519-566), a hand-made transition matrix / emission table.  tests/test_layout.py compiles this file on every
 * box (the header must stay valid C99); tests/test_gpu_parity.py runs it and compares with the Python binding.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "smcpp_b200.h"

#define M 4

static int fail(smcpp_b200_ctx *ctx, const char *what)
{
    fprintf(stderr, "%s: %s\n", what, smcpp_b200_last_error(ctx));
    return 1;
}

int main(void)
{
    /* observations: alternating runs of non-segregating bases and single sites, one missing stretch */
    enum { L0 = 41, L1 = 23 };
    static int32_t c0[L0][4], c1[L1][4];
    int i, j, k;
    for (i = 0; i < L0; ++i) {
        if (i % 2 == 0) { c0[i][0] = 3 + 7 * (i % 5); c0[i][1] = (i == 20) ? -1 : 0; c0[i][2] = 0; c0[i][3] = 0; }
        else { c0[i][0] = 1; c0[i][1] = 1 + (i % 3 == 0); c0[i][2] = (i % 7 == 1) ? 2 : 0; c0[i][3] = (i % 7 == 1) ? 3 : 0; }
    }
    for (i = 0; i < L1; ++i) {
        if (i % 2 == 0) { c1[i][0] = 2 + 11 * (i % 3); c1[i][1] = 0; c1[i][2] = 0; c1[i][3] = 0; }
        else { c1[i][0] = 1; c1[i][1] = 1; c1[i][2] = 0; c1[i][3] = 0; }
    }
    const int32_t *obs[2] = {&c0[0][0], &c1[0][0]};
    const int32_t lens[2] = {L0, L1};

    smcpp_b200_ctx *ctx = NULL;
    if (smcpp_b200_create(&ctx, 0)) return fail(NULL, "create");
    if (smcpp_b200_set_contigs(ctx, 2, obs, lens, 1, NULL, 0)) return fail(ctx, "set_contigs");
    const int K = smcpp_b200_num_keys(ctx);
    int32_t *keys = (int32_t *)malloc((size_t)K * 3 * sizeof(int32_t));
    smcpp_b200_get_keys(ctx, keys);

    /* model inputs: pi, a diagonally dominant transition matrix, emissions that depend on the key */
    double pi[M], T[M * M], *E = (double *)malloc((size_t)K * M * sizeof(double));
    for (i = 0; i < M; ++i) pi[i] = (i + 1.0) / (M * (M + 1) / 2.0);
    for (i = 0; i < M; ++i) {
        double s = 0.0;
        for (j = 0; j < M; ++j) { T[i * M + j] = (i == j) ? 50.0 : 1.0 / (1.0 + abs(i - j)); s += T[i * M + j]; }
        for (j = 0; j < M; ++j) T[i * M + j] /= s;
    }
    for (k = 0; k < K; ++k)
        for (i = 0; i < M; ++i) {
            const int a = keys[3 * k], b = keys[3 * k + 1];
            E[k * M + i] = a < 0 ? 1.0 : (a == 0 && b == 0) ? exp(-0.02 * (i + 1)) : 0.01 * (i + 1) * (1 + a) / (1.0 + b);
        }

    double ll[2], xisum[2 * M * M], gamma0[2 * M];
    double *gsums = (double *)malloc((size_t)2 * K * M * sizeof(double)), *reduced = (double *)malloc((size_t)(1 + M + M * M + K * M) * sizeof(double));
    /* P == NULL: the library computes the eigensystems of diag(e_key) T^T itself */
    if (smcpp_b200_estep(ctx, M, pi, T, E, 0, NULL, NULL, NULL, NULL, NULL, ll, xisum, gamma0, gsums, reduced)) return fail(ctx, "estep");

    smcpp_b200_stats_t st;
    smcpp_b200_get_stats(ctx, &st);
    printf("abi=%d K=%d eig_keys=%d blocks=%lld kernels=%d\n", smcpp_b200_abi_version(), K, smcpp_b200_num_eig_keys(ctx),
           (long long)smcpp_b200_total_blocks(ctx), st.kernel_launches);
    printf("ll %.12f %.12f sum %.12f\n", ll[0], ll[1], reduced[0]);
    printf("xisum[0][0][0] %.12e gamma0[1][%d] %.12e\n", xisum[0], M - 1, gamma0[M + M - 1]);
    free(keys); free(E); free(gsums); free(reduced);
    smcpp_b200_destroy(ctx);
    return 0;
}

/*
 * TEST INFRASTRUCTURE -- CPU restatement of the reference's observation pre-processing (the input side of the
 * E-step), row by row in the reference's own sequential order.  Only tests/ may call this.
 *
 * Pinned against the reference itself: tests/golden/make_obs_golden.py runs the reference's Cython functions
 * (smcpp/_estimation_tools.pyx, compiled in a scratch directory) and pure-Python functions on seeded inputs and stores
 * inputs + outputs in tests/golden/obs/obs_pipeline.npz; tests/test_obs_pipeline.py checks this file against them bit for bit.
 *
 * Rows are int32 [span, a_1, b_1, nb_1 (, a_2, b_2, nb_2)], W = 1 + 3 npop columns.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* thin_data(data, thinning, offset=0): reference smcpp/_estimation_tools.pyx:8-84.
 * Every `thinning`-th base keeps its full observation, all others keep only the distinguished-lineage part (a), or
 * nothing when the a's sum to 2.  Quirks kept: b_view / nb_view are never filled in the reference (the assignments are
 * commented out, :37-39), so `all_eq_memview(b_view, nb_view)` is always true and a full-SFS base whose a's sum to 2
 * is emitted as an all-zero key (:60-68).  Returns the number of output rows (-1 if `cap` is too small). */
long smcb_oracle_thin(const int32_t *data, long K, int npop, int thinning, int offset, int32_t *out, long cap)
{
    const int W = 1 + 3 * npop;
    long r = 0;
    int i = offset;
    int32_t thin[6], nonseg[6];
    memset(nonseg, 0, sizeof nonseg);
    for (long j = 0; j < K; ++j) {
        const int32_t *row = data + j * W;
        int span = row[0], sa = 0;
        memset(thin, 0, sizeof thin);
        for (int n = 0; n < npop; ++n) { sa += row[1 + 3 * n]; thin[3 * n] = row[1 + 3 * n]; }
        if (sa == 2)
            for (int n = 0; n < npop; ++n) thin[3 * n] = 0;
        while (span > 0) {
            if (i < thinning && i + span >= thinning) {
                if (thinning - i > 1) {
                    if (r >= cap) return -1;
                    out[r * W] = thinning - i - 1;
                    memcpy(out + r * W + 1, thin, 3 * npop * sizeof(int32_t));
                    ++r;
                }
                if (r >= cap) return -1;
                out[r * W] = 1;
                if (sa == 2) memcpy(out + r * W + 1, nonseg, 3 * npop * sizeof(int32_t));   /* nb_view is all zero */
                else memcpy(out + r * W + 1, row + 1, 3 * npop * sizeof(int32_t));
                ++r;
                span -= thinning - i;
                i = 0;
            } else {
                if (r >= cap) return -1;
                out[r * W] = span;
                memcpy(out + r * W + 1, thin, 3 * npop * sizeof(int32_t));
                ++r;
                i += span;
                break;
            }
        }
    }
    return r;
}

/* process_bin(data, new_data, na, i, j, k, thin=0): reference smcpp/_estimation_tools.pyx:110-143 */
static void process_bin(const int32_t *data, int W, int32_t *new_row, const long *na, long i, long j)
{
    const int K = (W - 1) / 3;
    int max_sample_size = -2;
    long mq = 0;
    for (long q = i; q <= j; ++q) {
        int seg = 0, sample_size = 0;
        if (data[q * W] == 0) continue;
        for (int aa = 0; aa < K; ++aa) {
            const int bb = 3 * aa;
            sample_size += data[q * W + bb + 3];
            sample_size += (int)(na[aa] * (data[q * W + bb + 1] >= 0));
            seg += data[q * W + bb + 1] > 0 ? data[q * W + bb + 1] : 0;
        }
        if (sample_size > max_sample_size) { mq = q; max_sample_size = sample_size; }
        if (max_sample_size == 2 && seg == 1) mq = q;
    }
    for (int aa = 0; aa < K; ++aa) {
        const int bb = 3 * aa;
        new_row[bb + 1] = data[mq * W + bb + 1];
        new_row[bb + 2] = data[mq * W + bb + 2];
        new_row[bb + 3] = data[mq * W + bb + 3];
    }
}

/* bin_observations(contig, w): reference smcpp/_estimation_tools.pyx:146-172.  One output row per w-bp bin holding the
 * observation with the largest sample size in the bin (ties: the first; among 2-sample rows a heterozygous one wins,
 * the last such).  `data` is scratch: the reference splits rows in place while it walks (:160-164).  `out` must hold
 * total_bp / w + 1 rows.  Returns the number of output rows. */
long smcb_oracle_bin(int32_t *data, long K, int npop, const long *na, long w, int32_t *out)
{
    const int W = 1 + 3 * npop;
    long i = 0, j = 0, k = 0, seen = 0;
    while (j < K) {
        const long span = data[j * W];
        if (seen + span > w) {
            data[j * W] = (int32_t)(w - seen);
            process_bin(data, W, out + k * W, na, i, j);
            data[j * W] = (int32_t)(span - (w - seen));
            seen = 0;
            ++k;
            i = j;
        } else {
            ++j;
            seen += span;
        }
    }
    process_bin(data, W, out + k * W, na, i, j - 1);
    for (long x = 0; x <= k; ++x) out[x * W] = 1;
    return k + 1;
}

/* RecodeMonomorphic._recode: reference smcpp/data_filter.py:331-336 (in place): rows whose a equals the number of
 * distinguished lineages in every population and whose b equals nb get a = b = 0. */
void smcb_oracle_recode_monomorphic(int32_t *data, long K, int npop, const long *a)
{
    const int W = 1 + 3 * npop;
    for (long j = 0; j < K; ++j) {
        int all = 1;
        for (int n = 0; n < npop; ++n)
            if (data[j * W + 1 + 3 * n] != a[n] || data[j * W + 2 + 3 * n] != data[j * W + 3 + 3 * n]) all = 0;
        if (all)
            for (int n = 0; n < npop; ++n) data[j * W + 1 + 3 * n] = data[j * W + 2 + 3 * n] = 0;
    }
}

/* compress_repeated_obs(dataset): reference smcpp/estimation_tools.py:51-61.  Consecutive rows with the same key are
 * merged, spans added.  Returns the number of output rows. */
long smcb_oracle_compress(const int32_t *data, long K, int W, int32_t *out)
{
    long r = 0;
    for (long j = 0; j < K; ++j) {
        const int32_t *row = data + j * W;
        if (r > 0 && memcmp(out + (r - 1) * W + 1, row + 1, (W - 1) * sizeof(int32_t)) == 0) out[(r - 1) * W] += row[0];
        else { memcpy(out + r * W, row, W * sizeof(int32_t)); ++r; }
    }
    return r;
}

/* recode_nonseg(contig, cutoff) with a cutoff given: reference smcpp/estimation_tools.py:88-114 (in place).  Runs of
 * homozygosity longer than `cutoff` -- span > cutoff, every a == 0 and every b == 0 -- become missing (a = -1, nb = 0).
 * Returns the number of rows changed.  (cutoff = None only warns in the reference and changes nothing.) */
long smcb_oracle_recode_nonseg(int32_t *data, long K, int npop, long cutoff)
{
    const int W = 1 + 3 * npop;
    long changed = 0;
    for (long j = 0; j < K; ++j) {
        int run = data[j * W] > cutoff;
        for (int n = 0; n < npop; ++n)
            if (data[j * W + 1 + 3 * n] != 0 || data[j * W + 2 + 3 * n] != 0) run = 0;
        if (run) {
            for (int n = 0; n < npop; ++n) { data[j * W + 1 + 3 * n] = -1; data[j * W + 3 + 3 * n] = 0; }
            ++changed;
        }
    }
    return changed;
}

/* break_long_spans(contig, span_cutoff): reference smcpp/estimation_tools.py:117-167.  The contig is cut at every long
 * missing row (span >= cutoff, every a == -1, every nb == 0); the long rows are dropped and every piece gets a one-base
 * missing row in front (:129-131, :143).  Output: the pieces one after the other in `out` (K + 1 rows: every long row's
 * place is taken by the next piece's leading row) and piece_off[0..n_pieces] (row offsets into out).  Returns n_pieces. */
long smcb_oracle_break_long_spans(const int32_t *data, long K, int npop, long cutoff, int32_t *out, long *piece_off)
{
    const int W = 1 + 3 * npop;
    int32_t miss[7];
    memset(miss, 0, sizeof miss);
    miss[0] = 1;
    for (int n = 0; n < npop; ++n) miss[1 + 3 * n] = -1;
    long r = 0, pieces = 0;
    piece_off[pieces++] = 0;
    memcpy(out + r * W, miss, W * sizeof(int32_t));
    ++r;
    for (long j = 0; j < K; ++j) {
        const int32_t *row = data + j * W;
        int lng = row[0] >= cutoff;
        for (int n = 0; n < npop; ++n)
            if (row[1 + 3 * n] != -1 || row[3 + 3 * n] != 0) lng = 0;
        if (lng) {
            piece_off[pieces++] = r;
            memcpy(out + r * W, miss, W * sizeof(int32_t));
        } else {
            memcpy(out + r * W, row, W * sizeof(int32_t));
        }
        ++r;
    }
    piece_off[pieces] = r;
    return pieces;
}

// smcpp_b200 -- observation pre-processing on the device: the input side of the E-step.
//
// The reference prepares every contig for `smc++ estimate` with a chain of sequential Python / Cython passes
// (smcpp/analysis/analysis.py:60-63):
//     Thin              thin_data             smcpp/_estimation_tools.pyx:8-84
//     BinObservations   bin_observations      smcpp/_estimation_tools.pyx:110-172 (process_bin + the walk)
//     RecodeMonomorphic _recode               smcpp/data_filter.py:331-336
//     Compress          compress_repeated_obs smcpp/estimation_tools.py:51-61
// Each walks the rows carrying a counter (bases since the last full observation, bases seen in the current bin), which
// looks sequential but is a function of the row's absolute base-pair position only.  With P[j] = sum of the spans
// before row j (one prefix sum) every output row can be computed independently:
//   thin : base p keeps its full observation iff (offset + p + 1) % thinning == 0; row j turns into a closed-form number
//          of pieces (head, then full / gap pairs, tail) -> second prefix sum -> one thread per OUTPUT row;
//   bin  : bin m is the base range [m w, (m+1) w); one thread per bin walks the rows that overlap it (one row for almost
//          every bin of run-length-encoded data) and applies process_bin's selection rule;
//   compress: run heads are rows whose key differs from their predecessor's; an output row is a run head's key with the
//          distance to the next run head as its span.
// All of it is integer work bound by HBM traffic (16-28 B per row in and out); results are bit-identical to the
// reference (tests/test_obs_pipeline.py, against goldens produced by the reference's own functions).  The prefix sums
// use CUB's single-pass DeviceScan (toolkit library); everything else is hand-written.
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>

#include "../../include/smcpp_b200.h"

namespace {

constexpr int kMaxW = 7;   // 1 + 3 * 2 populations

template <typename T>
struct Buf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count)
    {
        if (count <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct Params {
    int W, npop;
    int64_t a[2];
};

// One row in registers.  One-population rows (W = 4) are 16 bytes: a single 128-bit load / store per row, so a warp's
// accesses coalesce into whole lines instead of 4-byte words at a 16-byte stride.
struct Row { int32_t v[kMaxW]; };
__device__ __forceinline__ Row load_row(const int32_t *rows, int64_t j, int W)
{
    Row r;
    if (W == 4) {
        const int4 x = *reinterpret_cast<const int4 *>(rows + j * 4);
        r.v[0] = x.x; r.v[1] = x.y; r.v[2] = x.z; r.v[3] = x.w;
    } else {
        for (int c = 0; c < W; ++c) r.v[c] = rows[j * W + c];
    }
    return r;
}
__device__ __forceinline__ void store_row(int32_t *rows, int64_t j, int W, const Row &r)
{
    if (W == 4) *reinterpret_cast<int4 *>(rows + j * 4) = make_int4(r.v[0], r.v[1], r.v[2], r.v[3]);
    else
        for (int c = 0; c < W; ++c) rows[j * W + c] = r.v[c];
}

// spans as int64 (input of the position prefix sum)
__global__ void k_obs_spans(const int32_t *rows, int64_t n, int W, int64_t *span64)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) span64[j] = rows[j * W];
}

// last index j in [lo, hi] with v[j] <= key (v non-decreasing, v[lo] <= key)
__device__ __forceinline__ int64_t last_le(const int64_t *v, int64_t lo, int64_t hi, int64_t key)
{
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (v[mid] <= key) lo = mid; else hi = mid - 1;
    }
    return lo;
}
// The outputs of a thread block are consecutive, so their source rows form one short range.  A partition pass finds the
// first source row of every 256-output chunk (one thread per chunk, all searches in parallel); the main kernels then stage
// the chunk's slice of the scanned array in shared memory and every thread searches it there (merge-path style).
constexpr int kChunk = 256;        // outputs per thread block iteration (= block size)
constexpr int kStage = 2048;       // slice entries staged in shared memory; longer slices are searched in global memory
__global__ void k_partition(const int64_t *v, int64_t n, int64_t n_chunks, int64_t step, int64_t *part)
{
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c <= n_chunks; c += (int64_t)gridDim.x * blockDim.x)
        part[c] = c < n_chunks ? last_le(v, 0, n - 1, c * step) : n - 1;
}
struct Slice { int64_t lo, hi; bool staged; };
__device__ __forceinline__ Slice stage_slice(const int64_t *v, const int64_t *part, int64_t c, int64_t *s_v)
{
    Slice sl;
    sl.lo = part[c];
    sl.hi = part[c + 1];
    sl.staged = sl.hi - sl.lo < kStage;
    __syncthreads();                                   // the previous iteration's readers are done with s_v
    if (sl.staged)
        for (int64_t i = threadIdx.x; i <= sl.hi - sl.lo; i += blockDim.x) s_v[i] = v[sl.lo + i];
    __syncthreads();
    return sl;
}
__device__ __forceinline__ int64_t slice_last_le(const Slice &sl, const int64_t *v, const int64_t *s_v, int64_t key)
{
    return sl.staged ? sl.lo + last_le(s_v, 0, sl.hi - sl.lo, key) : last_le(v, sl.lo, sl.hi, key);
}

// ---- thin_data ---------------------------------------------------------------------------------------------
// pieces of row j (span s, counter i at its first base): f0 = local index of its first full base
struct ThinRow { int64_t f0, nf, rem; int head; };
__device__ __forceinline__ ThinRow thin_row(int64_t p0, int64_t s, int64_t thinning, int64_t offset)
{
    ThinRow t;
    if (offset >= thinning) { t.f0 = 0; t.nf = 0; t.rem = 0; t.head = 0; return t; }   // the reference's counter never resets then
    const int64_t i = (offset + p0) % thinning;
    t.f0 = thinning - 1 - i;
    t.nf = t.f0 < s ? 1 + (s - 1 - t.f0) / thinning : 0;
    t.head = t.nf > 0 && t.f0 > 0;
    t.rem = t.nf > 0 ? s - 1 - (t.f0 + (t.nf - 1) * thinning) : 0;
    return t;
}
__device__ __forceinline__ int64_t thin_count(const ThinRow &t, int64_t thinning)
{
    if (t.nf == 0) return 1;
    return t.head + t.nf + (thinning > 1 ? t.nf - 1 : 0) + (t.rem > 0);
}

__global__ void k_thin_count(const int32_t *rows, const int64_t *pos, int64_t n, int W, int64_t thinning, int64_t offset, int64_t *cnt)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        cnt[j] = thin_count(thin_row(pos[j], rows[j * W], thinning, offset), thinning);
}

// one thread per output row: find its input row (binary search in the scanned counts), then its piece
__global__ void __launch_bounds__(kChunk) k_thin_write(const int32_t *rows, const int64_t *pos, const int64_t *off, const int64_t *part, int64_t n,
                                                       int64_t n_out, Params P, int64_t thinning, int64_t offset, int32_t *out)
{
    const int W = P.W;
    __shared__ int64_t s_v[kStage];
    const int64_t n_chunks = (n_out + kChunk - 1) / kChunk;
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const Slice sl = stage_slice(off, part, c, s_v);
        const int64_t t = c * kChunk + threadIdx.x;
        if (t >= n_out) continue;
        const int64_t j = slice_last_le(sl, off, s_v, t), u = t - off[j];   // last j with off[j] <= t
        const Row row = load_row(rows, j, W);
        const int64_t s = row.v[0];
        const ThinRow tr = thin_row(pos[j], s, thinning, offset);
        int sa = 0;
        for (int p = 0; p < P.npop; ++p) sa += row.v[1 + 3 * p];
        int64_t span;
        bool full = false;
        if (tr.nf == 0) span = s;
        else if (u < tr.head) span = tr.f0;
        else {
            const int64_t v = u - tr.head;
            if (thinning > 1) {
                if ((v & 1) == 0) { span = 1; full = true; }
                else span = (v >> 1) < tr.nf - 1 ? thinning - 1 : tr.rem;
            } else { span = 1; full = true; }
        }
        Row o = row;
        if (!(full && sa != 2)) {                                     // (else the base keeps its full observation)
            for (int c = 1; c < W; ++c) o.v[c] = 0;
            if (!full && sa != 2)                                     // thinned: the distinguished lineages only
                for (int p = 0; p < P.npop; ++p) o.v[1 + 3 * p] = row.v[1 + 3 * p];
            // sa == 2: thinned bases carry a = 0, and a full base too (the reference's nb_view is never filled, :60-68)
        }
        o.v[0] = (int32_t)span;
        store_row(out, t, W, o);
    }
}

// ---- bin_observations (+ RecodeMonomorphic fused on request) --------------------------------------------------
__global__ void __launch_bounds__(kChunk) k_bin(const int32_t *rows, const int64_t *pos, const int64_t *part, int64_t n, int64_t total, int64_t n_bins,
                                                Params P, int64_t w, int32_t *out)
{
    const int W = P.W;
    __shared__ int64_t s_v[kStage];
    const int64_t n_chunks = (n_bins + kChunk - 1) / kChunk;
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const Slice sl = stage_slice(pos, part, c, s_v);
        const int64_t m = c * kChunk + threadIdx.x;
        if (m >= n_bins) continue;
        const int64_t b0 = m * w, b1 = (m + 1) * w < total ? (m + 1) * w : total;
        const int64_t lo = slice_last_le(sl, pos, s_v, b0);           // last row that starts at or before b0
        int max_sample = -2;
        Row best = load_row(rows, lo, W);
        for (int64_t q = lo; q < n && pos[q] < b1; ++q) {             // rows that overlap [b0, b1) (process_bin skips empty parts)
            const Row row = load_row(rows, q, W);
            int sample = 0, seg = 0;
            for (int p = 0; p < P.npop; ++p) {
                sample += row.v[3 + 3 * p];
                sample += (int)(P.a[p] * (row.v[1 + 3 * p] >= 0));
                seg += row.v[1 + 3 * p] > 0 ? row.v[1 + 3 * p] : 0;
            }
            bool take = sample > max_sample;
            if (take) max_sample = sample;
            take = take || (max_sample == 2 && seg == 1);
            if (take) best = row;
        }
        best.v[0] = 1;
        store_row(out, m, W, best);
    }
}

__global__ void k_recode_monomorphic(int32_t *rows, int64_t n, Params P)
{
    const int W = P.W;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        int32_t *row = rows + j * W;
        bool all = true;
        for (int p = 0; p < P.npop; ++p) all = all && row[1 + 3 * p] == P.a[p] && row[2 + 3 * p] == row[3 + 3 * p];
        if (all)
            for (int p = 0; p < P.npop; ++p) row[1 + 3 * p] = row[2 + 3 * p] = 0;
    }
}

// ---- compress_repeated_obs ---------------------------------------------------------------------------------
__global__ void k_run_heads(const int32_t *rows, int64_t n, int W, int64_t *flag)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        bool head = j == 0;
        if (!head) {
            const Row x = load_row(rows, j, W), y = load_row(rows, j - 1, W);
            for (int c = 1; c < W; ++c) head = head || x.v[c] != y.v[c];
        }
        flag[j] = head;
    }
}
__global__ void k_run_starts(const int64_t *flag_scan, const int32_t *rows, int64_t n, int W, int64_t *start)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        bool head = j == 0;
        if (!head) {
            const Row x = load_row(rows, j, W), y = load_row(rows, j - 1, W);
            for (int c = 1; c < W; ++c) head = head || x.v[c] != y.v[c];
        }
        if (head) start[flag_scan[j]] = j;
    }
}
__global__ void k_compress_write(const int32_t *rows, const int64_t *pos, const int64_t *start, int64_t n, int64_t total, int64_t n_runs, int W,
                                 int32_t *out)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_runs; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = start[r];
        const int64_t end = r + 1 < n_runs ? pos[start[r + 1]] : total;
        Row o = load_row(rows, j, W);
        o.v[0] = (int32_t)(end - pos[j]);
        store_row(out, r, W, o);
    }
}

// ---- recode_nonseg / break_long_spans ---------------------------------------------------------------------------
__global__ void k_recode_nonseg(int32_t *rows, int64_t n, Params P, int64_t cutoff)
{
    const int W = P.W;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        Row r = load_row(rows, j, W);
        bool run = r.v[0] > cutoff;
        for (int p = 0; p < P.npop; ++p) run = run && r.v[1 + 3 * p] == 0 && r.v[2 + 3 * p] == 0;
        if (run) {
            for (int p = 0; p < P.npop; ++p) { r.v[1 + 3 * p] = -1; r.v[3 + 3 * p] = 0; }
            store_row(rows, j, W, r);
        }
    }
}
// Every long missing row is dropped and every piece gets a one-base missing row in front, so row j simply moves to
// j + 1 and a long row's place is taken by the next piece's leading row: no compaction, only the piece offsets need a scan.
__global__ void k_break_long_spans(const int32_t *rows, int64_t n, Params P, int64_t cutoff, int32_t *out, int64_t *flag)
{
    const int W = P.W;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j <= n; j += (int64_t)gridDim.x * blockDim.x) {
        Row r;
        bool lng = j == 0;                       // slot 0: the first piece's leading row
        if (j > 0) {
            r = load_row(rows, j - 1, W);
            lng = r.v[0] >= cutoff;
            for (int p = 0; p < P.npop; ++p) lng = lng && r.v[1 + 3 * p] == -1 && r.v[3 + 3 * p] == 0;
        }
        if (lng) {
            for (int c = 0; c < W; ++c) r.v[c] = 0;
            r.v[0] = 1;
            for (int p = 0; p < P.npop; ++p) r.v[1 + 3 * p] = -1;
        }
        store_row(out, j, W, r);
        flag[j] = lng;
    }
}
__global__ void k_piece_offsets(const int64_t *flag_scan, const int64_t *flag, int64_t n1, int64_t n_pieces, int64_t *off)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n1; j += (int64_t)gridDim.x * blockDim.x)
        if (flag[j]) off[flag_scan[j]] = j;
    if (blockIdx.x == 0 && threadIdx.x == 0) off[n_pieces] = n1;
}

int grid_for(int64_t n)
{
    int64_t b = (n + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148 * 16) b = 148 * 16;
    return (int)b;
}

}  // namespace

struct smcpp_b200_obs {
    int device = 0;
    cudaStream_t st = nullptr;
    std::string err;
    int npop = 0, W = 0;
    int64_t n = 0;                     // rows currently held
    Buf<int32_t> rows, rows2;          // current rows / output of the running step (swapped)
    Buf<int32_t> pieces;               // result of break_long_spans (all pieces, one after the other)
    Buf<int64_t> piece_off;            // [n_pieces + 1] row offsets into `pieces`
    int64_t n_pieces = 0, n_piece_rows = 0;
    Buf<int64_t> a64, b64, c64, part;  // spans / counts / flags, their scans, run starts; chunk partition
    Buf<unsigned char> cub_tmp;
    int64_t *h_pin = nullptr;          // 2 pinned int64 for scalar read-backs
    float last_ms = 0.f;               // device time of the last step (CUDA events)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

static std::string g_obs_create_error;

#define OCU(call)                                                                \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) {                                                 \
            o->err = std::string(#call) + ": " + cudaGetErrorString(e_);         \
            return 1;                                                            \
        }                                                                        \
    } while (0)

// exclusive prefix sum of `in` into `outp` (n elements) + the grand total to *total
static int scan_total(smcpp_b200_obs *o, const int64_t *in, int64_t *outp, int64_t n, int64_t *total)
{
    size_t bytes = 0;
    OCU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, outp, (int)n, o->st));
    OCU(o->cub_tmp.ensure(bytes));
    OCU(cub::DeviceScan::ExclusiveSum(o->cub_tmp.p, bytes, in, outp, (int)n, o->st));
    // total = last scan value + last input value
    OCU(cudaMemcpyAsync(&o->h_pin[0], outp + (n - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, o->st));
    OCU(cudaMemcpyAsync(&o->h_pin[1], in + (n - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, o->st));
    OCU(cudaStreamSynchronize(o->st));
    *total = o->h_pin[0] + o->h_pin[1];
    return 0;
}

// positions of the current rows into a64 (spans) / b64 (exclusive scan); returns the total number of bases
static int positions(smcpp_b200_obs *o, int64_t *total)
{
    OCU(o->a64.ensure(o->n));
    OCU(o->b64.ensure(o->n));
    k_obs_spans<<<grid_for(o->n), 256, 0, o->st>>>(o->rows.p, o->n, o->W, o->a64.p);
    return scan_total(o, o->a64.p, o->b64.p, o->n, total);
}

static void swap_rows(smcpp_b200_obs *o)
{
    Buf<int32_t> t = o->rows;
    o->rows = o->rows2;
    o->rows2 = t;
}

static Params make_params(const smcpp_b200_obs *o, const int64_t *a)
{
    Params P;
    P.W = o->W;
    P.npop = o->npop;
    P.a[0] = a ? a[0] : 0;
    P.a[1] = (a && o->npop > 1) ? a[1] : 0;
    return P;
}

extern "C" {

int smcpp_b200_obs_create(smcpp_b200_obs **out, int device)
{
    if (!out) return 1;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        g_obs_create_error = "no usable CUDA device; smcpp_b200 has no CPU fallback";
        return 1;
    }
    smcpp_b200_obs *o = new smcpp_b200_obs();
    o->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&o->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost(&o->h_pin, 2 * sizeof(int64_t)) != cudaSuccess) {
        g_obs_create_error = "cuda init failed";
        delete o;
        return 1;
    }
    cudaEventCreate(&o->e0);
    cudaEventCreate(&o->e1);
    *out = o;
    return 0;
}

void smcpp_b200_obs_destroy(smcpp_b200_obs *o)
{
    if (!o) return;
    cudaSetDevice(o->device);
    cudaStreamSynchronize(o->st);
    o->rows.release(); o->rows2.release(); o->pieces.release(); o->piece_off.release(); o->a64.release(); o->b64.release(); o->c64.release(); o->part.release(); o->cub_tmp.release();
    if (o->h_pin) cudaFreeHost(o->h_pin);
    if (o->e0) cudaEventDestroy(o->e0);
    if (o->e1) cudaEventDestroy(o->e1);
    if (o->st) cudaStreamDestroy(o->st);
    delete o;
}

const char *smcpp_b200_obs_last_error(const smcpp_b200_obs *o) { return o ? o->err.c_str() : g_obs_create_error.c_str(); }

int smcpp_b200_obs_upload(smcpp_b200_obs *o, const int32_t *rows, int64_t n_rows, int npop)
{
    if (!o) return 1;
    if (!rows || n_rows <= 0 || npop < 1 || npop > 2) { o->err = "obs_upload: need rows, n_rows > 0 and npop in {1, 2}"; return 1; }
    if (n_rows > 0x7fffffffLL) { o->err = "obs_upload: more than 2^31 - 1 rows in one contig"; return 1; }
    OCU(cudaSetDevice(o->device));
    o->npop = npop;
    o->W = 1 + 3 * npop;
    o->n = n_rows;
    OCU(o->rows.ensure((size_t)n_rows * o->W));
    OCU(cudaMemcpyAsync(o->rows.p, rows, (size_t)n_rows * o->W * sizeof(int32_t), cudaMemcpyHostToDevice, o->st));
    OCU(cudaStreamSynchronize(o->st));
    return 0;
}

int64_t smcpp_b200_obs_rows(const smcpp_b200_obs *o) { return o ? o->n : -1; }
float smcpp_b200_obs_last_ms(const smcpp_b200_obs *o) { return o ? o->last_ms : -1.f; }

int smcpp_b200_obs_download(smcpp_b200_obs *o, int32_t *rows)
{
    if (!o || !rows || o->n <= 0) return 1;
    OCU(cudaSetDevice(o->device));
    OCU(cudaMemcpyAsync(rows, o->rows.p, (size_t)o->n * o->W * sizeof(int32_t), cudaMemcpyDeviceToHost, o->st));
    OCU(cudaStreamSynchronize(o->st));
    return 0;
}

int smcpp_b200_obs_thin(smcpp_b200_obs *o, int thinning, int offset)
{
    if (!o || o->n <= 0) return 1;
    if (thinning < 1 || offset < 0) { o->err = "obs_thin: thinning >= 1 and offset >= 0 required"; return 1; }
    OCU(cudaSetDevice(o->device));
    cudaEventRecord(o->e0, o->st);
    int64_t total = 0, n_out = 0;
    if (positions(o, &total)) return 1;
    OCU(o->c64.ensure(o->n));
    k_thin_count<<<grid_for(o->n), 256, 0, o->st>>>(o->rows.p, o->b64.p, o->n, o->W, thinning, offset, o->a64.p);
    if (scan_total(o, o->a64.p, o->c64.p, o->n, &n_out)) return 1;
    if (n_out > 0x7fffffffLL) { o->err = "obs_thin: result exceeds 2^31 - 1 rows"; return 1; }
    OCU(o->rows2.ensure((size_t)n_out * o->W));
    {
        const int64_t n_chunks = (n_out + kChunk - 1) / kChunk;
        OCU(o->part.ensure(n_chunks + 1));
        k_partition<<<grid_for(n_chunks + 1), 256, 0, o->st>>>(o->c64.p, o->n, n_chunks, kChunk, o->part.p);
        k_thin_write<<<(int)std::min<int64_t>(n_chunks, 148 * 32), kChunk, 0, o->st>>>(o->rows.p, o->b64.p, o->c64.p, o->part.p, o->n, n_out,
                                                                                       make_params(o, nullptr), thinning, offset, o->rows2.p);
    }
    cudaEventRecord(o->e1, o->st);
    OCU(cudaStreamSynchronize(o->st));
    OCU(cudaGetLastError());
    cudaEventElapsedTime(&o->last_ms, o->e0, o->e1);
    swap_rows(o);
    o->n = n_out;
    return 0;
}

int smcpp_b200_obs_bin(smcpp_b200_obs *o, const int64_t *a, int64_t w)
{
    if (!o || o->n <= 0) return 1;
    if (!a || w < 1) { o->err = "obs_bin: a[npop] and w >= 1 required"; return 1; }
    OCU(cudaSetDevice(o->device));
    cudaEventRecord(o->e0, o->st);
    int64_t total = 0;
    if (positions(o, &total)) return 1;
    const int64_t n_bins = (total + w - 1) / w;
    OCU(o->rows2.ensure((size_t)n_bins * o->W));
    {
        const int64_t n_chunks = (n_bins + kChunk - 1) / kChunk;
        OCU(o->part.ensure(n_chunks + 1));
        k_partition<<<grid_for(n_chunks + 1), 256, 0, o->st>>>(o->b64.p, o->n, n_chunks, (int64_t)kChunk * w, o->part.p);
        k_bin<<<(int)std::min<int64_t>(n_chunks, 148 * 32), kChunk, 0, o->st>>>(o->rows.p, o->b64.p, o->part.p, o->n, total, n_bins,
                                                                                make_params(o, a), w, o->rows2.p);
    }
    cudaEventRecord(o->e1, o->st);
    OCU(cudaStreamSynchronize(o->st));
    OCU(cudaGetLastError());
    cudaEventElapsedTime(&o->last_ms, o->e0, o->e1);
    swap_rows(o);
    o->n = n_bins;
    return 0;
}

int smcpp_b200_obs_recode_monomorphic(smcpp_b200_obs *o, const int64_t *a)
{
    if (!o || o->n <= 0 || !a) return 1;
    OCU(cudaSetDevice(o->device));
    cudaEventRecord(o->e0, o->st);
    k_recode_monomorphic<<<grid_for(o->n), 256, 0, o->st>>>(o->rows.p, o->n, make_params(o, a));
    cudaEventRecord(o->e1, o->st);
    OCU(cudaStreamSynchronize(o->st));
    OCU(cudaGetLastError());
    cudaEventElapsedTime(&o->last_ms, o->e0, o->e1);
    return 0;
}

int smcpp_b200_obs_recode_nonseg(smcpp_b200_obs *o, int64_t cutoff)
{
    if (!o || o->n <= 0) return 1;
    if (cutoff < 0) { o->err = "obs_recode_nonseg: cutoff >= 0 required"; return 1; }
    OCU(cudaSetDevice(o->device));
    cudaEventRecord(o->e0, o->st);
    k_recode_nonseg<<<grid_for(o->n), 256, 0, o->st>>>(o->rows.p, o->n, make_params(o, nullptr), cutoff);
    cudaEventRecord(o->e1, o->st);
    OCU(cudaStreamSynchronize(o->st));
    OCU(cudaGetLastError());
    cudaEventElapsedTime(&o->last_ms, o->e0, o->e1);
    return 0;
}

int smcpp_b200_obs_break_long_spans(smcpp_b200_obs *o, int64_t span_cutoff, int64_t *n_pieces)
{
    if (!o || o->n <= 0 || !n_pieces) return 1;
    OCU(cudaSetDevice(o->device));
    const int64_t n1 = o->n + 1;
    OCU(o->pieces.ensure((size_t)n1 * o->W));
    OCU(o->a64.ensure(n1));
    OCU(o->c64.ensure(n1));
    cudaEventRecord(o->e0, o->st);
    k_break_long_spans<<<grid_for(n1), 256, 0, o->st>>>(o->rows.p, o->n, make_params(o, nullptr), span_cutoff, o->pieces.p, o->a64.p);
    int64_t np = 0;
    if (scan_total(o, o->a64.p, o->c64.p, n1, &np)) return 1;
    OCU(o->piece_off.ensure(np + 1));
    k_piece_offsets<<<grid_for(n1), 256, 0, o->st>>>(o->c64.p, o->a64.p, n1, np, o->piece_off.p);
    cudaEventRecord(o->e1, o->st);
    OCU(cudaStreamSynchronize(o->st));
    OCU(cudaGetLastError());
    cudaEventElapsedTime(&o->last_ms, o->e0, o->e1);
    o->n_pieces = np;
    o->n_piece_rows = n1;
    *n_pieces = np;
    return 0;
}

int smcpp_b200_obs_piece_offsets(smcpp_b200_obs *o, int64_t *offsets)
{
    if (!o || !offsets || o->n_pieces <= 0) { if (o) o->err = "obs_piece_offsets: break_long_spans has not been run"; return 1; }
    OCU(cudaSetDevice(o->device));
    OCU(cudaMemcpyAsync(offsets, o->piece_off.p, (size_t)(o->n_pieces + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, o->st));
    OCU(cudaStreamSynchronize(o->st));
    return 0;
}

int smcpp_b200_obs_select_piece(smcpp_b200_obs *o, int64_t piece, int64_t row_begin, int64_t row_end)
{
    if (!o || o->n_pieces <= 0) { if (o) o->err = "obs_select_piece: break_long_spans has not been run"; return 1; }
    if (piece < 0 || piece >= o->n_pieces || row_begin < 0 || row_end > o->n_piece_rows || row_begin >= row_end) {
        o->err = "obs_select_piece: piece / row range out of bounds";
        return 1;
    }
    OCU(cudaSetDevice(o->device));
    const int64_t n = row_end - row_begin;
    OCU(o->rows.ensure((size_t)n * o->W));
    OCU(cudaMemcpyAsync(o->rows.p, o->pieces.p + (size_t)row_begin * o->W, (size_t)n * o->W * sizeof(int32_t), cudaMemcpyDeviceToDevice, o->st));
    OCU(cudaStreamSynchronize(o->st));
    o->n = n;
    return 0;
}

int smcpp_b200_obs_compress(smcpp_b200_obs *o)
{
    if (!o || o->n <= 0) return 1;
    OCU(cudaSetDevice(o->device));
    cudaEventRecord(o->e0, o->st);
    int64_t total = 0, n_runs = 0;
    if (positions(o, &total)) return 1;            // b64 = positions
    OCU(o->c64.ensure(o->n + 1));
    k_run_heads<<<grid_for(o->n), 256, 0, o->st>>>(o->rows.p, o->n, o->W, o->a64.p);
    if (scan_total(o, o->a64.p, o->c64.p, o->n, &n_runs)) return 1;   // c64 = run index of every row
    // run starts reuse a64 (the flags are recomputed instead of kept: 8 B per row less traffic than a third array)
    k_run_starts<<<grid_for(o->n), 256, 0, o->st>>>(o->c64.p, o->rows.p, o->n, o->W, o->a64.p);
    OCU(o->rows2.ensure((size_t)n_runs * o->W));
    k_compress_write<<<grid_for(n_runs), 256, 0, o->st>>>(o->rows.p, o->b64.p, o->a64.p, o->n, total, n_runs, o->W, o->rows2.p);
    cudaEventRecord(o->e1, o->st);
    OCU(cudaStreamSynchronize(o->st));
    OCU(cudaGetLastError());
    cudaEventElapsedTime(&o->last_ms, o->e0, o->e1);
    swap_rows(o);
    o->n = n_runs;
    return 0;
}

}  // extern "C"

// smcpp_b200 -- several GPUs inside ONE process (the drop-in case: `smc++ estimate` is a single Python process, it cannot be
// relaunched under torchrun).  A multi handle owns one single-device context per GPU, shards the contigs over them
// (contigs are independent HMMs, reference src/inference_manager.cpp:89-94), drives every device from its own host thread
// and sums the packed statistics [ll | gamma0 | xisum | gamma_sums] with ONE ncclAllReduce(ncclDouble, ncclSum) per
// E-step over NVLink, in place in the contexts' device buffers.
//
// This layer sits on top of the single-device C ABI (include/smcpp_b200.h) and needs nothing else from the library.
// NCCL is bound at run time (dlopen "libnccl.so.2" -- the copy the process already has, e.g. PyTorch's, or the system's):
// the library itself has no link-time dependency on it and a single GPU never loads it.
#include "../../include/smcpp_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

namespace {

// ---- the few NCCL entry points we use, resolved lazily ---------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;     // ncclSuccess == 0
struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load()
    {
        if (lib) return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { error = std::string("NCCL is not available: ") + dlerror(); return false; }
#define SYM(field, name)                                                                       \
    field = reinterpret_cast<decltype(field)>(dlsym(lib, name));                               \
    if (!field) { error = std::string("NCCL symbol missing: ") + name; lib = nullptr; return false; }
        SYM(CommInitAll, "ncclCommInitAll");
        SYM(CommDestroy, "ncclCommDestroy");
        SYM(AllReduce, "ncclAllReduce");
        SYM(GroupStart, "ncclGroupStart");
        SYM(GroupEnd, "ncclGroupEnd");
        SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
        return true;
    }
};
Nccl g_nccl;
constexpr int kNcclDouble = 8, kNcclSum = 0;     // ncclFloat64, ncclSum (nccl.h; stable since NCCL 2.0)

struct KeyRow {
    std::array<int32_t, 6> v;
    bool operator==(const KeyRow &o) const { return v == o.v; }
};
struct KeyRowHash {
    size_t operator()(const KeyRow &k) const
    {
        uint64_t h = 1469598103934665603ull;
        for (int32_t x : k.v) { h ^= (uint32_t)x; h *= 1099511628211ull; }
        return (size_t)h;
    }
};

thread_local std::string g_multi_create_error;

}  // namespace

struct smcpp_b200_multi {
    std::vector<int> devices;
    std::vector<smcpp_b200_ctx *> ctx;
    std::vector<ncclComm_t> comms;
    std::vector<std::vector<int>> shard;       // contig indices per device (ascending)
    std::vector<int32_t> keys;                 // K x 3P, the reference's std::map order
    int C = 0, npop = 0, K = 0;
    std::string err;
};

static int mfail(smcpp_b200_multi *m, const std::string &msg)
{
    m->err = msg;
    return 1;
}

extern "C" {

int smcpp_b200_multi_create(smcpp_b200_multi **out, const int *devices, int n_devices)
{
    if (!out) return 1;
    *out = nullptr;
    if (!devices || n_devices < 1) { g_multi_create_error = "multi_create: no devices"; return 1; }
    smcpp_b200_multi *m = new smcpp_b200_multi();
    m->devices.assign(devices, devices + n_devices);
    for (int i = 0; i < n_devices; ++i) {
        smcpp_b200_ctx *c = nullptr;
        if (smcpp_b200_create(&c, devices[i])) {
            g_multi_create_error = std::string("multi_create: device ") + std::to_string(devices[i]) + ": " + smcpp_b200_last_error(nullptr);
            for (auto *x : m->ctx) smcpp_b200_destroy(x);
            delete m;
            return 1;
        }
        m->ctx.push_back(c);
    }
    if (n_devices > 1) {
        // one communicator per device, single process (ncclCommInitAll)
        if (!g_nccl.load()) {
            g_multi_create_error = "multi_create: " + g_nccl.error;
            for (auto *x : m->ctx) smcpp_b200_destroy(x);
            delete m;
            return 1;
        }
        m->comms.assign(n_devices, nullptr);
        const ncclResult_t rc = g_nccl.CommInitAll(m->comms.data(), n_devices, devices);
        if (rc != 0) {
            g_multi_create_error = std::string("multi_create: ncclCommInitAll: ") + g_nccl.GetErrorString(rc);
            for (auto *x : m->ctx) smcpp_b200_destroy(x);
            delete m;
            return 1;
        }
    }
    *out = m;
    return 0;
}

void smcpp_b200_multi_destroy(smcpp_b200_multi *m)
{
    if (!m) return;
    for (auto c : m->comms)
        if (c) g_nccl.CommDestroy(c);
    for (auto *x : m->ctx) smcpp_b200_destroy(x);
    delete m;
}

const char *smcpp_b200_multi_last_error(const smcpp_b200_multi *m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

int smcpp_b200_multi_num_devices(const smcpp_b200_multi *m) { return m ? (int)m->ctx.size() : -1; }

int smcpp_b200_multi_context(smcpp_b200_multi *m, int i, smcpp_b200_ctx **ctx)
{
    if (!m || !ctx || i < 0 || i >= (int)m->ctx.size()) return 1;
    *ctx = m->ctx[i];
    return 0;
}

int smcpp_b200_multi_set_contigs(smcpp_b200_multi *m, int n_contigs, const int32_t *const *obs, const int32_t *lengths, int npop)
{
    if (!m) return 1;
    if (n_contigs <= 0 || !obs || !lengths) return mfail(m, "multi_set_contigs: no contigs");
    if (npop < 1 || npop > 2) return mfail(m, "multi_set_contigs: npop must be 1 or 2");
    const int W = 1 + 3 * npop, Q = 3 * npop, D = (int)m->ctx.size();
    // global key table = union over all contigs, lexicographic (reference include/block_key.h:51-60): every device packs
    // gamma_sums identically, so the all-reduce adds like to like
    std::unordered_set<KeyRow, KeyRowHash> seen;
    for (int c = 0; c < n_contigs; ++c) {
        KeyRow last{};
        bool have = false;
        for (int64_t l = 0; l < lengths[c]; ++l) {
            const int32_t *row = obs[c] + l * W;
            KeyRow kr{};
            for (int q = 0; q < Q; ++q) kr.v[q] = row[1 + q];
            if (!have || !(kr == last)) { seen.insert(kr); last = kr; have = true; }
        }
    }
    std::vector<KeyRow> table(seen.begin(), seen.end());
    std::sort(table.begin(), table.end(), [Q](const KeyRow &a, const KeyRow &b) {
        return std::lexicographical_compare(a.v.begin(), a.v.begin() + Q, b.v.begin(), b.v.begin() + Q);
    });
    m->K = (int)table.size();
    m->keys.assign((size_t)m->K * Q, 0);
    for (int k = 0; k < m->K; ++k)
        for (int q = 0; q < Q; ++q) m->keys[(size_t)k * Q + q] = table[k].v[q];
    // longest-processing-time sharding (ties: lower contig index, lower device)
    std::vector<int> order(n_contigs);
    for (int c = 0; c < n_contigs; ++c) order[c] = c;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lengths[a] > lengths[b]; });
    std::vector<int64_t> load(D, 0);
    m->shard.assign(D, {});
    for (int c : order) {
        int best = 0;
        for (int d = 1; d < D; ++d)
            if (load[d] < load[best]) best = d;
        m->shard[best].push_back(c);
        load[best] += lengths[c];
    }
    for (auto &s : m->shard) std::sort(s.begin(), s.end());
    m->C = n_contigs;
    m->npop = npop;
    // uploads run side by side (set_contigs is host-heavy: key encoding, span sorting)
    std::vector<int> rc(D, 0);
    std::vector<std::thread> th;
    for (int d = 0; d < D; ++d)
        th.emplace_back([&, d]() {
            if (m->shard[d].empty()) return;
            std::vector<const int32_t *> o;
            std::vector<int32_t> len;
            for (int c : m->shard[d]) { o.push_back(obs[c]); len.push_back(lengths[c]); }
            rc[d] = smcpp_b200_set_contigs(m->ctx[d], (int)o.size(), o.data(), len.data(), npop, m->keys.data(), m->K);
        });
    for (auto &t : th) t.join();
    for (int d = 0; d < D; ++d)
        if (rc[d]) return mfail(m, std::string("device ") + std::to_string(m->devices[d]) + ": " + smcpp_b200_last_error(m->ctx[d]));
    return 0;
}

int smcpp_b200_multi_num_keys(const smcpp_b200_multi *m) { return m ? m->K : -1; }
int smcpp_b200_multi_get_keys(const smcpp_b200_multi *m, int32_t *keys)
{
    if (!m || !keys) return 1;
    std::memcpy(keys, m->keys.data(), m->keys.size() * sizeof(int32_t));
    return 0;
}
int smcpp_b200_multi_get_shard(const smcpp_b200_multi *m, int device_index, int32_t *contigs, int32_t *n)
{
    if (!m || device_index < 0 || device_index >= (int)m->shard.size() || !n) return 1;
    *n = (int32_t)m->shard[device_index].size();
    if (contigs) std::memcpy(contigs, m->shard[device_index].data(), m->shard[device_index].size() * sizeof(int32_t));
    return 0;
}

int smcpp_b200_multi_estep(smcpp_b200_multi *m, int M, const double *pi, const double *T, const double *E, double *ll,
                           double *xisum, double *gamma0, double *gamma_sums, uint8_t *key_present, double *reduced)
{
    if (!m) return 1;
    if (m->C == 0) return mfail(m, "multi_estep: multi_set_contigs() has not been called");
    const int D = (int)m->ctx.size(), K = m->K;
    const size_t MM = (size_t)M * M, nred = 1 + M + MM + (size_t)K * M;
    // ---- every device: eigensystems, kernels, repair if needed (one host thread per device; returns synchronised)
    std::vector<int> rc(D, 0);
    {
        std::vector<std::thread> th;
        for (int d = 0; d < D; ++d)
            th.emplace_back([&, d]() {
                if (m->shard[d].empty()) return;
                rc[d] = smcpp_b200_estep_device(m->ctx[d], M, pi, T, E, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 1);
            });
        for (auto &t : th) t.join();
    }
    for (int d = 0; d < D; ++d)
        if (rc[d]) return mfail(m, std::string("device ") + std::to_string(m->devices[d]) + ": " + smcpp_b200_last_error(m->ctx[d]));
    // ---- one all-reduce of the packed statistics, in place in the device buffers
    std::vector<void *> red(D, nullptr);
    std::vector<void *> stream(D, nullptr);
    int root = -1;
    for (int d = 0; d < D; ++d) {
        if (m->shard[d].empty()) continue;
        int64_t cnt = 0;
        if (smcpp_b200_reduced_device_ptr(m->ctx[d], &red[d], &cnt) || (size_t)cnt != nred || smcpp_b200_stream(m->ctx[d], &stream[d]))
            return mfail(m, "multi_estep: reduced buffer mismatch");
        if (root < 0) root = d;
    }
    if (D > 1) {
        for (int d = 0; d < D; ++d)
            if (m->shard[d].empty()) return mfail(m, "multi_estep: fewer contigs than devices (create the handle with fewer devices)");
        ncclResult_t r = g_nccl.GroupStart();
        for (int d = 0; d < D && r == 0; ++d) {
            cudaSetDevice(m->devices[d]);
            r = g_nccl.AllReduce(red[d], red[d], nred, kNcclDouble, kNcclSum, m->comms[d], (cudaStream_t)stream[d]);
        }
        const ncclResult_t r2 = g_nccl.GroupEnd();
        if (r != 0 || r2 != 0) return mfail(m, std::string("multi_estep: ncclAllReduce: ") + g_nccl.GetErrorString(r != 0 ? r : r2));
    }
    // ---- per-contig results back in the caller's contig order (the Python API is per HMM), reduced from the first device
    std::vector<double> b_ll, b_x, b_g0, b_gs;
    std::vector<uint8_t> b_kp;
    for (int d = 0; d < D; ++d) {
        const size_t n = m->shard[d].size();
        if (!n) continue;
        b_ll.resize(n); b_x.resize(n * MM); b_g0.resize(n * M); b_gs.resize(n * K * M); b_kp.resize(n * K);
        if (smcpp_b200_fetch(m->ctx[d], ll ? b_ll.data() : nullptr, xisum ? b_x.data() : nullptr, gamma0 ? b_g0.data() : nullptr,
                             gamma_sums ? b_gs.data() : nullptr, (d == root && reduced) ? reduced : nullptr))
            return mfail(m, std::string("device ") + std::to_string(m->devices[d]) + ": " + smcpp_b200_last_error(m->ctx[d]));
        if (key_present && smcpp_b200_get_key_present(m->ctx[d], b_kp.data())) return mfail(m, "multi_estep: key_present");
        for (size_t i = 0; i < n; ++i) {
            const size_t c = (size_t)m->shard[d][i];
            if (ll) ll[c] = b_ll[i];
            if (xisum) std::memcpy(xisum + c * MM, b_x.data() + i * MM, MM * sizeof(double));
            if (gamma0) std::memcpy(gamma0 + c * M, b_g0.data() + i * M, (size_t)M * sizeof(double));
            if (gamma_sums) std::memcpy(gamma_sums + c * K * M, b_gs.data() + i * K * M, (size_t)K * M * sizeof(double));
            if (key_present) std::memcpy(key_present + c * K, b_kp.data() + i * K, K);
        }
    }
    return 0;
}

}  // extern "C"

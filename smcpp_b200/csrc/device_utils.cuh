// smcpp_b200 -- small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include <atomic>

namespace smcb {

// Kernel attributes (dynamic shared-memory limit, carve-out) are set per DEVICE, and one process may drive several
// GPUs (one context per device, one host thread each): "already configured" flags are kept per device.
constexpr int kMaxDevices = 64;
inline int current_device_slot()
{
    int d = 0;
    cudaGetDevice(&d);
    return d >= 0 && d < kMaxDevices ? d : 0;
}
// true when this device still needs cudaFuncSetAttribute for `bytes` of dynamic shared memory; records the new size
inline bool needs_smem_config(std::atomic<size_t> (&configured)[kMaxDevices], size_t bytes)
{
    std::atomic<size_t> &c = configured[current_device_slot()];
    if (c.load(std::memory_order_relaxed) >= bytes && bytes > 0) return false;
    c.store(bytes, std::memory_order_relaxed);
    return true;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// d^span for integer span >= 1 through the precomputed log|d| (the reference calls std::pow,
// src/hmm.cpp:75; the relative difference is <= |span log d| * 2^-53).
__device__ __forceinline__ double pow_span(double d, double logd, int span)
{
    if (d == 0.0) return 0.0;
    double r = exp((double)span * logd);
    return (d < 0.0 && (span & 1)) ? -r : r;
}

// Eigen 3.3.3 float sum() order (LinearVectorizedTraversal, SSE packets of 4, two accumulators), which
// is what `alpha_hat.col(ell).sum()` compiles to in the reference (src/hmm.cpp:87).
__device__ __forceinline__ float eigen_sum_f32(const float *v, int M, int astart)
{
    // `astart` = leading coefficients before the first 16-byte aligned one of the reference's column:
    // (4 - (ell*M) % 4) % 4 for column ell of the float matrix alpha_hat (0 whenever M % 4 == 0).
    float r;
    if (astart == 0) {
        const int n4 = M >> 2, n8 = M >> 3;
        if (n4) {
            const float4 *v4 = reinterpret_cast<const float4 *>(v);
            float4 p0 = v4[0];
            if (n4 > 1) {
                float4 p1 = v4[1];
                for (int q = 1; q < n8; ++q) {
                    float4 a = v4[2 * q], b = v4[2 * q + 1];
                    p0.x = __fadd_rn(p0.x, a.x); p0.y = __fadd_rn(p0.y, a.y); p0.z = __fadd_rn(p0.z, a.z); p0.w = __fadd_rn(p0.w, a.w);
                    p1.x = __fadd_rn(p1.x, b.x); p1.y = __fadd_rn(p1.y, b.y); p1.z = __fadd_rn(p1.z, b.z); p1.w = __fadd_rn(p1.w, b.w);
                }
                p0.x = __fadd_rn(p0.x, p1.x); p0.y = __fadd_rn(p0.y, p1.y); p0.z = __fadd_rn(p0.z, p1.z); p0.w = __fadd_rn(p0.w, p1.w);
                if (n4 > 2 * n8) {
                    float4 a = v4[2 * n8];
                    p0.x = __fadd_rn(p0.x, a.x); p0.y = __fadd_rn(p0.y, a.y); p0.z = __fadd_rn(p0.z, a.z); p0.w = __fadd_rn(p0.w, a.w);
                }
            }
            r = __fadd_rn(__fadd_rn(p0.x, p0.z), __fadd_rn(p0.y, p0.w));
            for (int i = n4 * 4; i < M; ++i) r = __fadd_rn(r, v[i]);
        } else {
            r = v[0];
            for (int i = 1; i < M; ++i) r = __fadd_rn(r, v[i]);
        }
        return r;
    }
    if (astart > M) astart = M;
    const int asize = ((M - astart) >> 2) << 2, asize2 = ((M - astart) >> 3) << 3;
    const int aend = astart + asize, aend2 = astart + asize2;
    if (asize) {
        float p0[4], p1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) p0[k] = v[astart + k];
        if (asize > 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) p1[k] = v[astart + 4 + k];
            for (int i = astart + 8; i < aend2; i += 8) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    p0[k] = __fadd_rn(p0[k], v[i + k]);
                    p1[k] = __fadd_rn(p1[k], v[i + 4 + k]);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) p0[k] = __fadd_rn(p0[k], p1[k]);
            if (aend > aend2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) p0[k] = __fadd_rn(p0[k], v[aend2 + k]);
            }
        }
        r = __fadd_rn(__fadd_rn(p0[0], p0[2]), __fadd_rn(p0[1], p0[3]));
        for (int i = 0; i < astart; ++i) r = __fadd_rn(r, v[i]);
        for (int i = aend; i < M; ++i) r = __fadd_rn(r, v[i]);
    } else {
        r = v[0];
        for (int i = 1; i < M; ++i) r = __fadd_rn(r, v[i]);
    }
    return r;
}


// Eigen 3.3.3 double sum() order (SSE packets of 2, two accumulators; the port's sum_f64): what
// `a.sum()` compiles to in the reference (src/hmm.cpp:77).  v in shared memory, every lane gets the same value.
__device__ __forceinline__ double eigen_sum_f64(const double *v, int M)
{
    const int n2 = M >> 1, n4 = M >> 2;
    double r;
    if (n2) {
        const double2 *v2 = reinterpret_cast<const double2 *>(v);
        double2 p0 = v2[0];
        if (n2 > 1) {
            double2 p1 = v2[1];
            for (int q = 1; q < n4; ++q) {
                const double2 a = v2[2 * q], b = v2[2 * q + 1];
                p0.x = __dadd_rn(p0.x, a.x); p0.y = __dadd_rn(p0.y, a.y);
                p1.x = __dadd_rn(p1.x, b.x); p1.y = __dadd_rn(p1.y, b.y);
            }
            p0.x = __dadd_rn(p0.x, p1.x); p0.y = __dadd_rn(p0.y, p1.y);
            if (n2 > 2 * n4) {
                const double2 a = v2[2 * n4];
                p0.x = __dadd_rn(p0.x, a.x); p0.y = __dadd_rn(p0.y, a.y);
            }
        }
        r = __dadd_rn(p0.x, p0.y);
        for (int i = n2 * 2; i < M; ++i) r = __dadd_rn(r, v[i]);
    } else {
        r = v[0];
        for (int i = 1; i < M; ++i) r = __dadd_rn(r, v[i]);
    }
    return r;
}

// exact power-of-two factor 2^-floor(log2|x|) (1.0 when x is zero, subnormal, inf or nan)
__device__ __forceinline__ double pow2_rescale(double x)
{
    const int ex = (__double2hiint(x) >> 20) & 0x7ff;
    if (ex == 0 || ex == 0x7ff) return 1.0;
    return __hiloint2double((2046 - ex) << 20, 0);
}

}  // namespace smcb

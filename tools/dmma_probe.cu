// Probe: FP64 throughput of mma.sync.m8n8k4.f64 (DMMA) vs DFMA on this GPU, and DMMA latency.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dmma(double *sink, int iters)
{
    double c[CHAINS][2];
    for (int i = 0; i < CHAINS; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = 1.0 + i; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double r = 0;
    for (int i = 0; i < CHAINS; ++i) r += c[i][0] + c[i][1];
    if (r == 1.2345) sink[0] = r;
}

__global__ void k_dfma(double *sink, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) sink[0] = r;
}

template <typename F> float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    double *sink; cudaMalloc(&sink, 64);
    const int iters = 1 << 14;
    for (int warps = 1; warps <= 16; warps *= 2) {
        const int ctas = 148 * 4, thr = 32 * warps;   // warps per CTA; 4 CTAs/SM
        float ms = timeit([&] { k_dmma<8><<<ctas, thr>>>(sink, iters); });
        double flop = 2.0 * 256 * 8.0 * iters * (double)ctas * warps;
        printf("DMMA 8 chains  warps/SM=%3d : %.2f TFLOP/s\n", warps * 4, flop / ms / 1e9);
    }
    {
        float ms = timeit([&] { k_dmma<1><<<148, 32>>>(sink, iters); });
        printf("DMMA latency (1 chain, 1 warp/SM): %.1f ns per mma = %.1f cycles @1.9GHz\n", ms * 1e6 / iters, ms * 1e6 / iters * 1.9);
    }
    {
        float ms = timeit([&] { k_dmma<2><<<148, 32>>>(sink, iters); });
        printf("DMMA 2 chains 1 warp/SM: %.1f ns per mma\n", ms * 1e6 / iters / 2);
        ms = timeit([&] { k_dmma<4><<<148, 32>>>(sink, iters); });
        printf("DMMA 4 chains 1 warp/SM: %.1f ns per mma\n", ms * 1e6 / iters / 4);
        ms = timeit([&] { k_dmma<8><<<148, 32>>>(sink, iters); });
        printf("DMMA 8 chains 1 warp/SM: %.1f ns per mma\n", ms * 1e6 / iters / 8);
        ms = timeit([&] { k_dmma<8><<<148, 128>>>(sink, iters); });
        printf("DMMA 8 chains 4 warps/SM (1/SMSP): %.2f TFLOP/s\n", 2.0 * 256 * 8.0 * iters * 148 * 4 / ms / 1e9);
    }
    {
        float ms = timeit([&] { k_dfma<<<148 * 8, 256>>>(sink, iters); });
        printf("DFMA peak: %.2f TFLOP/s\n", 2.0 * 8.0 * iters * 256.0 * 148 * 8 / ms / 1e9);
        ms = timeit([&] { k_dfma<<<148, 32>>>(sink, iters); });
        printf("DFMA 8 chains 1 warp/SM: %.2f ns per warp-fma (%.1f cycles)\n", ms * 1e6 / iters / 8, ms * 1e6 / iters / 8 * 1.9);
    }
    return 0;
}

// Microbenchmark: MUFU.EX2 throughput per SM sub-partition and the cost of a softmax-like chunk (32 FFMA + 32 EX2 + 32 FADD + 16 F2FP).
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t *>(&v); }
__global__ void __launch_bounds__(512) k(long long *out, float *sink, int nwarps, int iters, int mode, float seed)
{
    const int warp = threadIdx.x >> 5;
    long long t0 = 0, t1 = 0;
    float v[32];
    for (int j = 0; j < 32; j++) v[j] = seed * (float)(j + threadIdx.x);
    float s0 = 0.f, s1 = 0.f;
    uint32_t wacc = 0;
    if (warp < nwarps)
    {
        t0 = clock64();
        for (int i = 0; i < iters; i++)
        {
            if (mode == 0)
            { // MUFU only: 32 independent ex2 per iteration
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = ex2(v[j]);
            }
            else if (mode == 1)
            { // softmax chunk: ffma, ex2, add, pack
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; j++) x[j] = ex2(fmaf(v[j], 0.18f, -seed));
#pragma unroll
                for (int j = 0; j < 16; j++)
                {
                    s0 += x[2 * j], s1 += x[2 * j + 1];
                    wacc ^= pack(x[2 * j], x[2 * j + 1]);
                }
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] += 1.0f;
            }
            else
            { // the same without the exponentials (issue cost of everything else)
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; j++) x[j] = fmaf(v[j], 0.18f, -seed);
#pragma unroll
                for (int j = 0; j < 16; j++)
                {
                    s0 += x[2 * j], s1 += x[2 * j + 1];
                    wacc ^= pack(x[2 * j], x[2 * j + 1]);
                }
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] += 1.0f;
            }
        }
        t1 = clock64();
    }
    float acc = s0 + s1 + __uint_as_float(wacc);
    for (int j = 0; j < 32; j++) acc += v[j];
    if (acc == 12345.678f) sink[threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0 && warp < nwarps) out[warp] = t1 - t0;
}
int main()
{
    long long *d, h[16];
    float *sink;
    cudaMalloc(&d, 16 * 8);
    cudaMalloc(&sink, 4096);
    const int iters = 2000;
    const char *names[3] = {"32 EX2 only", "32 FFMA + 32 EX2 + 32 FADD + 16 F2FP + 32 FADD", "the same without EX2"};
    for (int mode = 0; mode < 3; mode++)
        for (int nw : {1, 4, 8, 12, 16})
        {
            k<<<1, 512>>>(d, sink, nw, iters, mode, 0.001f);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
            cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < nw; i++) mx = h[i] > mx ? h[i] : mx;
            printf("%-50s warps %2d (%d per sub-partition): %7.1f cycles per 32-element chunk and warp\n", names[mode], nw, (nw + 3) / 4, (double)mx / iters);
        }
    return 0;
}

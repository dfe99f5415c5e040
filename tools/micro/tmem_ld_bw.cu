// Microbenchmark: tcgen05.ld throughput per SM (bytes per cycle) as a function of the number of warps loading.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu && ./tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr)
                 : "memory");
}
__global__ void __launch_bounds__(512) k(long long *out, int nwarps, int iters, int mode)
{
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps)
    {
        uint32_t v[32], u[32];
        t0 = clock64();
        for (int i = 0; i < iters; i++)
        {
            if (mode == 0)
            { // one load in flight, consumed (max-like dependency on 1 value)
                tmem_ld_32x32(base + (i & 7) * 32, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += v[0] + v[13] + v[31];
            }
            else
            { // two loads in flight
                tmem_ld_32x32(base + (i & 7) * 32, v);
                tmem_ld_32x32(base + 256 + (i & 7) * 32, u);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += v[0] + v[13] + v[31] + u[0] + u[17] + u[31];
            }
        }
        t1 = clock64();
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && warp < nwarps) out[blockIdx.x * 16 + warp] = t1 - t0 + (acc == 0x12345 ? 1 : 0);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512u) : "memory");
}
int main()
{
    long long *d, h[16];
    cudaMalloc(&d, 148 * 16 * 8);
    const int iters = 2000;
    for (int mode = 0; mode < 2; mode++)
        for (int nw : {1, 2, 4, 8, 12, 16})
        {
            cudaMemset(d, 0, 148 * 16 * 8);
            k<<<1, 512>>>(d, nw, iters, mode);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < nw; i++) mx = h[i] > mx ? h[i] : mx;
            const double bytes = (double)nw * iters * 4096.0 * (mode ? 2 : 1);
            printf("mode %d (%s) warps %2d: %lld cycles, %.1f B/clk per SM, %.1f B/clk per warp, %.0f cycles per ld\n", mode,
                   mode ? "2 loads in flight" : "1 load in flight ", nw, mx, bytes / mx, bytes / mx / nw, (double)mx / iters / (mode ? 2 : 1));
        }
    return 0;
}

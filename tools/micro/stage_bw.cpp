// Host staging-copy bandwidth (pageable -> page-locked write-combined) of the library's copy pool, against plain memcpy.
// g++ -O2 -std=gnu++14 stage_bw.cpp -o stage_bw -L../../vit-fpga_b200/lib -lnetcuda -Wl,-rpath,'$ORIGIN/../../vit-fpga_b200/lib' -pthread
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime_api.h>
namespace nc { void staging_copy(void *, const void *, size_t); int staging_threads(); }
int main()
{
    const size_t n = 616u << 20;
    std::vector<char> a(n);
    for (size_t i = 0; i < n; i += 4096) a[i] = (char)i;
    void *wc = nullptr, *pin = nullptr;
    if (cudaHostAlloc(&wc, n, cudaHostAllocWriteCombined) != cudaSuccess || cudaHostAlloc(&pin, n, cudaHostAllocDefault) != cudaSuccess) { printf("cudaHostAlloc failed\n"); return 1; }
    auto bw = [&](void *dst, bool pool) {
        double best = 0;
        for (int rep = 0; rep < 4; rep++)
        {
            auto t0 = std::chrono::steady_clock::now();
            if (pool) nc::staging_copy(dst, a.data(), n); else memcpy(dst, a.data(), n);
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            best = n / s / 1e9 > best ? n / s / 1e9 : best;
        }
        return best;
    };
    printf("copy pool, %d threads: pageable -> write-combined pinned %.1f GB/s, -> ordinary pinned %.1f GB/s; one-thread memcpy -> pinned %.1f GB/s\n",
           nc::staging_threads(), bw(wc, true), bw(pin, true), bw(pin, false));
    return 0;
}

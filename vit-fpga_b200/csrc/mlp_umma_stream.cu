// mlp_umma_stream.cu -- whole INT8 MLP forward for up to 128 samples in ONE persistent tcgen05 kernel.
//
// Reference counterpart: the task `network_v1` that walks all layers of the net (src/netFPGA.cpp:250,275), batched.  Between the
// register-resident weight-streaming kernel (mlp_stream.cu, <= 32 samples) and the plain tcgen05 GEMM (hundreds of samples and up),
// config C5 (8 x 4096 x 4096 int8) is still weight streaming -- 128 MiB of weights against <= 512 KB of activations per layer -- but
// the activations no longer fit in registers.  Per layer the split-K GEMM + finalize pair (16 launches in a CUDA graph) spends most
// of its ~17 us on launch boundaries and prologues.  This kernel keeps the stream going across layers instead:
//   * cooperative launch, one CTA per SM; CTA c owns the output-neuron tiles c, c + G, ... (32 neurons each) of EVERY layer;
//   * warp 0 streams this CTA's weight tiles ([32 rows x 128 bytes] per k-block, TMA, 128B swizzle; four k-blocks per ring slot) and
//     never waits for activations: it runs ahead across the grid barrier between layers;
//   * warp 1 loads the activation tiles ([samples x 128 bytes] per k-block, one per ring slot: the box holds the batch rounded up to 8 rows; the MMA
//     reads 128 rows, and whatever stale bytes sit in the rest of the slot only reach accumulator rows that are never stored) of
//     layer l once every output tile of layer l - 1 is stored (a global counter of finished TILES, cumulative over the layers,
//     release / acquire at GPU scope, reset by the last CTA; only the CTAs that load wait, so -- unlike mlp_stream.cu, where every CTA
//     waits at every layer -- one arrival per CTA would let the CTAs without a tile in a narrow layer run ahead of the count);
//   * warps 2 and 3 issue tcgen05.mma kind::i8, M = 128 (samples) x N = 32 (neurons) x K = 32 per instruction -- each takes every
//     other four-k-block weight group of a tile (split K) into its own int32 accumulator in tensor memory (two stages x two
//     accumulators of 32 columns), from ring slots of its own;
//   * warps 4-7 (one per TMEM lane quarter, thread = sample): + bias, ReLU, >> 7, clamp -- the integers of EPI_REQUANT[_RELU] and of the
//     oracle -- 32 bytes of the next layer's activation row per thread (or 32 int32 of the output row after the last layer).
// Integer arithmetic: order-independent, bit-exact.
#include "gemm_tcgen05.cuh"
#include "kernels.h"

#include <algorithm>
#include <cstdlib>

namespace nc
{

constexpr int MU_ISSUERS = 2;                   // MMA-issuing warps: issuer i takes the weight groups i, i + 2, ... of every tile (split K)
constexpr int MU_EPI_WARP0 = 4;                 // first epilogue warp (a multiple of 4: TMEM lane quarters)
constexpr int MU_THREADS = (MU_EPI_WARP0 + 4) * 32;
// Rings.  A timeline of the first form of this kernel (one issuing thread, one k-block per slot) showed the MMA-issuing thread, not
// the memory system, setting the pace: ~650 cycles per 128-byte k-block -- mbarrier waits of ~90 cycles each even when already
// complete, four tcgen05.mma of ~160 cycles issue-to-issue at N = 32 and the commits.  So weights travel four k-blocks per slot, and
// the K range of a tile is split over two issuing threads with an accumulator each (110 -> 70 us at 128 samples on config C5).
// Every issuer owns its slots: a slot has ONE producer and ONE consumer, who visits its phases in order.  (A ring shared by both
// issuers is wrong: a parity wait can only tell a phase from the one before it, and an issuer asking for the phase after next of a
// slot whose next phase -- the other issuer's -- has not landed yet is told "complete" for the one before.  With four issuers on a
// shared ring that hung config C5; with two it would have been a rare wrong answer.)
constexpr int MU_W_GROUP = 4;                   // k-blocks per weight slot
constexpr int MU_W_SLOTS_PER = 2;               // weight slots per issuer: 2 x 4 x [32 rows x 128 B]
constexpr int MU_W_SLOTS = MU_ISSUERS * MU_W_SLOTS_PER;
constexpr int MU_W_TILE_BYTES = 32 * 128;
constexpr int MU_W_SLOT_BYTES = MU_W_GROUP * MU_W_TILE_BYTES;
constexpr int MU_A_SLOTS_PER = 5;               // activation slots per issuer: one k-block each, [128 samples x 128 B] (only the batch's rows are loaded)
constexpr int MU_A_SLOTS = MU_ISSUERS * MU_A_SLOTS_PER;
constexpr int MU_A_TILE_BYTES = 128 * 128;
constexpr int MU_A_SLOT_BYTES = MU_A_TILE_BYTES;
constexpr int MU_OFF_A = MU_W_SLOTS * MU_W_SLOT_BYTES;                 // 64 KB, 1024-byte aligned
constexpr int MU_OFF_BARS = MU_OFF_A + MU_A_SLOTS * MU_A_SLOT_BYTES;   // + 160 KB
constexpr int MU_NUM_BARS = 2 * MU_W_SLOTS + 2 * MU_A_SLOTS + 4;
constexpr int MU_OFF_TMEM_PTR = MU_OFF_BARS + MU_NUM_BARS * 8;
constexpr int MU_SMEM = MU_OFF_TMEM_PTR + 16;
static_assert(MU_SMEM <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
constexpr int MU_TILE_N = 32;
constexpr uint32_t MU_TMEM_COLS = 2 * MU_ISSUERS * 32; // two accumulator stages x one 32-column accumulator per issuer
static_assert((MU_TMEM_COLS & (MU_TMEM_COLS - 1)) == 0 && MU_TMEM_COLS >= 32 && MU_TMEM_COLS <= 512, "TMEM allocations are powers of two");

enum : int
{
    KERR_MU_W_PRODUCER = 31,
    KERR_MU_A_PRODUCER = 32,
    KERR_MU_MMA = 33,
    KERR_MU_EPILOGUE = 34,
    KERR_MU_GRID_BARRIER = 35,
};

struct MlpUmmaMaps
{
    CUtensorMap w[MLP_STREAM_MAX_LAYERS]; // weights of layer l: {fan_in, fan_out}, box {128 B, 32 rows}
    CUtensorMap a[MLP_STREAM_MAX_LAYERS]; // input activations of layer l: {fan_in, batch}, box {128 B, batch rounded up to 8 rows}
};

struct MlpUmmaParams
{
    int n_layers, batch;
    unsigned relu_mask;                     // bit l: ReLU after layer l
    int fan_in[MLP_STREAM_MAX_LAYERS], fan_out[MLP_STREAM_MAX_LAYERS];
    const int32_t *bias[MLP_STREAM_MAX_LAYERS];
    int8_t *act_out[MLP_STREAM_MAX_LAYERS]; // where layer l writes its int8 outputs (null for the last layer)
    long long ld_out[MLP_STREAM_MAX_LAYERS]; // ... and their row pitch in bytes
    int32_t *out;                           // [batch][fan_out of the last layer]
    unsigned *barrier;                      // two zero-initialised counters in device memory (left at zero by every launch)
    int *error_flag;
    long long *debug;                       // optional [n_layers][8] clock64 stamps of CTA 0 (NETCUDA_DEBUG_TIMELINE builds)
    const int8_t *w_tiled[MLP_STREAM_MAX_LAYERS]; // cluster kernel: the weights in 16 KB streaming blocks (launch_retile_i8_weights)
    int a_slot_shift = 14;                  // cluster kernel: log2 of the activation slot size (4, 8 or 16 KB: the batch's rows x 128 bytes, rounded up)
};

__device__ __forceinline__ void mu_red_release_gpu_add(unsigned *p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned mu_ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(MU_THREADS, 1)
mlp_i8_umma_stream_kernel(const __grid_constant__ MlpUmmaMaps maps, const MlpUmmaParams p)
{
    extern __shared__ __align__(1024) uint8_t mu_smem[];
    const uint32_t base = smem_u32(mu_smem);
    if ((base & 1023u) != 0)
    {
        if (threadIdx.x == 0 && p.error_flag) atomicExch(p.error_flag, KERR_SMEM_ALIGN);
        return;
    }
    const uint32_t bars = base + MU_OFF_BARS;
    auto wfull = [&](int s) { return bars + 8u * s; };
    auto wempty = [&](int s) { return bars + 8u * (MU_W_SLOTS + s); };
    auto afull = [&](int s) { return bars + 8u * (2 * MU_W_SLOTS + s); };
    auto aempty = [&](int s) { return bars + 8u * (2 * MU_W_SLOTS + MU_A_SLOTS + s); };
    auto tfull = [&](int a) { return bars + 8u * (2 * MU_W_SLOTS + 2 * MU_A_SLOTS + a); };
    auto tempty = [&](int a) { return bars + 8u * (2 * MU_W_SLOTS + 2 * MU_A_SLOTS + 2 + a); };
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(mu_smem + MU_OFF_TMEM_PTR);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31; // (uniform: the issuers' descriptors stay in uniform registers, see ptx.cuh)
    const int cta = blockIdx.x, grid = gridDim.x;
#ifdef NETCUDA_DEBUG_TIMELINE
    long long *const dbg = cta == 0 ? p.debug : nullptr;
#else
    constexpr long long *dbg = nullptr;
#endif

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < MU_W_SLOTS; s++) mbar_init(wfull(s), 1), mbar_init(wempty(s), 1);
        for (int s = 0; s < MU_A_SLOTS; s++) mbar_init(afull(s), 1), mbar_init(aempty(s), 1);
        for (int a = 0; a < 2; a++) mbar_init(tfull(a), MU_ISSUERS), mbar_init(tempty(a), 4); // one arrive per issuer / per epilogue warp
        fence_barrier_init();
    }
    if (warp == 3)
    {
        tmem_alloc(base + MU_OFF_TMEM_PTR, MU_TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0)
    {
        // ===================== weight producer: never waits for activations =====================
        if (lane == 0)
        {
            for (int l = 0; l < p.n_layers; l++) tma_prefetch_desc(&maps.w[l]);
            uint32_t cnt[MU_ISSUERS] = {}; // weight groups handed to each issuer so far: slot and phase of its next one
            for (int l = 0; l < p.n_layers; l++)
            {
                const int tiles = (p.fan_out[l] + MU_TILE_N - 1) / MU_TILE_N, nkb = (p.fan_in[l] + 127) >> 7;
                for (int tile = cta; tile < tiles; tile += grid)
                    for (int kb = 0, g = 0; kb < nkb; kb += MU_W_GROUP, g++)
                    {
                        const int is = g % MU_ISSUERS;
                        uint32_t c = 0;
#pragma unroll
                        for (int i = 0; i < MU_ISSUERS; i++)
                            if (i == is) c = cnt[i]++;
                        const int s = is * MU_W_SLOTS_PER + (int)(c % MU_W_SLOTS_PER), nb = min(MU_W_GROUP, nkb - kb);
                        mbar_wait(wempty(s), ((c / MU_W_SLOTS_PER) & 1u) ^ 1u, p.error_flag, KERR_MU_W_PRODUCER);
                        mbar_arrive_expect_tx(wfull(s), (uint32_t)(nb * MU_W_TILE_BYTES));
                        for (int i = 0; i < nb; i++)
                            tma_load_2d(base + s * MU_W_SLOT_BYTES + i * MU_W_TILE_BYTES, &maps.w[l], wfull(s), (kb + i) * 128, tile * MU_TILE_N);
                    }
            }
        }
    }
    else if (warp == 1)
    {
        // ===================== activation producer: layer l follows the grid barrier after layer l - 1 =====================
        if (lane == 0)
        {
            for (int l = 0; l < p.n_layers; l++) tma_prefetch_desc(&maps.a[l]);
            uint32_t cnt[MU_ISSUERS] = {}; // activation k-blocks handed to each issuer so far
            unsigned target = 0; // output tiles of all the layers before this one (the counter starts every launch at zero: reset below)
            for (int l = 0; l < p.n_layers; l++)
            {
                const int tiles = (p.fan_out[l] + MU_TILE_N - 1) / MU_TILE_N, nkb = (p.fan_in[l] + 127) >> 7;
                if (l > 0) target += (unsigned)((p.fan_out[l - 1] + MU_TILE_N - 1) / MU_TILE_N);
                if (l > 0 && cta < tiles)
                {
                    const long long t0 = clock64();
                    while (mu_ld_acquire_gpu(p.barrier) < target)
                        if (clock64() - t0 > 4000000000LL)
                        {
                            if (p.error_flag) atomicExch(p.error_flag, KERR_MU_GRID_BARRIER);
                            __threadfence_system();
                            __trap();
                        }
                    // the other CTAs' generic-proxy stores are visible; order this thread's async-proxy (TMA) reads after them
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                }
                if (dbg) dbg[l * 8 + 0] = clock64(); // barrier passed
                for (int tile = cta; tile < tiles; tile += grid)
                    for (int kb = 0; kb < nkb; kb++)
                    {
                        const int is = (kb / MU_W_GROUP) % MU_ISSUERS; // the issuer of the weight group this k-block belongs to
                        uint32_t c = 0;
#pragma unroll
                        for (int i = 0; i < MU_ISSUERS; i++)
                            if (i == is) c = cnt[i]++;
                        const int s = is * MU_A_SLOTS_PER + (int)(c % MU_A_SLOTS_PER);
                        mbar_wait(aempty(s), ((c / MU_A_SLOTS_PER) & 1u) ^ 1u, p.error_flag, KERR_MU_A_PRODUCER);
                        mbar_arrive_expect_tx(afull(s), (uint32_t)(((p.batch + 7) & ~7) * 128));
                        tma_load_2d(base + MU_OFF_A + s * MU_A_SLOT_BYTES, &maps.a[l], afull(s), kb * 128, 0);
                    }
            }
        }
    }
    else if (warp == 2 || warp == 3)
    {
        // ===================== MMA issuers (warp 3 also owns the TMEM allocation) =====================
        const int issuer = warp - 2;
        if (lane == 0)
        {
            constexpr uint32_t IDESC = KindTraits<KIND_I8>::idesc(128, MU_TILE_N);
            uint32_t wcnt = 0, acnt = 0, tseq = 0; // this issuer's weight groups / activation k-blocks so far; tiles so far
            for (int l = 0; l < p.n_layers; l++)
            {
                const int tiles = (p.fan_out[l] + MU_TILE_N - 1) / MU_TILE_N, nkb = (p.fan_in[l] + 127) >> 7;
                const int ngw = (nkb + MU_W_GROUP - 1) / MU_W_GROUP;
                for (int tile = cta; tile < tiles; tile += grid, tseq++)
                {
                    const uint32_t acc = tseq & 1u;
                    mbar_wait(tempty(acc), ((tseq >> 1) & 1u) ^ 1u, p.error_flag, KERR_MU_MMA);
                    tcgen05_fence_after();
                    if (issuer >= ngw) // a short K leaves this issuer without a group: its accumulator is not read either
                    {
                        mbar_arrive(tfull(acc));
                        continue;
                    }
                    const uint32_t d_tmem = tmem_base + (acc * MU_ISSUERS + issuer) * MU_TILE_N;
                    for (int g = issuer; g < ngw; g += MU_ISSUERS, wcnt++)
                    {
                        const int ws = issuer * MU_W_SLOTS_PER + (int)(wcnt % MU_W_SLOTS_PER), kb1 = min(nkb, (g + 1) * MU_W_GROUP);
                        mbar_wait(wfull(ws), (wcnt / MU_W_SLOTS_PER) & 1u, p.error_flag, KERR_MU_MMA);
                        for (int kb = g * MU_W_GROUP; kb < kb1; kb++, acnt++)
                        {
                            const int as = issuer * MU_A_SLOTS_PER + (int)(acnt % MU_A_SLOTS_PER), wi = kb % MU_W_GROUP;
                            mbar_wait(afull(as), (acnt / MU_A_SLOTS_PER) & 1u, p.error_flag, KERR_MU_MMA);
                            tcgen05_fence_after();
                            if (dbg && issuer == 0 && kb == 0) dbg[l * 8 + 1] = clock64(); // first operands of the layer landed
                            const uint64_t a_desc = umma_smem_desc_sw128(base + MU_OFF_A + as * MU_A_SLOT_BYTES);
                            const uint64_t b_desc = umma_smem_desc_sw128(base + ws * MU_W_SLOT_BYTES + wi * MU_W_TILE_BYTES);
#pragma unroll
                            for (int k = 0; k < 4; k++) // 4 x 32 bytes of K per k-block
                                umma_ss<KIND_I8>(d_tmem, a_desc + 2u * k, b_desc + 2u * k, IDESC, (g != issuer || wi != 0 || k != 0) ? 1u : 0u);
                            tcgen05_commit(aempty(as));
                        }
                        tcgen05_commit(wempty(ws));
                    }
                    tcgen05_commit(tfull(acc));
                    if (dbg && issuer == 0) dbg[l * 8 + 2] = clock64(); // this issuer's last MMA of the layer issued
                }
            }
        }
    }
    else if (warp >= MU_EPI_WARP0)
    {
        // ===================== epilogue: thread = sample =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t tseq = 0;
        for (int l = 0; l < p.n_layers; l++)
        {
            const int tiles = (p.fan_out[l] + MU_TILE_N - 1) / MU_TILE_N;
            const int nacc = min(MU_ISSUERS, (((p.fan_in[l] + 127) >> 7) + MU_W_GROUP - 1) / MU_W_GROUP); // accumulators that were written
            const bool last = l + 1 == p.n_layers;
            const bool relu = (p.relu_mask >> l) & 1u;
            for (int tile = cta; tile < tiles; tile += grid, tseq++)
            {
                const uint32_t acc = tseq & 1u;
                const int col0 = tile * MU_TILE_N;
                // bias of the tile's 32 neurons (the same for every thread: broadcast loads), fetched before the accumulator is waited for
                int bias[32];
#pragma unroll
                for (int j4 = 0; j4 < 8; j4++)
                {
                    int4 b4 = make_int4(0, 0, 0, 0);
                    if (col0 + 4 * j4 + 3 < p.fan_out[l])
                        b4 = __ldg(reinterpret_cast<const int4 *>(p.bias[l] + col0) + j4);
                    else
                    {
                        int t[4] = {0, 0, 0, 0};
                        for (int e = 0; e < 4; e++)
                            if (col0 + 4 * j4 + e < p.fan_out[l]) t[e] = __ldg(p.bias[l] + col0 + 4 * j4 + e);
                        b4 = make_int4(t[0], t[1], t[2], t[3]);
                    }
                    bias[4 * j4] = b4.x, bias[4 * j4 + 1] = b4.y, bias[4 * j4 + 2] = b4.z, bias[4 * j4 + 3] = b4.w;
                }
                mbar_wait(tfull(acc), (tseq >> 1) & 1u, p.error_flag, KERR_MU_EPILOGUE);
                tcgen05_fence_after();
                if (dbg && threadIdx.x == MU_EPI_WARP0 * 32) dbg[l * 8 + 3] = clock64(); // accumulators complete
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * MU_ISSUERS * MU_TILE_N, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 1; i < MU_ISSUERS; i++)
                    if (i < nacc) // (uniform)
                    {
                        uint32_t u[32];
                        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (acc * MU_ISSUERS + i) * MU_TILE_N, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] += u[j];
                    }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty(acc)); // the accumulator may be overwritten by the tile after next
                if (row < p.batch)
                {
                    if (last)
                    {
                        int32_t *dst = p.out + (long long)row * p.fan_out[l] + col0;
#pragma unroll
                        for (int j = 0; j < 32; j++)
                        {
                            int a = (int)v[j] + bias[j];
                            if (relu) a = max(a, 0);
                            if (col0 + j < p.fan_out[l]) dst[j] = a;
                        }
                    }
                    else
                    {
                        uint32_t w[8];
#pragma unroll
                        for (int j4 = 0; j4 < 8; j4++)
                        {
                            uint32_t word = 0;
#pragma unroll
                            for (int e = 0; e < 4; e++)
                            {
                                int a = (int)v[4 * j4 + e] + bias[4 * j4 + e];
                                if (relu) a = max(a, 0);
                                a = min(127, max(-128, a >> 7));
                                word |= ((uint32_t)a & 0xFFu) << (8 * e);
                            }
                            w[j4] = word;
                        }
                        int8_t *dst = p.act_out[l] + (long long)row * p.ld_out[l] + col0;
                        if (col0 + MU_TILE_N <= p.fan_out[l])
                        {
                            reinterpret_cast<uint4 *>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                            reinterpret_cast<uint4 *>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                        }
                        else
                        {
                            for (int j = 0; j < 32; j++)
                                if (col0 + j < p.fan_out[l]) dst[j] = (int8_t)(w[j >> 2] >> (8 * (j & 3)));
                        }
                        // these generic-proxy stores are read by other CTAs' TMA (async proxy) after the grid barrier
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                    }
                }
            }
            // Publish this CTA's output tiles of the layer.  The counter counts TILES, cumulatively over the layers, and a CTA without a
            // tile in a layer adds nothing: only a CTA that has passed the wait for layer l (every tile of the layers before it is
            // stored) can add a tile of layer l, so the count reaches "all tiles up to layer l" only when they all are.  (Counting one
            // arrival per CTA and layer is wrong here: the CTAs without a tile in a narrow layer arrive for it at once -- they wait for
            // nothing -- and their early arrivals would stand in for CTAs still storing the layer before.)
            const int my_tiles = cta < tiles ? (tiles - cta + grid - 1) / grid : 0;
            if (!last && my_tiles > 0)
            {
                named_bar_sync(1, 128); // orders every epilogue thread's stores before the release-add (cumulativity)
                if (threadIdx.x == MU_EPI_WARP0 * 32) mu_red_release_gpu_add(p.barrier, (unsigned)my_tiles);
                if (dbg && threadIdx.x == MU_EPI_WARP0 * 32) dbg[l * 8 + 4] = clock64(); // outputs published
            }
        }
        // The last CTA to get here leaves both counters at zero for the next launch (launches of one handle are stream-ordered).
        if (threadIdx.x == MU_EPI_WARP0 * 32 && atomicAdd(p.barrier + 1, 1u) == (unsigned)grid - 1u)
        {
            p.barrier[0] = 0u;
            p.barrier[1] = 0u;
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 3)
    {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, MU_TMEM_COLS);
    }
}

// =====================================================================================================
// Split-K CTA clusters (mlp_i8_umma_cluster_kernel): the same net-in-one-launch kernel, CL = 2 or 4 CTAs per neuron tile
// =====================================================================================================
// What bounds the kernel above (measured): (1) at 128 samples every CTA pulls the whole activation matrix of a layer
// (128 x fan_in bytes, 512 KB on config C5) out of L2 -- 128 CTAs x 512 KB = 64 MB per layer against 16 MB of weights, and the L2 -> SM
// fabric moves ~6300 B/clk chip-wide (TMA multicast over a cluster of 2..4 saves nothing there: L2 already merges those requests);
// (2) at every batch the tensor pipe itself: a kind::i8 MMA of 128 samples x N neurons x 32 bytes costs ~70 cycles for N = 32 and for
// N = 64 alike (tools/umma_pair_timeline.py: 64 MMAs per layer and CTA in 4.5 k cycles, the same with two or four issuing threads,
// 17 or 128 samples, with or without the weights prefetched into L2) -- the 4 KB of the 128-row A operand are read from shared memory
// for every MMA, whatever N is.  Both shrink when a CLUSTER of CL CTAs shares a tile of 32 CL neurons and splits K:
//   * CTA r of the cluster streams the weight rows [32 CL t, 32 CL (t + 1)) over ITS 1 / CL of the k-blocks (same bytes per CTA as
//     before) and loads the activations of that K range only (1 / CL of the bytes); its MMAs are 128 x 32 CL x 32 (1 / CL as many);
//   * the partial sums meet in distributed shared memory: every epilogue thread (thread = sample) sends each peer the 32 columns that
//     peer finalises -- 128 bytes as eight st.async, which land in a window of the peer's shared memory and are counted on the peer's
//     mbarrier like TMA bytes: no fence, no arrive (plain st.shared::cluster + a release.cluster arrive cost a MEMBAR.GPU per thread
//     and tile: 62 -> 57 us) -- and adds the CL - 1 rows of 32 columns it receives to its own; CTA r then requantises and stores
//     neurons [32 CL t + 32 r, + 32).  Integer sums: order-independent, still bit-exact.  Only the batch's rows travel.
//   * everything else is the kernel above: per-issuer rings, the weight producer running ahead across the layers, the cumulative
//     tile counter as grid barrier (every CTA publishes the tiles of its cluster: the target is CL x tiles).
// Needs every layer to have at least CL k-blocks; the launcher picks CL = 4 (fan-ins >= 512), CL = 2 (> 128) or the single-CTA kernel.
// Config C5, us per forward at 17 / 64 / 128 samples: single CTAs 66.6 / 67.9 / 70.2; pairs 55.5 / 56.2 / 59.6 (two issuers: 57.5 / 58.9 / 61.8).
constexpr int MP_FIN_N = 32;                     // neurons a CTA finalises per tile
constexpr int MP_X_BYTES = 128 * MP_FIN_N * 4;   // one exchange window, 16 KB: [8 column quads][128 samples][4 x int32]
// NI MMA-issuing threads per CTA (2 or 4), each with weight slots, activation slots and an accumulator of its own.
template <int CL, int NI>
struct MpLayout
{
    static_assert(CL == 2 || CL == 4, "two or four CTAs per tile");
    static_assert(NI == 2 || NI == 4, "two or four issuers");
    static constexpr int TILE_N = MP_FIN_N * CL;           // neurons per cluster tile = N of the MMAs
    static constexpr int W_TILE_BYTES = TILE_N * 128;      // one k-block of a tile's weight rows
    static constexpr int W_SLOTS_PER = 2;
    static constexpr int W_SLOTS = NI * W_SLOTS_PER;
    static constexpr int W_GROUP = 65536 / (W_SLOTS * W_TILE_BYTES); // k-blocks per weight slot: the ring holds 64 KB
    static_assert(W_GROUP >= 1, "weight ring too small for this tile");
    static constexpr int W_SLOT_BYTES = W_GROUP * W_TILE_BYTES;
    // Activations: an equal region per issuer, cut into slots of the batch's size (128 samples: 16 KB, <= 64: 8 KB, <= 32: 4 KB; at
    // most 8 slots per issuer) -- a small batch gets a deeper ring, i.e. fewer L2 round trips in the critical path of a layer.
    // 128 KB for pairs; 64 KB for clusters of four (a quarter of K per CTA, and three exchange windows per stage instead of one).
    static constexpr int A_REGION_LOG = (CL == 2 ? 17 : 16) - (NI == 4 ? 2 : 1);
    static constexpr int A_SLOTS_PER = 8;                  // barriers per issuer (slots in use: region >> slot shift, capped)
    static constexpr int A_SLOTS = NI * A_SLOTS_PER;
    static constexpr int OFF_A = W_SLOTS * W_SLOT_BYTES;                // 64 KB
    static constexpr int OFF_X = OFF_A + (NI << A_REGION_LOG);          // the exchange windows: [2 stages][CL - 1 senders]
    static constexpr int OFF_BARS = OFF_X + 2 * (CL - 1) * MP_X_BYTES;
    static constexpr int NUM_BARS = 2 * W_SLOTS + 2 * A_SLOTS + 4 + 2;
    static constexpr int OFF_TMEM_PTR = OFF_BARS + NUM_BARS * 8;
    static constexpr int SMEM = OFF_TMEM_PTR + 16;
    static_assert(SMEM <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
    static constexpr uint32_t TMEM_COLS = 2 * NI * TILE_N;              // two accumulator stages x one accumulator per issuer
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM allocations are powers of two, at most 512 columns");
    // warps: 0 weight producer, 1 activation producer, 2 / 3 issuers 0 / 1, 4..7 epilogue (the TMEM lane quarters), 8 / 9 issuers 2 / 3 (NI = 4)
    static constexpr int W_ISSUER2 = MU_EPI_WARP0 + 4;
    static constexpr int THREADS = (W_ISSUER2 + NI - 2) * 32;
};

enum : int
{
    KERR_MP_EXCHANGE = 36,
};

template <int CL, int NI>
__global__ void __launch_bounds__((MpLayout<CL, NI>::THREADS), 1)
mlp_i8_umma_cluster_kernel(const __grid_constant__ MlpUmmaMaps maps, const MlpUmmaParams p)
{
    using L = MpLayout<CL, NI>;
    extern __shared__ __align__(1024) uint8_t mp_smem[];
    const uint32_t base = smem_u32(mp_smem);
    if ((base & 1023u) != 0)
    {
        if (threadIdx.x == 0 && p.error_flag) atomicExch(p.error_flag, KERR_SMEM_ALIGN);
        return; // uniform over the grid
    }
    const uint32_t bars = base + L::OFF_BARS;
    auto wfull = [&](int s) { return bars + 8u * s; };
    auto wempty = [&](int s) { return bars + 8u * (L::W_SLOTS + s); };
    auto afull = [&](int s) { return bars + 8u * (2 * L::W_SLOTS + s); };
    auto aempty = [&](int s) { return bars + 8u * (2 * L::W_SLOTS + L::A_SLOTS + s); };
    auto tfull = [&](int a) { return bars + 8u * (2 * L::W_SLOTS + 2 * L::A_SLOTS + a); };
    auto tempty = [&](int a) { return bars + 8u * (2 * L::W_SLOTS + 2 * L::A_SLOTS + 2 + a); };
    auto xfull = [&](int x) { return bars + 8u * (2 * L::W_SLOTS + 2 * L::A_SLOTS + 4 + x); };
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(mp_smem + L::OFF_TMEM_PTR);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31; // (uniform: the issuers' descriptors stay in uniform registers, see ptx.cuh)
    const int rank = (int)cluster_ctarank();                     // which part of K, which 32 of the tile's neurons
    const int grp = blockIdx.x / CL, ngrps = gridDim.x / CL;     // cluster index: walks the tiles
    // this CTA's k-blocks of a layer (never empty: nkb >= CL)
    auto k_lo = [&](int nkb) { return rank * nkb / CL; };
    auto k_hi = [&](int nkb) { return (rank + 1) * nkb / CL; };
    const int a_slots_log = min(3, L::A_REGION_LOG - p.a_slot_shift);  // activation slots per issuer in use (2, 4 or 8)
    const uint32_t a_mask = (1u << a_slots_log) - 1u;
    auto a_slot_addr = [&](int issuer, uint32_t slot) { return base + L::OFF_A + ((uint32_t)issuer << L::A_REGION_LOG) + (slot << p.a_slot_shift); };
#ifdef NETCUDA_DEBUG_TIMELINE
    long long *const dbg = blockIdx.x == 0 ? p.debug : nullptr; // [n_layers][8] clock64 stamps of CTA 0 (tools/umma_pair_timeline.py)
#else
    constexpr long long *dbg = nullptr;
#endif

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < L::W_SLOTS; s++) mbar_init(wfull(s), 1), mbar_init(wempty(s), 1);
        for (int s = 0; s < L::A_SLOTS; s++) mbar_init(afull(s), 1), mbar_init(aempty(s), 1);
        for (int a = 0; a < 2; a++) mbar_init(tfull(a), NI), mbar_init(tempty(a), 4);
        for (int x = 0; x < 2; x++) mbar_init(xfull(x), 1); // one local arrive.expect_tx per tile; the peers' bytes complete the phase
        fence_barrier_init();
    }
    if (warp == 3)
    {
        tmem_alloc(base + L::OFF_TMEM_PTR, L::TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    cluster_sync_all(); // the peers' exchange barriers are initialised before anything is counted on them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0)
    {
        // ===================== weight producer: never waits for activations =====================
        // The weights come from the tiled copy (elementwise.cu: 16 KB blocks of 128 neurons x 128 bytes of K, already in the 128B-swizzled
        // byte order): what a CTA streams is contiguous in HBM -- its K range of a 128-neuron tile is ONE run of bytes, of a 64-neuron tile
        // 8 KB out of every 16 -- and arrives by 1-D bulk copies.  Out of the row-major matrix a k-block of a tile is 64..128 pieces of
        // 128 bytes, 4 KB apart: the 2-D TMA loads of that took 2.6-3 k cycles each with every CTA at it (tools/umma_pair_timeline.py),
        // and the issuers spent a layer waiting for their second pair of weight slots.
        if (lane == 0)
        {
            const uint64_t stream_once = l2_policy_evict_first();
            uint32_t cnt[NI] = {};
            for (int l = 0; l < p.n_layers; l++)
            {
                const int tiles = (p.fan_out[l] + L::TILE_N - 1) / L::TILE_N, nkb = (p.fan_in[l] + 127) >> 7;
                const int lo = k_lo(nkb), hi = k_hi(nkb);
                for (int tile = grp; tile < tiles; tile += ngrps)
                {
                    // first k-block of this tile's rows: a whole 16 KB block per k-block (CL = 4), or one 8 KB half of it (CL = 2)
                    const int8_t *src0 = p.w_tiled[l] + (CL == 4 ? (size_t)tile * nkb * 16384 : ((size_t)(tile >> 1) * nkb * 16384 + (size_t)(tile & 1) * 8192));
                    for (int kb = lo, g = 0; kb < hi; kb += L::W_GROUP, g++)
                    {
                        const int is = g % NI;
                        uint32_t c = 0;
#pragma unroll
                        for (int i = 0; i < NI; i++)
                            if (i == is) c = cnt[i]++;
                        const int s = is * L::W_SLOTS_PER + (int)(c % L::W_SLOTS_PER), nb = min(L::W_GROUP, hi - kb);
                        mbar_wait(wempty(s), ((c / L::W_SLOTS_PER) & 1u) ^ 1u, p.error_flag, KERR_MU_W_PRODUCER);
                        mbar_arrive_expect_tx(wfull(s), (uint32_t)(nb * L::W_TILE_BYTES));
                        if constexpr (CL == 4)
                            bulk_load_1d(base + s * L::W_SLOT_BYTES, src0 + (size_t)kb * 16384, (uint32_t)(nb * L::W_TILE_BYTES), wfull(s), stream_once);
                        else
                            for (int i = 0; i < nb; i++)
                                bulk_load_1d(base + s * L::W_SLOT_BYTES + i * L::W_TILE_BYTES, src0 + (size_t)(kb + i) * 16384, L::W_TILE_BYTES, wfull(s), stream_once);
                    }
                }
            }
        }
    }
    else if (warp == 1)
    {
        // ===================== activation producer: this CTA's K range of layer l, after the grid barrier behind layer l - 1 =====================
        // (one producing thread per issuer, and the weights beyond the ring prefetched into L2, were measured: no gain -- the MMA phase
        // of a layer is paced by the tensor pipe, see the header)
        if (lane == 0)
        {
            for (int l = 0; l < p.n_layers; l++) tma_prefetch_desc(&maps.a[l]);
            uint32_t cnt[NI] = {}; // k-blocks handed to each issuer so far
            unsigned target = 0; // every CTA of a cluster publishes the cluster's tiles: CL x the output tiles of all the layers before this one
            for (int l = 0; l < p.n_layers; l++)
            {
                const int tiles = (p.fan_out[l] + L::TILE_N - 1) / L::TILE_N, nkb = (p.fan_in[l] + 127) >> 7;
                const int lo = k_lo(nkb), hi = k_hi(nkb);
                if (l > 0) target += (unsigned)CL * (unsigned)((p.fan_out[l - 1] + L::TILE_N - 1) / L::TILE_N);
                if (l > 0 && grp < tiles)
                {
                    const long long t0 = clock64();
                    while (mu_ld_acquire_gpu(p.barrier) < target)
                        if (clock64() - t0 > 4000000000LL)
                        {
                            if (p.error_flag) atomicExch(p.error_flag, KERR_MU_GRID_BARRIER);
                            __threadfence_system();
                            __trap();
                        }
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                }
                if (dbg) dbg[l * 8 + 0] = clock64(); // barrier passed
                for (int tile = grp; tile < tiles; tile += ngrps)
                    for (int kb = lo; kb < hi; kb++)
                    {
                        const int is = ((kb - lo) / L::W_GROUP) % NI; // the issuer of the weight group this k-block belongs to
                        uint32_t c = 0;
#pragma unroll
                        for (int i = 0; i < NI; i++)
                            if (i == is) c = cnt[i]++;
                        const int s = is * L::A_SLOTS_PER + (int)(c & a_mask);
                        mbar_wait(aempty(s), ((c >> a_slots_log) & 1u) ^ 1u, p.error_flag, KERR_MU_A_PRODUCER);
                        mbar_arrive_expect_tx(afull(s), (uint32_t)(((p.batch + 7) & ~7) * 128));
                        tma_load_2d(a_slot_addr(is, c & a_mask), &maps.a[l], afull(s), kb * 128, 0);
                    }
            }
        }
    }
    else if (warp == 2 || warp == 3 || warp >= L::W_ISSUER2)
    {
        // ===================== MMA issuers: 128 samples x 32 CL neurons x 32 bytes of K per instruction =====================
        const int issuer = warp < 4 ? warp - 2 : warp - L::W_ISSUER2 + 2;
        if (lane == 0)
        {
            constexpr uint32_t IDESC = KindTraits<KIND_I8>::idesc(128, L::TILE_N);
            uint32_t wcnt = 0, acnt = 0, tseq = 0;
            for (int l = 0; l < p.n_layers; l++)
            {
                const int tiles = (p.fan_out[l] + L::TILE_N - 1) / L::TILE_N, nkb = (p.fan_in[l] + 127) >> 7;
                const int mykb = k_hi(nkb) - k_lo(nkb);
                const int ngw = (mykb + L::W_GROUP - 1) / L::W_GROUP;
                for (int tile = grp; tile < tiles; tile += ngrps, tseq++)
                {
                    const uint32_t acc = tseq & 1u;
                    mbar_wait(tempty(acc), ((tseq >> 1) & 1u) ^ 1u, p.error_flag, KERR_MU_MMA);
                    tcgen05_fence_after();
                    if (issuer >= ngw) // a short K range leaves this issuer without a group: its accumulator is not read either
                    {
                        mbar_arrive(tfull(acc));
                        continue;
                    }
                    const uint32_t d_tmem = tmem_base + (acc * NI + issuer) * L::TILE_N;
                    for (int g = issuer; g < ngw; g += NI, wcnt++)
                    {
                        const int ws = issuer * L::W_SLOTS_PER + (int)(wcnt % L::W_SLOTS_PER), kr1 = min(mykb, (g + 1) * L::W_GROUP);
                        mbar_wait(wfull(ws), (wcnt / L::W_SLOTS_PER) & 1u, p.error_flag, KERR_MU_MMA);
                        if (dbg && l == 3 && g * L::W_GROUP < 16) dbg[128 + (issuer * 16 + g * L::W_GROUP) * 2] = clock64(); // weights of the group landed
                        for (int kr = g * L::W_GROUP; kr < kr1; kr++, acnt++) // kr: k-block relative to this CTA's first one
                        {
                            const int as = issuer * L::A_SLOTS_PER + (int)(acnt & a_mask), wi = kr % L::W_GROUP;
                            mbar_wait(afull(as), (acnt >> a_slots_log) & 1u, p.error_flag, KERR_MU_MMA);
                            tcgen05_fence_after();
                            if (dbg && l == 3 && kr < 16) dbg[128 + (issuer * 16 + kr) * 2 + 1] = clock64(); // activations of the k-block landed
                            if (dbg && issuer == 0 && kr == 0) dbg[l * 8 + 1] = clock64(); // first operands of the layer landed
                            const uint64_t a_desc = umma_smem_desc_sw128(a_slot_addr(issuer, acnt & a_mask));
                            const uint64_t b_desc = umma_smem_desc_sw128(base + ws * L::W_SLOT_BYTES + wi * L::W_TILE_BYTES);
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                umma_ss<KIND_I8>(d_tmem, a_desc + 2u * k, b_desc + 2u * k, IDESC, (g != issuer || wi != 0 || k != 0) ? 1u : 0u);
                            tcgen05_commit(aempty(as));
                        }
                        tcgen05_commit(wempty(ws));
                    }
                    tcgen05_commit(tfull(acc));
                    if (dbg && issuer == 0) dbg[l * 8 + 2] = clock64(); // issuer 0's last MMA of the layer issued
                }
            }
        }
    }
    else if (warp >= MU_EPI_WARP0 && warp < MU_EPI_WARP0 + 4)
    {
        // ===================== epilogue: thread = sample; partial sums of the CL K ranges meet through DSMEM =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool live = row < p.batch; // only the batch's rows are exchanged and stored
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t tseq = 0;
        for (int l = 0; l < p.n_layers; l++)
        {
            const int tiles = (p.fan_out[l] + L::TILE_N - 1) / L::TILE_N, nkb = (p.fan_in[l] + 127) >> 7;
            const int ngw = (k_hi(nkb) - k_lo(nkb) + L::W_GROUP - 1) / L::W_GROUP;
            const int nacc = min(NI, ngw); // accumulators this CTA's issuers wrote
            const bool last = l + 1 == p.n_layers;
            const bool relu = (p.relu_mask >> l) & 1u;
            for (int tile = grp; tile < tiles; tile += ngrps, tseq++)
            {
                const uint32_t acc = tseq & 1u, xs = tseq & 1u;
                const int col0 = tile * L::TILE_N + rank * MP_FIN_N; // the 32 neurons this CTA finalises
                int bias[32];
#pragma unroll
                for (int j4 = 0; j4 < 8; j4++)
                {
                    int4 b4 = make_int4(0, 0, 0, 0);
                    if (col0 + 4 * j4 + 3 < p.fan_out[l])
                        b4 = __ldg(reinterpret_cast<const int4 *>(p.bias[l] + col0) + j4);
                    else
                    {
                        int t[4] = {0, 0, 0, 0};
                        for (int e = 0; e < 4; e++)
                            if (col0 + 4 * j4 + e < p.fan_out[l]) t[e] = __ldg(p.bias[l] + col0 + 4 * j4 + e);
                        b4 = make_int4(t[0], t[1], t[2], t[3]);
                    }
                    bias[4 * j4] = b4.x, bias[4 * j4 + 1] = b4.y, bias[4 * j4 + 2] = b4.z, bias[4 * j4 + 3] = b4.w;
                }
                // (window stage xs was last waited for two tiles ago by this very thread: its barrier can be armed for this tile)
                if (threadIdx.x == MU_EPI_WARP0 * 32) mbar_arrive_expect_tx(xfull(xs), (uint32_t)((CL - 1) * p.batch * 128));
                mbar_wait(tfull(acc), (tseq >> 1) & 1u, p.error_flag, KERR_MU_EPILOGUE);
                tcgen05_fence_after();
                if (dbg && threadIdx.x == MU_EPI_WARP0 * 32) dbg[l * 8 + 3] = clock64(); // accumulators complete
                const uint32_t t_acc = lane_base + acc * NI * L::TILE_N;
                // the peers' columns first: they are on their way while this thread reads its own
#pragma unroll
                for (int d = 1; d < CL; d++)
                {
                    const uint32_t peer = (uint32_t)(rank + d) % CL;
                    uint32_t u[32];
                    tmem_ld_32x32(t_acc + peer * MP_FIN_N, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 1; i < NI; i++)
                        if (i < nacc) // (uniform)
                        {
                            uint32_t t[32];
                            tmem_ld_32x32(t_acc + i * L::TILE_N + peer * MP_FIN_N, t);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; j++) u[j] += t[j];
                        }
                    if (live)
                    {
                        // Stage xs of the peer's windows is free: the peer's threads read it for tile t - 2 (and used what they read)
                        // before they sent for tile t - 1, and this thread has waited for all of those bytes (its own wait of tile t - 1).
                        // A peer keeps one window per sender: this CTA's is number rank (or rank - 1 behind the peer's own rank).
                        const uint32_t win = (uint32_t)(rank < (int)peer ? rank : rank - 1);
                        const uint32_t dst = mapa_shared(base + L::OFF_X + (xs * (CL - 1) + win) * MP_X_BYTES, peer) + (uint32_t)row * 16u;
                        const uint32_t peer_bar = mapa_shared(xfull(xs), peer);
#pragma unroll
                        for (int j4 = 0; j4 < 8; j4++)
                            st_async_cluster_v4(dst + j4 * 2048u, u[4 * j4], u[4 * j4 + 1], u[4 * j4 + 2], u[4 * j4 + 3], peer_bar);
                    }
                }
                if (dbg && threadIdx.x == MU_EPI_WARP0 * 32) dbg[l * 8 + 4] = clock64(); // the peers' columns are on their way
                uint32_t v[32];
                tmem_ld_32x32(t_acc + rank * MP_FIN_N, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 1; i < NI; i++)
                    if (i < nacc)
                    {
                        uint32_t t[32];
                        tmem_ld_32x32(t_acc + i * L::TILE_N + rank * MP_FIN_N, t);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] += t[j];
                    }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty(acc)); // the accumulators may be overwritten by the tile after next
                mbar_wait_acquire_cluster(xfull(xs), (tseq >> 1) & 1u, p.error_flag, KERR_MP_EXCHANGE);
                if (dbg && threadIdx.x == MU_EPI_WARP0 * 32) dbg[l * 8 + 5] = clock64(); // the peers' partial sums are here
                if (live && col0 < p.fan_out[l])
                {
#pragma unroll
                    for (int w = 0; w < CL - 1; w++)
                    {
                        const uint4 *src = reinterpret_cast<const uint4 *>(mp_smem + L::OFF_X + (xs * (CL - 1) + w) * MP_X_BYTES) + row;
#pragma unroll
                        for (int j4 = 0; j4 < 8; j4++)
                        {
                            const uint4 r4 = src[j4 * 128];
                            v[4 * j4] += r4.x, v[4 * j4 + 1] += r4.y, v[4 * j4 + 2] += r4.z, v[4 * j4 + 3] += r4.w;
                        }
                    }
                    if (last)
                    {
                        int32_t *dst = p.out + (long long)row * p.fan_out[l] + col0;
#pragma unroll
                        for (int j = 0; j < 32; j++)
                        {
                            int a = (int)v[j] + bias[j];
                            if (relu) a = max(a, 0);
                            if (col0 + j < p.fan_out[l]) dst[j] = a;
                        }
                    }
                    else
                    {
                        uint32_t w[8];
#pragma unroll
                        for (int j4 = 0; j4 < 8; j4++)
                        {
                            uint32_t word = 0;
#pragma unroll
                            for (int e = 0; e < 4; e++)
                            {
                                int a = (int)v[4 * j4 + e] + bias[4 * j4 + e];
                                if (relu) a = max(a, 0);
                                a = min(127, max(-128, a >> 7));
                                word |= ((uint32_t)a & 0xFFu) << (8 * e);
                            }
                            w[j4] = word;
                        }
                        int8_t *dst = p.act_out[l] + (long long)row * p.ld_out[l] + col0;
                        if (col0 + MP_FIN_N <= p.fan_out[l])
                        {
                            reinterpret_cast<uint4 *>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                            reinterpret_cast<uint4 *>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                        }
                        else
                        {
                            for (int j = 0; j < 32; j++)
                                if (col0 + j < p.fan_out[l]) dst[j] = (int8_t)(w[j >> 2] >> (8 * (j & 3)));
                        }
                        // these generic-proxy stores are read by other CTAs' TMA (async proxy) after the grid barrier
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                    }
                }
            }
            // publish the cluster's tiles of this layer (see the kernel above for why the counter counts tiles, not CTAs)
            const int my_tiles = grp < tiles ? (tiles - grp + ngrps - 1) / ngrps : 0;
            if (!last && my_tiles > 0)
            {
                if (dbg && threadIdx.x == MU_EPI_WARP0 * 32) dbg[l * 8 + 6] = clock64(); // outputs stored
                named_bar_sync(1, 128);
                if (threadIdx.x == MU_EPI_WARP0 * 32) mu_red_release_gpu_add(p.barrier, (unsigned)my_tiles);
                if (dbg && threadIdx.x == MU_EPI_WARP0 * 32) dbg[l * 8 + 7] = clock64(); // ... and published
            }
        }
        if (threadIdx.x == MU_EPI_WARP0 * 32 && atomicAdd(p.barrier + 1, 1u) == gridDim.x - 1u)
        {
            p.barrier[0] = 0u;
            p.barrier[1] = 0u;
        }
    }

    tcgen05_fence_before();
    cluster_sync_all(); // no CTA exits while a peer may still write into its exchange windows or count bytes on its barriers
    if (warp == 3)
    {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, L::TMEM_COLS);
    }
}

bool mlp_umma_stream_supported(const MlpStreamParams &p)
{
    if (p.n_layers < 1 || p.n_layers > MLP_STREAM_MAX_LAYERS || p.batch < 1 || p.batch > MLP_UMMA_STREAM_MAX_BATCH) return false;
    for (int l = 0; l < p.n_layers; l++)
    {
        const MlpStreamLayer &ly = p.layers[l];
        if (ly.fan_in < 16 || (ly.fan_in & 15) || ly.fan_out < 1) return false;           // TMA: 16-byte row pitch
        if (l + 1 < p.n_layers && (ly.fan_out & 15)) return false;                          // (the next layer's rows are unpadded)
        if ((reinterpret_cast<uintptr_t>(ly.w) & 15u) != 0 || (reinterpret_cast<uintptr_t>(ly.bias) & 15u) != 0) return false;
    }
    return true;
}

cudaError_t launch_mlp_i8_umma_stream(const MlpStreamParams &sp, int num_sms, cudaStream_t stream)
{
    if (!mlp_umma_stream_supported(sp)) return cudaErrorInvalidValue;
    static bool opted[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !opted[dev])
    {
        cudaError_t e = cudaFuncSetAttribute(mlp_i8_umma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MU_SMEM);
        if (e != cudaSuccess) return e;
        opted[dev] = true;
    }
    MlpUmmaMaps maps;
    MlpUmmaParams p;
    p.n_layers = sp.n_layers, p.batch = sp.batch, p.relu_mask = sp.relu_mask;
    p.out = sp.out, p.barrier = sp.barrier, p.error_flag = sp.error_flag;
    p.debug = sp.debug;
    int max_tiles = 1;
    for (int l = 0; l < sp.n_layers; l++)
    {
        const MlpStreamLayer &ly = sp.layers[l];
        const bool last = l + 1 == sp.n_layers;
        p.fan_in[l] = ly.fan_in, p.fan_out[l] = ly.fan_out, p.bias[l] = ly.bias;
        p.act_out[l] = last ? nullptr : sp.act[(l + 1) & 1];
        p.ld_out[l] = ly.fan_out; // hidden activations are written unpadded: row pitch = fan_out = the next layer's fan_in
        const int8_t *a_src = l == 0 ? sp.in : sp.act[l & 1];
        cudaError_t e = encode_tma_2d(&maps.a[l], 1, a_src, ly.fan_in, sp.batch, ly.fan_in, 128, (sp.batch + 7) & ~7, true);
        if (e != cudaSuccess) return e;
        e = encode_tma_2d(&maps.w[l], 1, ly.w, ly.fan_in, ly.fan_out, ly.fan_in, 128, MU_TILE_N, true);
        if (e != cudaSuccess) return e;
        max_tiles = std::max(max_tiles, (ly.fan_out + MU_TILE_N - 1) / MU_TILE_N);
    }
    const int grid = std::min(num_sms > 0 ? num_sms : 148, max_tiles);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3(MU_THREADS), cfg.dynamicSmemBytes = MU_SMEM, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; // all CTAs co-resident: the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, mlp_i8_umma_stream_kernel, maps, p);
}

// Cluster size the net can run with: 4 (every fan-in has at least four k-blocks), 2 (at least two), 0 = single-CTA kernel only
int mlp_umma_cluster_size(const MlpStreamParams &p, int num_sms)
{
    if (num_sms < 2 || !mlp_umma_stream_supported(p)) return 0;
    int min_fan_in = 1 << 30;
    for (int l = 0; l < p.n_layers; l++)
    {
        if (!p.layers[l].w_tiled) return 0; // (the runtime builds the tiled copy at upload for the nets this kernel can serve)
        min_fan_in = std::min(min_fan_in, p.layers[l].fan_in);
    }
    return min_fan_in > 384 && num_sms >= 4 ? 4 : min_fan_in > 128 ? 2 : 0;
}

template <int CL, int NI>
static cudaError_t launch_cluster(const MlpStreamParams &sp, int num_sms, cudaStream_t stream)
{
    using L = MpLayout<CL, NI>;
    static bool opted[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !opted[dev])
    {
        cudaError_t e = cudaFuncSetAttribute(mlp_i8_umma_cluster_kernel<CL, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM);
        if (e != cudaSuccess) return e;
        opted[dev] = true;
    }
    MlpUmmaMaps maps;
    MlpUmmaParams p;
    p.n_layers = sp.n_layers, p.batch = sp.batch, p.relu_mask = sp.relu_mask;
    p.out = sp.out, p.barrier = sp.barrier, p.error_flag = sp.error_flag;
    p.debug = sp.debug;
    const int box_bytes = ((sp.batch + 7) & ~7) * 128;
    p.a_slot_shift = box_bytes <= 4096 ? 12 : box_bytes <= 8192 ? 13 : 14;
    int max_tiles = 1;
    for (int l = 0; l < sp.n_layers; l++)
    {
        const MlpStreamLayer &ly = sp.layers[l];
        const bool last = l + 1 == sp.n_layers;
        p.fan_in[l] = ly.fan_in, p.fan_out[l] = ly.fan_out, p.bias[l] = ly.bias;
        p.act_out[l] = last ? nullptr : sp.act[(l + 1) & 1];
        p.ld_out[l] = ly.fan_out;
        const int8_t *a_src = l == 0 ? sp.in : sp.act[l & 1];
        cudaError_t e = encode_tma_2d(&maps.a[l], 1, a_src, ly.fan_in, sp.batch, ly.fan_in, 128, (sp.batch + 7) & ~7, true);
        if (e != cudaSuccess) return e;
        p.w_tiled[l] = ly.w_tiled; // (maps.w stays unused: the weights arrive by 1-D bulk copies)
        max_tiles = std::max(max_tiles, (ly.fan_out + L::TILE_N - 1) / L::TILE_N);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(L::THREADS), cfg.dynamicSmemBytes = L::SMEM, cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative; // all CTAs co-resident: the grid barrier cannot deadlock
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    // How many clusters the device holds at once: the SMs of a cluster share a GPC, so it can be fewer than num_sms / CL (a cooperative
    // launch of more is refused).  Asked once per device.
    static int max_clusters[64] = {};
    if (dev >= 64 || max_clusters[dev] == 0)
    {
        int n = 0;
        cfg.gridDim = dim3((unsigned)(CL * ((num_sms > 0 ? num_sms : 148) / CL))), cfg.numAttrs = 1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, mlp_i8_umma_cluster_kernel<CL, NI>, &cfg);
        if (e != cudaSuccess) return e;
        if (n < 1) return cudaErrorCooperativeLaunchTooLarge;
        if (dev < 64) max_clusters[dev] = n;
        else max_clusters[0] = n; // (not cached)
    }
    const int ngrps = std::min(std::min((num_sms > 0 ? num_sms : 148) / CL, max_clusters[dev < 64 ? dev : 0]), max_tiles);
    cfg.gridDim = dim3((unsigned)(CL * ngrps)), cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, mlp_i8_umma_cluster_kernel<CL, NI>, maps, p);
}

// mode (A/B codes): 1 = the product choice -- clusters of four (two issuers) where the net allows them and the batch is at most
// MLP_UMMA_QUAD_MAX_BATCH, else pairs (four issuers); 2 = pairs, four issuers; 3 = pairs, two issuers; 4 = clusters of four at every batch.
// Config C5, us per forward at 17 / 64 / 96 / 128 samples: clusters of four 49.7 / 52.9 / 57.5 / 62.0 (three 16 KB windows per CTA
// travel at 128 samples), pairs 53.6 / 55.6 / 56.8 / 59.5.
constexpr int MLP_UMMA_QUAD_MAX_BATCH = 88;
cudaError_t launch_mlp_i8_umma_cluster(const MlpStreamParams &sp, int num_sms, int mode, cudaStream_t stream)
{
    const int cl = mlp_umma_cluster_size(sp, num_sms);
    if (cl == 0) return cudaErrorInvalidValue;
    if (cl == 4 && ((mode == 1 && sp.batch <= MLP_UMMA_QUAD_MAX_BATCH) || mode == 4)) return launch_cluster<4, 2>(sp, num_sms, stream);
    if (mode == 3) return launch_cluster<2, 2>(sp, num_sms, stream);
    return launch_cluster<2, 4>(sp, num_sms, stream);
}

} // namespace nc

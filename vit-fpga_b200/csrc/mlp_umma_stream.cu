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
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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

} // namespace nc

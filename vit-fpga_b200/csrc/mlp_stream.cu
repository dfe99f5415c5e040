// mlp_stream.cu -- whole INT8 MLP forward for a handful of samples in ONE persistent kernel: weight streaming at HBM pace.
//
// Reference counterpart: the single-work-item task `network_v1` that walks all layers of the net for one sample
// (src/netFPGA.cpp:250,275; argument list :427-436,499-502).  For 1..32 samples of config C5 (8 x 4096 x 4096 int8) the forward is
// 128 MiB of weights against a few KB of activations: HBM-bound integer work, ~2 int-ops per weight byte and sample.  Padding such a
// batch to a 128-row tcgen05 tile and launching two kernels per layer (split-K GEMM + finalize) runs at ~1.5 TB/s; this kernel
// keeps the byte stream going instead:
//   * one CTA per SM (cooperative launch), every CTA owns a contiguous slice of output neurons of EVERY layer; a producer thread
//     pulls its weight rows, 16 at a time, through a shared-memory ring with 1-D bulk copies (cp.async.bulk + mbarrier
//     complete_tx; row pitch = 64 mod 128 bytes so that the fragment loads below are bank-conflict free).  Weights do not depend
//     on activations, so the ring keeps filling across the grid barrier between layers -- the stream does not drain;
//   * the 8 consumer warps split K: warp w keeps its K slice of all (<= 32) activation rows in REGISTERS for the whole layer, laid
//     out as the B fragments of mma.sync.m16n8k32.s8 (the CUDA-core dp4a form of this kernel needed as many shuffles as multiply-
//     adds to reduce over the lanes and was instruction-bound from 8 samples up); K is permuted so that every activation and
//     weight load is 16 bytes wide: 16 weight rows cost 2 shared-memory loads and 2-4 MMAs per 64 bytes of K.  The legacy
//     warp-level MMA is the right size here: M = 16 weight rows, N = 8 samples;
//   * per layer and CTA the 8 partial sums per (row, sample) meet in shared memory: + bias, ReLU, >> 7 and clamp exactly like the
//     GEMM epilogue (EPI_REQUANT[_RELU]) -- integer arithmetic, order-independent, bit-identical to the oracle;
//   * layers are separated by a grid barrier (global arrival counter, release / acquire at GPU scope, reset by the last CTA at kernel
//     end; activations read with ld.global.cg).  Its cost is the release: `fence.acq_rel.gpu` takes 1.3-1.5 us here (0.5 us once the
//     weight stream has ended) -- the acknowledgements of the few activation stores queue behind ~112 KB of weight data per layer on
//     this SM's return path -- out of a ~4.1 us chain per layer (barrier seen 0.4, activations 0.5, two tiles 1.0, partial sums 0.3,
//     release 1.5-1.8; tools/stream_timeline_all.py) against 2.6 us of pure weight streaming;
//   * up to 4 samples the hidden activations therefore travel WITHOUT barrier or fence, the way NCCL's LL protocol moves data: every
//     8-byte store carries 4 activations and a 4-byte tag (launch epoch, layer), a reader polls the words it needs (ld.volatile, 16
//     bytes = two tagged words) and has the data the moment the tags match.  C5: 36.9 -> 35.4 us per forward at 1-2 samples; from 8
//     samples on the doubled read volume costs more than the fence (39.6 vs 37.0 us at 8, 58 vs 43 at 9-16): grid barrier there;
//   * the weights are also prefetched into L2 three tiles ahead of the ring (HBM keeps streaming while the ring is full; 39 -> 37 us)
//     and read with an L2 evict-first policy (128 MiB of them would otherwise flush the biases out of L2 every launch, and a bias
//     load that misses L2 waits 2-4 us behind the queued weight reads); biases are requested a layer ahead.
#include "kernels.h"
#include "ptx.cuh"

namespace nc
{

constexpr int MS_CONSUMER_WARPS = 8;
constexpr int MS_THREADS = (MS_CONSUMER_WARPS + 1) * 32;
constexpr int MS_TILE_ROWS = 16;       // weight rows per ring slot = M of the MMA
constexpr int MS_RING_BYTES = 199680;  // 3 slots at fan_in 4096 (pitch 4160); more, smaller slots for narrower layers
constexpr int MS_MAX_SLOTS = 8;
constexpr int MS_MAX_K = 4096;
constexpr int MS_DSTEPS = MS_MAX_K / 64 / MS_CONSUMER_WARPS; // 64-byte K steps (two MMAs) per warp: 8
constexpr int MS_MAX_PARTIAL = 1008;   // (rows per CTA) x (samples, padded to 8 / 16 / 32) int32 per consumer warp: what is left of 227 KB
constexpr int MS_OFF_PARTIAL = MS_RING_BYTES + 64;
constexpr int MS_OFF_BARS = MS_OFF_PARTIAL + MS_CONSUMER_WARPS * MS_MAX_PARTIAL * 4;
constexpr int MS_SMEM = MS_OFF_BARS + 2 * MS_MAX_SLOTS * 8;

// row pitch in the ring: >= fan_in rounded up to 64, and = 64 (mod 128) -- the 16-byte fragment loads of rows r and r + 1 then
// fall into different halves of the 32 banks
__host__ __device__ inline int ms_pitch(int k) { return ((k - 64 + 127) / 128) * 128 + 64; }
__host__ __device__ inline int ms_slots(int k)
{
    const int n = MS_RING_BYTES / (MS_TILE_ROWS * ms_pitch(k));
    return n < MS_MAX_SLOTS ? n : MS_MAX_SLOTS;
}

enum : int
{
    KERR_MS_PRODUCER = 21,
    KERR_MS_CONSUMER = 22,
    KERR_MS_GRID_BARRIER = 23,
};

__device__ __forceinline__ int4 ld_cg_int4(const void *p)
{
    int4 v;
    asm volatile("ld.global.cg.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned *p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned ld_cg_u32(const void *p)
{
    unsigned v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void *p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_v2(void *p, unsigned a, unsigned b)
{
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
// rows of a layer per CTA: a multiple of 4, so that a tagged word (4 neurons of one sample) never straddles two CTAs
__host__ __device__ inline int ms_rows_per_cta(int fan_out, int grid) { return (((fan_out + grid - 1) / grid) + 3) & ~3; }

__device__ __forceinline__ void mma_s8_16832(int *d, const unsigned *a, unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NT: 8-sample groups (1: up to 8 samples, 2: up to 16, 4: up to 32); LL: tagged-word exchange of the hidden activations
template <int NT, bool LL>
__global__ void __launch_bounds__(MS_THREADS, 1)
mlp_i8_stream_kernel(const MlpStreamParams p)
{
    extern __shared__ __align__(128) uint8_t ms_smem[];
    constexpr int BTP = 8 * NT; // padded samples
    const uint32_t base = smem_u32(ms_smem);
    const uint32_t bars = base + MS_OFF_BARS;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (MS_MAX_SLOTS + s); };
    int32_t *partial = reinterpret_cast<int32_t *>(ms_smem + MS_OFF_PARTIAL);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x, grid = gridDim.x;

    auto stamp_raw = [&](int row, int slot) { // (timeline of every CTA: kernel entry / exit)
        if (p.debug && p.debug_cta < 0 && threadIdx.x == 0)
        {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.debug[cta * 32 * 6 + row * 6 + slot] = (long long)t;
        }
    };
    stamp_raw(24, 0);
    if (threadIdx.x == 0)
    {
        for (int s = 0; s < MS_MAX_SLOTS; s++)
        {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), MS_CONSUMER_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    // The ring is re-cut per layer (slot size follows the layer's fan-in).  A slot index and its phase come from one running tile
    // counter per slot COUNT, so both sides must see a layer's tiles on fresh slots: every layer starts at slot 0 with the ring
    // empty in the consumers' view only once they have drained the previous layer -- which the producer does not wait for.  To keep
    // that simple and still prefetch across the barrier, the slot size is fixed for the whole launch: the widest fan-in of the net.
    const int pitch = ms_pitch(p.max_fan_in), slot_bytes = MS_TILE_ROWS * pitch, nslots = ms_slots(p.max_fan_in);

    if (warp == MS_CONSUMER_WARPS)
    {
        // ===================== producer: this CTA's weight rows of every layer, in order =====================
        if (lane == 0)
        {
            uint32_t seq = 0;
            const uint64_t stream_once = l2_policy_evict_first();
            // The ring (~1.5 layers of this CTA's rows) is not deep enough to keep HBM busy through the latency chain of a layer
            // (activation exchange, tiles, partial sums: ~4 us against 2.6 us of streaming), and a ring slot only frees when its
            // tile has been consumed.  So the rows are also prefetched into L2 -- which holds 7 layers of this net -- a few tiles
            // ahead of the ring: HBM streams from the first microsecond on, and the ring fills from L2.
            int pf_l = 0, pf_r = 0, pf_r1 = 0; // prefetch cursor: layer, next row, end of this CTA's rows in that layer
            auto pf_layer = [&]() {
                const int rpc = ms_rows_per_cta(p.layers[pf_l].fan_out, grid);
                pf_r = cta * rpc, pf_r1 = min(p.layers[pf_l].fan_out, pf_r + rpc);
            };
            auto prefetch_tile = [&]() {
                while (pf_l < p.n_layers && pf_r >= pf_r1)
                    if (++pf_l < p.n_layers) pf_layer();
                if (pf_l >= p.n_layers) return;
                const MlpStreamLayer &ly = p.layers[pf_l];
                const int nrows = min(MS_TILE_ROWS, pf_r1 - pf_r);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ly.w + (long long)pf_r * ly.fan_in),
                             "r"((uint32_t)nrows * (uint32_t)ly.fan_in)
                             : "memory"); // (rows of a tile are contiguous in global memory: one request per tile)
                pf_r += MS_TILE_ROWS;
            };
            pf_layer();
            for (int i = 0; i < p.l2_prefetch_tiles; i++) prefetch_tile();
            for (int l = 0; l < p.n_layers; l++)
            {
                const MlpStreamLayer &ly = p.layers[l];
                const int rpc = ms_rows_per_cta(ly.fan_out, grid);
                const int r0 = cta * rpc, r1 = min(ly.fan_out, r0 + rpc);
                for (int r = r0; r < r1; r += MS_TILE_ROWS, seq++)
                {
                    const int slot = seq % nslots;
                    if (p.l2_prefetch_tiles > 0) prefetch_tile();
                    mbar_wait(empty_bar(slot), ((seq / nslots) & 1u) ^ 1u, p.error_flag, KERR_MS_PRODUCER);
                    const int nrows = min(MS_TILE_ROWS, r1 - r);
                    mbar_arrive_expect_tx(full_bar(slot), (uint32_t)nrows * (uint32_t)ly.fan_in);
                    for (int i = 0; i < nrows; i++)
                        bulk_load_1d(base + slot * slot_bytes + i * pitch, ly.w + (long long)(r + i) * ly.fan_in, (uint32_t)ly.fan_in, full_bar(slot), stream_once);
                }
            }
        }
        return;
    }

    // ===================== consumers =====================
    const int gid = lane >> 2, tig = lane & 3;
    auto stamp = [&](int l, int slot) { // optional timeline of CTA `debug_cta` (profiling aid): [layer][6] globaltimer ns
        if (p.debug && (cta == p.debug_cta || p.debug_cta < 0) && threadIdx.x == 0) // (debug_cta < 0: every CTA, [cta][32][6])
        {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.debug[(p.debug_cta < 0 ? cta * 32 * 6 : 0) + l * 6 + slot] = (long long)t;
        }
    };
    // tag of the words layer l reads (= what layer l - 1 wrote): (launch epoch, l).  The epoch lives in device memory and is advanced
    // by the last CTA of every launch, so a replayed CUDA graph still gets fresh tags; the word buffers are zeroed at allocation and
    // no tag is zero.  (A stale word could only pass for a fresh one after exactly 2^27 launches that never rewrote it.)
    const unsigned epoch = LL ? ld_cg_u32(p.barrier + 2) : 0u;
    // Biases: the weight stream keeps ~29 MB of reads queued at the memory controllers (195 KB per SM), so a load that misses L2
    // -- the biases are 16 KB per layer, flushed out of L2 by every launch's 128 MiB of weights -- comes back after 2-4 us.  Asked for
    // between the activation load and the tiles (~1 us ahead) it stalled the finalize of most layers; asked for a layer ahead it is free.
    constexpr int FIN = (MS_MAX_PARTIAL + MS_CONSUMER_WARPS * 32 - 1) / (MS_CONSUMER_WARPS * 32); // plain finalize: values per thread
    int bias_r_next[FIN], bias4_next[4] = {0, 0, 0, 0};
    auto fetch_bias = [&](int l) {
        const MlpStreamLayer &ly = p.layers[l];
        const int rpc = ms_rows_per_cta(ly.fan_out, grid);
        const int r0 = cta * rpc, r1 = min(ly.fan_out, r0 + rpc);
        const int nvals = (r1 - r0) * BTP;
#pragma unroll
        for (int j = 0; j < FIN; j++)
        {
            const int i = threadIdx.x + j * MS_CONSUMER_WARPS * 32;
            bias_r_next[j] = (!(LL && l + 1 < p.n_layers) && i < nvals) ? __ldg(ly.bias + r0 + i / BTP) : 0;
        }
        // (tagged-word finalize: 4 consecutive neurons of one sample per thread)
#pragma unroll
        for (int e = 0; e < 4; e++)
            bias4_next[e] = (LL && l + 1 < p.n_layers && (threadIdx.x / BTP) * 4 < r1 - r0) ? __ldg(ly.bias + r0 + (threadIdx.x / BTP) * 4 + e) : 0;
    };
    fetch_bias(0);
    uint32_t seq = 0;
    for (int l = 0; l < p.n_layers; l++)
    {
        const MlpStreamLayer &ly = p.layers[l];
        const int K = ly.fan_in;
        const int dsteps = (K + 63) >> 6;                                     // 64-byte K steps per row
        const int dw = (dsteps + MS_CONSUMER_WARPS - 1) / MS_CONSUMER_WARPS; // ... per warp (<= MS_DSTEPS)
        const int ds0 = warp * dw, ds1 = min(dsteps, ds0 + dw);
        const int rpc = ms_rows_per_cta(ly.fan_out, grid);
        const int r0 = cta * rpc, r1 = min(ly.fan_out, r0 + rpc);

        stamp(l, 0);
        // the previous layer's outputs of every CTA must be visible
        if (!LL && l > 0)
        {
            if (threadIdx.x == 0)
            {
                const unsigned target = (unsigned)l * (unsigned)grid; // the counter starts every launch at zero (reset below)
                long long t0 = clock64();
                while (ld_acquire_gpu(p.barrier) < target)
                {
                    if (clock64() - t0 > 4000000000LL)
                    {
                        if (p.error_flag) atomicExch(p.error_flag, KERR_MS_GRID_BARRIER);
                        __threadfence_system();
                        __trap();
                    }
                }
            }
            named_bar_sync(1, MS_CONSUMER_WARPS * 32);
        }

        stamp(l, 1);
        // This warp's K slice of the activations as B fragments, all samples, in registers for the whole layer.  The MMA does not
        // care which byte of K sits at which k position as long as A and B agree, so K is PERMUTED to make every load 16 bytes
        // wide: lane (gid, tig) owns bytes [tig*16, +16) of each 64-byte step -- word 0 / 1 are the b0 / b1 (k positions tig*4 and
        // 16 + tig*4) of the step's first MMA, word 2 / 3 those of its second MMA; the weight rows are read the same way below.
        const int8_t *act = l == 0 ? p.in : p.act[l & 1];
        int4 bf[MS_DSTEPS][NT];
        if (!LL || l == 0)
        {
#pragma unroll
            for (int s = 0; s < MS_DSTEPS; s++)
            {
                const int k = (ds0 + s) * 64 + tig * 16;
#pragma unroll
                for (int nt = 0; nt < NT; nt++)
                {
                    const int smp = nt * 8 + gid;
                    const bool ok = ds0 + s < ds1 && smp < p.batch && k < K;
                    bf[s][nt] = ok ? ld_cg_int4(act + (long long)smp * K + k) : make_int4(0, 0, 0, 0);
                }
            }
        }
        else
        {
            // Tagged words: 16 bytes of activations are 32 bytes here, {a0..3, tag, a4..7, tag} {a8..11, tag, a12..15, tag}.  Every lane
            // requests all its units at once and again, only the missing ones, until each carries this layer's tag: the data is in
            // registers one L2 round trip after the last writer's store lands.  (At <= 4 samples a round is <= 32 KB per CTA next
            // to 112 KB of weights per layer; polling ONE unit first and fetching the rest after it -- less polling traffic -- costs
            // a second round trip and measured slower than the grid barrier at every batch size.)
            const unsigned tag = epoch * 32u + (unsigned)l;
            const uint8_t *ll = reinterpret_cast<const uint8_t *>(p.ll[l & 1]);
            const long long t0 = clock64();
            unsigned pending = 0;
#pragma unroll
            for (int s = 0; s < MS_DSTEPS; s++)
#pragma unroll
                for (int nt = 0; nt < NT; nt++)
                {
                    const int k = (ds0 + s) * 64 + tig * 16, smp = nt * 8 + gid;
                    if (ds0 + s < ds1 && smp < p.batch && k < K) pending |= 1u << (s * NT + nt);
                    bf[s][nt] = make_int4(0, 0, 0, 0);
                }
            while (pending)
            {
                uint4 ra[MS_DSTEPS][NT], rb[MS_DSTEPS][NT];
#pragma unroll
                for (int s = 0; s < MS_DSTEPS; s++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++)
                        if (pending >> (s * NT + nt) & 1u)
                        {
                            const int k = (ds0 + s) * 64 + tig * 16, smp = nt * 8 + gid;
                            const uint8_t *u = ll + ((long long)smp * K + k) * 2;
                            ra[s][nt] = ld_volatile_v4(u), rb[s][nt] = ld_volatile_v4(u + 16);
                        }
#pragma unroll
                for (int s = 0; s < MS_DSTEPS; s++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++)
                        if ((pending >> (s * NT + nt) & 1u) && ra[s][nt].y == tag && ra[s][nt].w == tag && rb[s][nt].y == tag && rb[s][nt].w == tag)
                        {
                            bf[s][nt] = make_int4((int)ra[s][nt].x, (int)ra[s][nt].z, (int)rb[s][nt].x, (int)rb[s][nt].z);
                            pending &= ~(1u << (s * NT + nt));
                        }
                if (pending && clock64() - t0 > 4000000000LL)
                {
                    if (p.error_flag) atomicExch(p.error_flag, KERR_MS_GRID_BARRIER);
                    __threadfence_system();
                    __trap();
                }
            }
        }
        if (bf[0][0].x == 0x12345678) stamp(l, 5); // (keeps the loads above ahead of the next stamp)
        stamp(l, 2);
        // biases of the outputs this thread finalises below: requested one layer AHEAD (see fetch_bias)
        const int nvals = (r1 - r0) * BTP;
        int bias_r[FIN], bias4[4];
#pragma unroll
        for (int j = 0; j < FIN; j++) bias_r[j] = bias_r_next[j];
#pragma unroll
        for (int e = 0; e < 4; e++) bias4[e] = bias4_next[e];
        if (l + 1 < p.n_layers) fetch_bias(l + 1);

        int32_t *my_partial = partial + warp * MS_MAX_PARTIAL;
        for (int r = r0; r < r1; r += MS_TILE_ROWS, seq++)
        {
            const int slot = seq % nslots;
            mbar_wait(full_bar(slot), (seq / nslots) & 1u, p.error_flag, KERR_MS_CONSUMER);
            // A fragments: 16 bytes of row gid and of row gid + 8 per 64-byte step, the same K permutation as above
            // (rows past the CTA's slice and 16-byte chunks past a row's end are stale ring content: they are never stored, or
            //  meet zero activations)
            const uint8_t *w = ms_smem + slot * slot_bytes + gid * pitch + ds0 * 64 + tig * 16;
            constexpr int CH = NT >= 4 ? 1 : 2; // independent accumulation chains per sample group (four groups are chains enough)
            int acc[2][NT][4];
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
                for (int nt = 0; nt < NT; nt++) acc[c][nt][0] = acc[c][nt][1] = acc[c][nt][2] = acc[c][nt][3] = 0;
#pragma unroll
            for (int s = 0; s < MS_DSTEPS; s++)
            {
                if (ds0 + s < ds1) // warp-uniform
                {
                    const int4 lo = *reinterpret_cast<const int4 *>(w + s * 64);
                    const int4 hi = *reinterpret_cast<const int4 *>(w + s * 64 + 8 * pitch);
                    const unsigned a1[4] = {(unsigned)lo.x, (unsigned)hi.x, (unsigned)lo.y, (unsigned)hi.y};
                    const unsigned a2[4] = {(unsigned)lo.z, (unsigned)hi.z, (unsigned)lo.w, (unsigned)hi.w};
#pragma unroll
                    for (int nt = 0; nt < NT; nt++)
                    {
                        mma_s8_16832(acc[0][nt], a1, (unsigned)bf[s][nt].x, (unsigned)bf[s][nt].y);
                        mma_s8_16832(acc[CH - 1][nt], a2, (unsigned)bf[s][nt].z, (unsigned)bf[s][nt].w);
                    }
                }
            }
            // accumulator fragment: c0 / c1 = row gid, samples tig*2, +1; c2 / c3 = row gid + 8
            int32_t *dst = my_partial + (r - r0) * BTP;
            const int rows_left = r1 - r; // (rows of the tile past the CTA's slice are stale ring content: not stored)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
            {
                if (gid < rows_left)
                    *reinterpret_cast<int2 *>(dst + gid * BTP + nt * 8 + tig * 2) = make_int2(acc[0][nt][0] + acc[1][nt][0], acc[0][nt][1] + acc[1][nt][1]);
                if (gid + 8 < rows_left)
                    *reinterpret_cast<int2 *>(dst + (gid + 8) * BTP + nt * 8 + tig * 2) = make_int2(acc[0][nt][2] + acc[1][nt][2], acc[0][nt][3] + acc[1][nt][3]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar(slot));
        }

        // ---- this CTA's outputs of the layer: sum of the 8 K slices + bias, activation, requantisation ----
        stamp(l, 3);
        named_bar_sync(1, MS_CONSUMER_WARPS * 32);
        stamp(l, 4);
        const bool last = l + 1 == p.n_layers;
        const bool relu = (p.relu_mask >> l) & 1u;
        if (LL && !last)
        {
            // one thread = 4 consecutive neurons of one sample = one tagged word (rows per CTA and hidden widths are multiples of 4)
            const int i = threadIdx.x, g = i / BTP, b = i - g * BTP;
            if (g * 4 < r1 - r0 && b < p.batch)
            {
                unsigned word = 0;
#pragma unroll
                for (int e = 0; e < 4; e++)
                {
                    const int row = g * 4 + e;
                    int v = bias4[e];
#pragma unroll
                    for (int ww = 0; ww < MS_CONSUMER_WARPS; ww++) v += partial[ww * MS_MAX_PARTIAL + row * BTP + b];
                    if (relu) v = max(v, 0);
                    word |= ((unsigned)min(127, max(-128, v >> 7)) & 0xFFu) << (8 * e);
                }
                stamp(16 + l, 0);
                st_volatile_v2(reinterpret_cast<uint8_t *>(p.ll[(l + 1) & 1]) + ((long long)b * ly.fan_out + r0 + g * 4) * 2, word,
                               epoch * 32u + (unsigned)(l + 1));
                stamp(16 + l, 1);
            }
            named_bar_sync(1, MS_CONSUMER_WARPS * 32); // `partial` is rewritten by the next layer's tiles
            stamp(l, 5);
            continue;
        }
#pragma unroll
        for (int j = 0; j < FIN; j++)
        {
            const int i = threadIdx.x + j * MS_CONSUMER_WARPS * 32;
            const int row = i / BTP, b = i - row * BTP;
            if (i >= nvals || b >= p.batch) continue;
            int v = bias_r[j];
#pragma unroll
            for (int ww = 0; ww < MS_CONSUMER_WARPS; ww++) v += partial[ww * MS_MAX_PARTIAL + i];
            if (relu) v = max(v, 0);
            if (last)
                p.out[(long long)b * ly.fan_out + r0 + row] = v;
            else
                p.act[(l + 1) & 1][(long long)b * ly.fan_out + r0 + row] = (int8_t)min(127, max(-128, v >> 7));
        }
        if (!last)
        {
            // release: the barrier orders every consumer thread's stores before thread 0's release-add (cumulativity)
            stamp(16 + l, 0);
            named_bar_sync(1, MS_CONSUMER_WARPS * 32);
            stamp(16 + l, 1);
            if (threadIdx.x == 0)
            {
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                stamp(16 + l, 2);
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p.barrier), "r"(1u) : "memory");
            }
        }
        stamp(l, 5);
    }
    // The last CTA to get here leaves both counters at zero for the next launch: nobody polls `barrier` any more (every CTA is
    // past its last wait), and launches of one handle are stream-ordered.  No host-side state, so the launch can sit in a CUDA graph.
    stamp_raw(24, 1);
    if (threadIdx.x == 0 && atomicAdd(p.barrier + 1, 1u) == (unsigned)grid - 1u)
    {
        p.barrier[0] = 0u;
        p.barrier[1] = 0u;
        if (LL) p.barrier[2] = epoch + 1u; // (every CTA read the epoch before it arrived here)
    }
}

template <int NT, bool LL>
static cudaError_t launch_one(const MlpStreamParams &p, int grid, cudaStream_t stream)
{
    static bool opted[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !opted[dev])
    {
        cudaError_t e = cudaFuncSetAttribute(mlp_i8_stream_kernel<NT, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, MS_SMEM);
        if (e != cudaSuccess) return e;
        opted[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3(MS_THREADS), cfg.dynamicSmemBytes = MS_SMEM, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; // all CTAs co-resident: the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, mlp_i8_stream_kernel<NT, LL>, p);
}

bool mlp_stream_supported(const MlpStreamParams &p, int grid)
{
    if (p.n_layers < 1 || p.n_layers > MLP_STREAM_MAX_LAYERS || p.batch < 1 || p.batch > MLP_STREAM_MAX_BATCH || grid < 1) return false;
    const int btp = p.batch <= 8 ? 8 : p.batch <= 16 ? 16 : 32;
    int max_k = 0;
    for (int l = 0; l < p.n_layers; l++)
    {
        const MlpStreamLayer &ly = p.layers[l];
        if (ly.fan_in < 16 || (ly.fan_in & 15) || ly.fan_in > MS_MAX_K || ly.fan_out < 1) return false;
        const int rpc = ms_rows_per_cta(ly.fan_out, grid);
        if (rpc * btp > MS_MAX_PARTIAL) return false;
        if ((reinterpret_cast<uintptr_t>(ly.w) & 15u) != 0) return false;
        max_k = ly.fan_in > max_k ? ly.fan_in : max_k;
    }
    return p.max_fan_in == max_k;
}

cudaError_t launch_mlp_i8_stream(const MlpStreamParams &p, int grid, cudaStream_t stream)
{
    if (!mlp_stream_supported(p, grid)) return cudaErrorInvalidValue;
    const bool ll = p.ll[0] && p.ll[1] && p.batch <= MLP_STREAM_LL_MAX_BATCH;
    if (p.batch <= 8) return ll ? launch_one<1, true>(p, grid, stream) : launch_one<1, false>(p, grid, stream);
    return p.batch <= 16 ? launch_one<2, false>(p, grid, stream) : launch_one<4, false>(p, grid, stream);
}

} // namespace nc

// gemm_tcgen05.cuh -- the dense hot kernel:  out[M][N] = epilogue(A[M][K] . W[N][K]^T + bias[N])
//
// This is the B200 replacement for the arithmetic of the reference's `network_v1` task
// (src/netFPGA.cpp:250,275: one fully-connected layer after another over W[out][in] matrices laid
// out as in src/netFPGA.cpp:91-106) and for every linear layer of the ViT.  Both operands are
// K-major, exactly the reference's row-major W[out][in] and sample-major activations, so no
// transposes are needed anywhere.
//
// CG = 2 (the large-problem configuration): two CTAs on neighbouring SMs form a cluster and share one
// 256 x BN accumulator tile (tcgen05 cta_group::2).  Each CTA loads its own 128 rows of A and HALF of the
// W tile, so per SM a k-block costs 32 KB of L2->smem traffic instead of 48 KB and the ring holds 6 stages;
// the leader CTA's thread issues the 256-row UMMAs for the pair, both CTAs run their own epilogue.
// CG = 1: one CTA per tile (small M).
//
// Structure (one persistent CTA per SM, 384 threads, warp-specialised):
//   warp 8     TMA producer: 128B-swizzled [128 x 128B] A tile + [BN x 128B] W tile per stage
//   warp 9     MMA issuer:   one thread issues tcgen05.mma (UMMA 128 x BN x 32B) into TMEM
//   warp 10    TMEM allocator (2 accumulator stages x BN fp32/int32 columns)
//   warps 0-7  epilogue: warp w reads TMEM lanes [32(w%4), +32) x columns [w/4 * BN/2, +BN/2):
//              tcgen05.ld -> +bias -> activation -> convert -> 128B-swizzled smem slab -> one TMA
//              store (or TMA reduce-add for the residual epilogue) per 32-row x 128-byte slab.
//              Runs concurrently with the next tile's MMAs (double-buffered accumulator).
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue).
// All byte geometry is independent of the operand type: a stage row is always 128 bytes
// (64 bf16 / 32 tf32 / 128 int8), one UMMA consumes 32 bytes of K.
//
// Outputs that a TMA store cannot address (row pitch not a multiple of 16 bytes, or the row-remapping
// patch-embedding epilogue) take the "direct" epilogue: each lane writes its own accumulator row.
#pragma once

#include "ptx.cuh"

namespace nc
{

enum : int
{
    OUT_F32 = 0,
    OUT_BF16 = 1,
    OUT_S8 = 2,
    OUT_S32 = 3
};
enum : int
{
    EPI_NONE = 0,
    EPI_RELU = 1,
    EPI_GELU = 2,
    EPI_RESIDUAL = 3, // out(fp32) += acc + bias
    EPI_REQUANT = 4,  // int8: q = clamp(relu?(acc + bias) >> 7)
    EPI_REQUANT_RELU = 5,
    EPI_PATCH = 6, // fp32 out with row remap + position embedding (patch embedding)
    EPI_SPLITK = 7, // int32 out: this CTA's K slice is ADDED to the output (TMA reduce-add), no bias -- see GemmParams::k_splits
    EPI_GELU_X2 = 8 // EPI_GELU evaluated two elements at a time on the packed fp32 pipe (same bits; chosen by the launcher)
};

struct GemmParams
{
    int M, N, K;       // K in elements of the operand type
    const void *bias;  // float[N] (bf16 / tf32) or int32[N] (int8); may be null
    void *out;         // OUT_* typed, row pitch ldc elements
    long long ldc;
    int epi;
    int tma_store; // 1: epilogue through smem slabs + TMA store (tma_out valid); 0: direct stores
    // Split-K (integer kinds, few tiles, long K: weight streaming at small batch): every tile is computed by k_splits CTAs,
    // each over a contiguous range of k-blocks, and the partial sums meet in the int32 output through TMA reduce-adds
    // (exact and order-independent for integers).  The caller zeroes the output first and applies bias / requantisation after.
    int k_splits;
    // EPI_PATCH: input row r = b * remap_in + t  ->  output row b * remap_out + 1 + t, and
    // pos[(1 + t) * N + col] is added (cls token occupies output row b * remap_out).
    int remap_in, remap_out;
    const float *pos;
    int *error_flag;
    long long *debug; // optional clock64 stamps of cluster 0: [tile < 40][warp 0..11 of CTA 0, 12..23 of CTA 1][4] (profiling aid)
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_STAGE_ROW_BYTES = 128;
// EW = number of epilogue warps (8 or 16): warp w serves TMEM lanes [32 (w % 4), +32) and columns
// [(w - 4) / 4 * BN / (EW / 4), + BN / (EW / 4)).  16 warps (4 per scheduler) are for epilogues whose per-element work is
// heavy (exact-erf GELU: 14 issue slots) so that the epilogue keeps pace with the MMAs; they cost 32 KB more smem (one stage).
__host__ __device__ constexpr int gemm_threads(int ew) { return 128 + 32 * ew; }
constexpr int GEMM_SLAB_BYTES = 32 * 128; // 32 rows x 128 bytes per epilogue warp

template <int OUT>
struct OutTraits;
template <>
struct OutTraits<OUT_F32>
{
    static constexpr int ELEM = 4;
};
template <>
struct OutTraits<OUT_S32>
{
    static constexpr int ELEM = 4;
};
template <>
struct OutTraits<OUT_BF16>
{
    static constexpr int ELEM = 2;
};
template <>
struct OutTraits<OUT_S8>
{
    static constexpr int ELEM = 1;
};

// Columns of the output covered by one TMA-store slab of an epilogue warp (host and device agree on this):
// a 128-byte row, or the warp's whole column range when that is narrower than 128 bytes.
template <int BN, int OUT, int EW>
__host__ __device__ constexpr int slab_cols()
{
    return (128 / OutTraits<OUT>::ELEM) < (BN / (EW / 4)) ? (128 / OutTraits<OUT>::ELEM) : (BN / (EW / 4));
}

template <int BN, int STAGES, int CG, int EW, int DS = 0>
struct GemmSmem
{
    static constexpr int A_BYTES = GEMM_BM * GEMM_STAGE_ROW_BYTES;
    static constexpr int B_BYTES = (BN / CG) * GEMM_STAGE_ROW_BYTES; // a CTA of a pair stages half of the W tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int OFF_SLABS = STAGES * STAGE_BYTES; // 1024-byte aligned (stage sizes are multiples of 1024)
    static constexpr int SLABS_PER_WARP = DS ? 2 : 1; // DS: slab i + 1 is filled while the TMA store of slab i drains
    static constexpr int OFF_BIAS = OFF_SLABS + EW * SLABS_PER_WARP * GEMM_SLAB_BYTES;
    static constexpr int OFF_BARS = OFF_BIAS + 2 * BN * 4; // bias slice, double-buffered by tile parity
    static constexpr int NUM_BARS = 2 * STAGES + 4;
    static constexpr int OFF_TMEM_PTR = OFF_BARS + NUM_BARS * 8;
    static constexpr int TOTAL = OFF_TMEM_PTR + 16; // the dynamic smem window itself is 1024-byte aligned (checked in the kernel)
    static_assert(TOTAL <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
};

template <int KIND>
struct KindTraits;
template <>
struct KindTraits<KIND_BF16>
{
    static constexpr int ELEM = 2;
    __host__ __device__ static constexpr uint32_t idesc(int m, int bn) { return umma_idesc(1, 1, m, bn); }
};
template <>
struct KindTraits<KIND_TF32>
{
    static constexpr int ELEM = 4;
    __host__ __device__ static constexpr uint32_t idesc(int m, int bn) { return umma_idesc(1, 2, m, bn); }
};
template <>
struct KindTraits<KIND_I8>
{
    static constexpr int ELEM = 1;
    __host__ __device__ static constexpr uint32_t idesc(int m, int bn) { return umma_idesc(2, 1, m, bn); }
};

// ---- epilogue helpers --------------------------------------------------------------------------

__device__ __forceinline__ float epi_act_f32(float v, int epi)
{
    if (epi == EPI_RELU) return fmaxf(v, 0.0f);
    if (epi == EPI_GELU || epi == EPI_GELU_X2) return gelu_erf(v); // (the packed form lives in epi_convert_chunk)
    return v;
}

// 32 accumulator columns of one row (+ bias, activation / requantisation) -> 32 output elements packed
// into `w` (32 words for 4-byte outputs, 16 for bf16, 8 for int8).  `bias` points at 32 words in smem.
template <int OUT>
__device__ __forceinline__ void epi_convert32(const uint32_t *v, const uint32_t *bias, int epi, uint32_t *w)
{
    if constexpr (OUT == OUT_F32)
    {
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++)
        {
            const uint4 b = *reinterpret_cast<const uint4 *>(bias + 4 * j4);
            w[4 * j4 + 0] = __float_as_uint(epi_act_f32(__uint_as_float(v[4 * j4 + 0]) + __uint_as_float(b.x), epi));
            w[4 * j4 + 1] = __float_as_uint(epi_act_f32(__uint_as_float(v[4 * j4 + 1]) + __uint_as_float(b.y), epi));
            w[4 * j4 + 2] = __float_as_uint(epi_act_f32(__uint_as_float(v[4 * j4 + 2]) + __uint_as_float(b.z), epi));
            w[4 * j4 + 3] = __float_as_uint(epi_act_f32(__uint_as_float(v[4 * j4 + 3]) + __uint_as_float(b.w), epi));
        }
    }
    else if constexpr (OUT == OUT_S32)
    {
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++)
        {
            const uint4 b = *reinterpret_cast<const uint4 *>(bias + 4 * j4);
            const uint32_t bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 4; e++)
            {
                int a = (int)v[4 * j4 + e] + (int)bb[e];
                if (epi == EPI_RELU) a = max(a, 0);
                w[4 * j4 + e] = (uint32_t)a;
            }
        }
    }
    else if constexpr (OUT == OUT_BF16)
    {
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++)
        {
            const uint4 b = *reinterpret_cast<const uint4 *>(bias + 4 * j4);
            const float f0 = epi_act_f32(__uint_as_float(v[4 * j4 + 0]) + __uint_as_float(b.x), epi);
            const float f1 = epi_act_f32(__uint_as_float(v[4 * j4 + 1]) + __uint_as_float(b.y), epi);
            const float f2 = epi_act_f32(__uint_as_float(v[4 * j4 + 2]) + __uint_as_float(b.z), epi);
            const float f3 = epi_act_f32(__uint_as_float(v[4 * j4 + 3]) + __uint_as_float(b.w), epi);
            w[2 * j4 + 0] = pack_bf16x2(f0, f1);
            w[2 * j4 + 1] = pack_bf16x2(f2, f3);
        }
    }
    else
    {
        // int8 requantisation: q = clamp((relu?)(acc + bias) >> 7, -128, 127)
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++)
        {
            const uint4 b = *reinterpret_cast<const uint4 *>(bias + 4 * j4);
            const uint32_t bb[4] = {b.x, b.y, b.z, b.w};
            uint32_t word = 0;
#pragma unroll
            for (int e = 0; e < 4; e++)
            {
                int a = (int)v[4 * j4 + e] + (int)bb[e];
                if (epi == EPI_REQUANT_RELU) a = max(a, 0);
                a = min(127, max(-128, a >> 7));
                word |= ((uint32_t)a & 0xFFu) << (8 * e);
            }
            w[j4] = word;
        }
    }
}

// The same for the 16 / sizeof(out) columns that make one 16-byte chunk of an output row (8 bf16, 4 fp32 / int32, 16 int8):
// `v` holds that many accumulator columns, `b` the matching bias words (already in registers: the caller reads them from smem
// before it waits for the accumulator columns, so their latency hides behind the TMEM load); returns the 4 packed words.
template <int N>
__device__ __forceinline__ void load_bias_chunk(const uint32_t *bias, uint32_t (&b)[N])
{
#pragma unroll
    for (int i = 0; i < N / 4; i++)
    {
        const uint4 t = *reinterpret_cast<const uint4 *>(bias + 4 * i);
        b[4 * i] = t.x, b[4 * i + 1] = t.y, b[4 * i + 2] = t.z, b[4 * i + 3] = t.w;
    }
}

template <int OUT>
__device__ __forceinline__ uint4 epi_convert_chunk(const uint32_t *v, const uint32_t *b, int epi)
{
    uint4 w;
    if constexpr (OUT == OUT_F32)
    {
        w.x = __float_as_uint(epi_act_f32(__uint_as_float(v[0]) + __uint_as_float(b[0]), epi));
        w.y = __float_as_uint(epi_act_f32(__uint_as_float(v[1]) + __uint_as_float(b[1]), epi));
        w.z = __float_as_uint(epi_act_f32(__uint_as_float(v[2]) + __uint_as_float(b[2]), epi));
        w.w = __float_as_uint(epi_act_f32(__uint_as_float(v[3]) + __uint_as_float(b[3]), epi));
    }
    else if constexpr (OUT == OUT_S32)
    {
        int a0 = (int)v[0] + (int)b[0], a1 = (int)v[1] + (int)b[1], a2 = (int)v[2] + (int)b[2], a3 = (int)v[3] + (int)b[3];
        if (epi == EPI_RELU) a0 = max(a0, 0), a1 = max(a1, 0), a2 = max(a2, 0), a3 = max(a3, 0);
        w = make_uint4((uint32_t)a0, (uint32_t)a1, (uint32_t)a2, (uint32_t)a3);
    }
    else if constexpr (OUT == OUT_BF16)
    {
        float f[8];
        if (epi == EPI_GELU_X2)
        {
#pragma unroll
            for (int e = 0; e < 8; e += 2)
                gelu_erf_x2(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(b[e]), __uint_as_float(b[e + 1]), f[e], f[e + 1]);
        }
        else
        {
#pragma unroll
            for (int e = 0; e < 8; e++) f[e] = epi_act_f32(__uint_as_float(v[e]) + __uint_as_float(b[e]), epi);
        }
        w = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
    else
    {
        uint32_t word[4];
#pragma unroll
        for (int j4 = 0; j4 < 4; j4++)
        {
            word[j4] = 0;
#pragma unroll
            for (int e = 0; e < 4; e++)
            {
                int a = (int)v[4 * j4 + e] + (int)b[4 * j4 + e];
                if (epi == EPI_REQUANT_RELU) a = max(a, 0);
                a = min(127, max(-128, a >> 7));
                word[j4] |= ((uint32_t)a & 0xFFu) << (8 * e);
            }
        }
        w = make_uint4(word[0], word[1], word[2], word[3]);
    }
    return w;
}

template <int KIND, int BN, int OUT, int STAGES, int CG, int EW, int DS = 0>
__global__ void __launch_bounds__(gemm_threads(EW), 1)
gemm_tn_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w,
                       const __grid_constant__ CUtensorMap tma_out, const GemmParams p)
{
    using L = GemmSmem<BN, STAGES, CG, EW, DS>;
    constexpr int ELEM = KindTraits<KIND>::ELEM;
    constexpr int BK = GEMM_STAGE_ROW_BYTES / ELEM; // elements of K per stage
    constexpr uint32_t TMEM_COLS = 2 * BN;          // two accumulator stages (power of two >= 32)
    constexpr uint32_t IDESC = KindTraits<KIND>::idesc(GEMM_BM * CG, BN);
    constexpr int TILE_M = GEMM_BM * CG;            // rows of one accumulator tile (per CTA pair when CG = 2)
    static_assert(CG == 1 || CG == 2, "a tile belongs to one CTA or to a CTA pair");
    const int epi = p.epi; // (a compile-time epilogue was measured: same speed, so one instantiation serves all epilogues of a type)
    // Warp roles.  The epilogue warps take the LOW warp ids: the scheduler favours the highest warp id among ready warps,
    // and the single MMA-issuing thread must never queue behind epilogue arithmetic (measured on the GELU GEMM: the issue
    // loop of a tile took 8.5 k cycles instead of 6.1 k when the issuer was warp 1 and the epilogue warps 4..11).
    constexpr int W_PRODUCER = EW, W_MMA = EW + 1, W_ALLOC = EW + 2;
    constexpr int OELEM = OutTraits<OUT>::ELEM;
    constexpr int WARP_COLS = BN / (EW / 4);        // columns per epilogue warp
    constexpr int SLAB_COLS = slab_cols<BN, OUT, EW>(); // columns per TMA-store slab
    constexpr int SLAB_ROW_BYTES = SLAB_COLS * OELEM;
    constexpr bool SLAB_SWIZZLED = SLAB_ROW_BYTES == 128; // narrower slab rows (32 / 64 bytes) are stored unswizzled
    static_assert(BN == 128 || BN == 256, "BN must be 128 or 256");
    static_assert(EW == 8 || EW == 16, "8 or 16 epilogue warps");
    static_assert(SLAB_ROW_BYTES == 128 || SLAB_ROW_BYTES == 64 || SLAB_ROW_BYTES == 32, "slab rows are 128, 64 or 32 bytes");
    static_assert(WARP_COLS % 32 == 0, "an epilogue warp handles whole 32-column TMEM groups");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw); // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t *smem = smem_raw;
    if ((base & 1023u) != 0)
    {
        if (threadIdx.x == 0 && p.error_flag) atomicExch(p.error_flag, KERR_SMEM_ALIGN);
        return; // uniform: every thread of every CTA sees the same window offset
    }

    const uint32_t bars = base + L::OFF_BARS;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(smem + L::OFF_TMEM_PTR);

    const int warp = uniform_warp_idx(); // (uniform: the producer's and the issuer's operands stay in uniform registers, see ptx.cuh)
    const int lane = threadIdx.x & 31;

    const int tiles_m = (p.M + TILE_M - 1) / TILE_M;
    const int tiles_n = (p.N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = (p.K + BK - 1) / BK;
    const int S = p.k_splits > 1 ? p.k_splits : 1;
    const int num_work = num_tiles * S; // work item = (tile, K slice); K slice fastest
    const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0; // 0 = leader of the pair
    const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG; // both CTAs of a pair walk the same tiles

    if (warp == W_PRODUCER && lane == 0)
    {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_w);
        if (p.tma_store) tma_prefetch_desc(&tma_out);
    }
    if (warp == W_MMA && lane == 0)
    {
        for (int s = 0; s < STAGES; s++)
        {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; a++)
        {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), EW * CG); // one arrive per epilogue warp (of both CTAs of a pair)
        }
        fence_barrier_init();
    }
    if (warp == W_ALLOC)
    {
        if constexpr (CG == 2)
        {
            tmem_alloc_pair(base + L::OFF_TMEM_PTR, TMEM_COLS);
            tmem_relinquish_pair();
        }
        else
        {
            tmem_alloc(base + L::OFF_TMEM_PTR, TMEM_COLS);
            tmem_relinquish();
        }
    }
    tcgen05_fence_before();
    if constexpr (CG == 2)
        cluster_sync_all(); // the peer's barriers must be initialised before anything signals them
    else
        __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // everything above overlapped the previous kernel's tail; from here on its outputs are read (and its inputs overwritten)
    griddep_launch_dependents();
    griddep_wait();

    if (warp >= EW)
    {
    // 16 epilogue warps: 640 threads x 96 registers at launch; the four non-epilogue warps hand back 56 each, which lets
    // every epilogue thread grow to 104 (128 x 40 + 512 x 104 <= 640 x 96: a larger request would block forever)
    if constexpr (EW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == W_PRODUCER)
    {
        // ===================== TMA producer =====================
        if (lane == 0)
        {
            uint32_t stage = 0, phase = 0;
            for (int work = tile0; work < num_work; work += tile_step)
            {
                // n-fastest tile order: the CTAs of one wave share a few A row-blocks and all of W in L2
                const int tile = work / S, ks = work - tile * S;
                const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
                const int kb0 = ks * num_kb / S, kb1 = (ks + 1) * num_kb / S;
                for (int kb = kb0; kb < kb1; kb++)
                {
                    mbar_wait(empty_bar(stage), phase ^ 1u, p.error_flag, KERR_PRODUCER_EMPTY);
                    const uint32_t a_dst = base + stage * L::STAGE_BYTES;
                    if constexpr (CG == 2)
                    {
                        // the leader's barrier collects the bytes of both CTAs' loads of this stage
                        if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);
                        tma_load_2d_pair(a_dst, &tma_a, full_bar(stage), kb * BK, m_blk * TILE_M + cta_rank * GEMM_BM);
                        tma_load_2d_pair(a_dst + L::A_BYTES, &tma_w, full_bar(stage), kb * BK, n_blk * BN + cta_rank * (BN / 2));
                    }
                    else
                    {
                        mbar_arrive_expect_tx(full_bar(stage), L::STAGE_BYTES);
                        tma_load_2d(a_dst, &tma_a, full_bar(stage), kb * BK, m_blk * GEMM_BM);
                        tma_load_2d(a_dst + L::A_BYTES, &tma_w, full_bar(stage), kb * BK, n_blk * BN);
                    }
                    if (++stage == STAGES)
                    {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    }
    else if (warp == W_MMA)
    {
        // ===================== MMA issuer =====================
        if (lane == 0 && cta_rank == 0)
        {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int work = tile0; work < num_work; work += tile_step)
            {
                const int ks = work % S, kb0 = ks * num_kb / S, kb1 = (ks + 1) * num_kb / S;
                int tl = (work - tile0) / tile_step;
                const bool dbg = p.debug != nullptr && tile0 == 0 && tl < 40;
                if (dbg) p.debug[(tl * 24 + 11) * 4 + 0] = clock64();
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.error_flag, KERR_MMA_TMEM_EMPTY);
                tcgen05_fence_after();
                if (dbg) p.debug[(tl * 24 + 11) * 4 + 1] = clock64();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; kb++)
                {
                    mbar_wait(full_bar(stage), phase, p.error_flag, KERR_MMA_FULL);
                    tcgen05_fence_after();
                    const uint32_t a_addr = base + stage * L::STAGE_BYTES;
                    const uint64_t a_desc = umma_smem_desc_sw128(a_addr);
                    const uint64_t b_desc = umma_smem_desc_sw128(a_addr + L::A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; k++) // 4 x 32 bytes of K per stage; +2 = 32 B >> 4
                    {
                        if constexpr (CG == 2)
                            umma_ss_pair<KIND>(d_tmem, a_desc + 2u * k, b_desc + 2u * k, IDESC, (kb != kb0 || k != 0) ? 1u : 0u);
                        else
                            umma_ss<KIND>(d_tmem, a_desc + 2u * k, b_desc + 2u * k, IDESC, (kb != kb0 || k != 0) ? 1u : 0u);
                    }
                    // frees the smem slot (in both CTAs of a pair) once these MMAs retire
                    if constexpr (CG == 2)
                        tcgen05_commit_pair(empty_bar(stage));
                    else
                        tcgen05_commit(empty_bar(stage));
                    if (++stage == STAGES)
                    {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                // accumulator complete -> epilogue warps (of both CTAs of a pair)
                if constexpr (CG == 2)
                    tcgen05_commit_pair(tfull_bar(acc));
                else
                    tcgen05_commit(tfull_bar(acc));
                if (dbg) p.debug[(tl * 24 + 11) * 4 + 2] = clock64();
                if (++acc == 2)
                {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    }
    }
    else
    {
        // ===================== epilogue =====================
        if constexpr (EW == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int ew = warp;              // 0..EW-1
        const int q = warp & 3;           // TMEM lane quarter this warp may read: lanes [32q, 32q+32)
        const int wcol = (ew >> 2) * WARP_COLS; // first tile column of this warp
        uint32_t *bias_all = reinterpret_cast<uint32_t *>(smem + L::OFF_BIAS);
        uint8_t *slab = smem + L::OFF_SLABS + ew * L::SLABS_PER_WARP * GEMM_SLAB_BYTES;
        const uint32_t slab_addr = base + L::OFF_SLABS + ew * L::SLABS_PER_WARP * GEMM_SLAB_BYTES;
        [[maybe_unused]] uint32_t slab_seq = 0; // DS: slabs stored so far (parity = buffer)
        uint32_t acc = 0, acc_phase = 0, parity = 0;
        for (int work = tile0; work < num_work; work += tile_step, parity ^= 1u)
        {
            const int tile = work / S;
            const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
            const int col0 = n_blk * BN;
            const int row0 = m_blk * TILE_M + cta_rank * GEMM_BM + q * 32;

            // This tile's bias slice (bit pattern: float or int32).  Every warp fetches the WARP_COLS values of its own column range
            // into registers before it waits for the accumulator and parks them in smem afterwards (the four warps of a column
            // range write identical words).  No barrier among the epilogue warps: once tfull(t) has completed, every warp has
            // handed back the accumulator of tile t - 2, i.e. is done reading the slice of the same parity.
            uint32_t *bias_s = bias_all + parity * BN;
            uint32_t bias_r[WARP_COLS / 32];
#pragma unroll
            for (int i = 0; i < WARP_COLS / 32; i++)
            {
                const int c = col0 + wcol + i * 32 + lane;
                bias_r[i] = (p.bias != nullptr && c < p.N) ? __ldg(reinterpret_cast<const uint32_t *>(p.bias) + c) : 0u;
            }

            const int tl = (work - tile0) / tile_step;
            const bool dbg = p.debug != nullptr && tile0 == 0 && tl < 40 && lane == 0 && warp < 11;
            long long *dslot = p.debug + (tl * 24 + cta_rank * 12 + warp) * 4;
            if (dbg) dslot[0] = clock64();
            mbar_wait(tfull_bar(acc), acc_phase, p.error_flag, KERR_EPI_TMEM_FULL);
            tcgen05_fence_after();
            if (dbg) dslot[1] = clock64();
#pragma unroll
            for (int i = 0; i < WARP_COLS / 32; i++) bias_s[wcol + i * 32 + lane] = bias_r[i];
            __syncwarp();
            const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + wcol;

            if (p.tma_store)
            {
                // ---- smem slab + TMA store: 32 rows x SLAB_COLS columns per store ----
                // A slab row is produced 16 bytes at a time in a ROLLED loop (the next chunk's accumulator columns are in
                // flight while the current chunk is converted).  Keeping this loop small matters more than unrolling it: with the
                // 32-column groups fully unrolled the GELU epilogue was ~25 KB of code that streamed through the 6 KB L0
                // instruction cache of every scheduler and evicted the MMA issuer's loop (fc1 lost 15 % to that).
                constexpr int CHUNK_COLS = 16 / OELEM;              // accumulator columns per 16-byte output chunk
                constexpr int CHUNKS_PER_ROW = SLAB_ROW_BYTES / 16; // 8 (4 / 2 for the narrow slabs); always even
                for (int sb = 0; sb < WARP_COLS / SLAB_COLS; sb++)
                {
                    const int scol = wcol + sb * SLAB_COLS; // tile column of this slab
                    if (col0 + scol >= p.N) break;          // warp-uniform
                    // the previous store of this warp out of this slab buffer must have finished reading it
                    const uint32_t buf_off = DS ? (slab_seq & 1u) * GEMM_SLAB_BYTES : 0u;
                    slab_seq++;
                    if (lane == 0)
                    {
                        if constexpr (DS)
                            tma_store_wait_read_1();
                        else
                            tma_store_wait_read();
                    }
                    __syncwarp();
                    const uint32_t t_slab = t_base + sb * SLAB_COLS;
                    const uint32_t *bias_slab = bias_s + scol;
                    uint8_t *row = slab + buf_off + lane * SLAB_ROW_BYTES;
                    uint32_t v0[CHUNK_COLS], v1[CHUNK_COLS], b0[CHUNK_COLS], b1[CHUNK_COLS];
                    tmem_ld_32xN<CHUNK_COLS>(t_slab, v0);
#pragma unroll 1
                    for (int j = 0; j < CHUNKS_PER_ROW; j += 2)
                    {
                        load_bias_chunk<CHUNK_COLS>(bias_slab + j * CHUNK_COLS, b0);
                        load_bias_chunk<CHUNK_COLS>(bias_slab + (j + 1) * CHUNK_COLS, b1);
                        tmem_ld_wait();
                        tmem_ld_32xN<CHUNK_COLS>(t_slab + (j + 1) * CHUNK_COLS, v1);
                        // lane = row; 16-byte chunk j of the row goes to chunk (j ^ (row & 7)) (128B swizzle)
                        *reinterpret_cast<uint4 *>(row + ((SLAB_SWIZZLED ? (j ^ (lane & 7)) : j) << 4)) = epi_convert_chunk<OUT>(v0, b0, epi);
                        tmem_ld_wait();
                        if (j + 2 < CHUNKS_PER_ROW) tmem_ld_32xN<CHUNK_COLS>(t_slab + (j + 2) * CHUNK_COLS, v0);
                        *reinterpret_cast<uint4 *>(row + ((SLAB_SWIZZLED ? ((j + 1) ^ (lane & 7)) : (j + 1)) << 4)) =
                            epi_convert_chunk<OUT>(v1, b1, epi);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0)
                    {
                        if (epi == EPI_RESIDUAL || epi == EPI_SPLITK)
                            tma_reduce_add_2d(&tma_out, slab_addr + buf_off, col0 + scol, row0);
                        else
                            tma_store_2d(&tma_out, slab_addr + buf_off, col0 + scol, row0);
                        tma_store_commit();
                    }
                }
            }
            else
            {
                // ---- direct epilogue: every lane writes its own accumulator row ----
                const int r = row0 + lane;
                long long orow = r;
                int prow = 0;
                if (epi == EPI_PATCH)
                {
                    const int b = r / p.remap_in, t = r - b * p.remap_in;
                    orow = (long long)b * p.remap_out + 1 + t;
                    prow = 1 + t;
                }
                const bool row_ok = r < p.M;
                for (int g = 0; g < WARP_COLS / 32; g++)
                {
                    const int gcol = col0 + wcol + g * 32; // first global column of the group
                    if (gcol >= p.N) break;                // warp-uniform
                    uint32_t v[32];
                    tmem_ld_32x32(t_base + g * 32, v);
                    tmem_ld_wait();
                    uint32_t w[32 * OELEM / 4];
                    const int act = (epi == EPI_RESIDUAL || epi == EPI_PATCH) ? EPI_NONE : epi;
                    epi_convert32<OUT>(v, bias_s + wcol + g * 32, act, w);
                    if (!row_ok) continue;
                    if constexpr (OUT == OUT_F32)
                    {
                        float *dst = reinterpret_cast<float *>(p.out) + orow * p.ldc + gcol;
                        const bool vec = gcol + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
                        if (vec)
                        {
                            float4 add[8];
                            if (epi == EPI_RESIDUAL || epi == EPI_PATCH)
                            {
                                const float4 *src = epi == EPI_RESIDUAL
                                                        ? reinterpret_cast<const float4 *>(dst)
                                                        : reinterpret_cast<const float4 *>(p.pos + (long long)prow * p.N + gcol);
#pragma unroll
                                for (int j = 0; j < 8; j++) add[j] = src[j]; // 8 independent loads in flight
                            }
                            else
                            {
#pragma unroll
                                for (int j = 0; j < 8; j++) add[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
#pragma unroll
                            for (int j = 0; j < 8; j++)
                                reinterpret_cast<float4 *>(dst)[j] =
                                    make_float4(__uint_as_float(w[4 * j]) + add[j].x, __uint_as_float(w[4 * j + 1]) + add[j].y,
                                                __uint_as_float(w[4 * j + 2]) + add[j].z, __uint_as_float(w[4 * j + 3]) + add[j].w);
                        }
                        else
                        {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (gcol + j < p.N)
                                {
                                    float val = __uint_as_float(w[j]);
                                    if (epi == EPI_RESIDUAL) val += dst[j];
                                    if (epi == EPI_PATCH) val += p.pos[(long long)prow * p.N + gcol + j];
                                    dst[j] = val;
                                }
                        }
                    }
                    else if constexpr (OUT == OUT_S32)
                    {
                        uint32_t *dst = reinterpret_cast<uint32_t *>(p.out) + orow * p.ldc + gcol;
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (gcol + j < p.N) dst[j] = w[j];
                    }
                    else if constexpr (OUT == OUT_BF16)
                    {
                        uint16_t *dst = reinterpret_cast<uint16_t *>(p.out) + orow * p.ldc + gcol;
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (gcol + j < p.N) dst[j] = (uint16_t)(w[j >> 1] >> (16 * (j & 1)));
                    }
                    else
                    {
                        uint8_t *dst = reinterpret_cast<uint8_t *>(p.out) + orow * p.ldc + gcol;
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (gcol + j < p.N) dst[j] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
                    }
                }
            }

            // all of this warp's TMEM reads of the tile are complete -> hand the accumulator back
            if (dbg) dslot[2] = clock64();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0)
            {
                if constexpr (CG == 2)
                    mbar_arrive_leader(tempty_bar(acc)); // the MMA issuer lives in the leader CTA
                else
                    mbar_arrive(tempty_bar(acc));
            }
            if (++acc == 2)
            {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        // The bulk stores must have COMPLETED (not only read their smem source) before this CTA exits: a dependent kernel
        // released by griddepcontrol.wait is only ordered after what the exited CTAs of this grid have made visible.
        if (lane == 0) tma_store_wait_all();
    }

    tcgen05_fence_before();
    if constexpr (CG == 2)
        cluster_sync_all(); // neither CTA may exit (or free TMEM) while the other can still touch its smem / barriers
    else
        __syncthreads();
    if (warp == W_ALLOC)
    {
        tcgen05_fence_after();
        if constexpr (CG == 2)
            tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else
            tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

} // namespace nc

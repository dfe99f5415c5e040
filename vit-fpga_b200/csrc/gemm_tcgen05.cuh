// gemm_tcgen05.cuh -- the dense hot kernel:  out[M][N] = epilogue(A[M][K] . W[N][K]^T + bias[N])
//
// This is the B200 replacement for the arithmetic of the reference's `network_v1` task
// (src/netFPGA.cpp:250,275: one fully-connected layer after another over W[out][in] matrices laid
// out as in src/netFPGA.cpp:91-106) and for every linear layer of the ViT.  Both operands are
// K-major, exactly the reference's row-major W[out][in] and sample-major activations, so no
// transposes are needed anywhere.
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0    TMA producer: 128B-swizzled [128 x 128B] A tile + [BN x 128B] W tile per stage
//   warp 1    MMA issuer:   one thread issues tcgen05.mma (UMMA 128 x BN x 32B) into TMEM
//   warp 2    TMEM allocator (2 accumulator stages x BN fp32/int32 columns)
//   warps 4-7 epilogue: tcgen05.ld -> +bias -> activation -> warp-private smem transpose ->
//             row-contiguous (coalesced) global stores; overlaps the next tile's MMAs
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue).
// All byte geometry is independent of the operand type: a stage row is always 128 bytes
// (64 bf16 / 32 tf32 / 128 int8), one UMMA consumes 32 bytes of K.
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
    EPI_PATCH = 6 // fp32 out with row remap + position embedding (patch embedding)
};

struct GemmParams
{
    int M, N, K;       // K in elements of the operand type
    const void *bias;  // float[N] (bf16 / tf32) or int32[N] (int8); may be null
    void *out;         // OUT_* typed, row pitch ldc elements
    long long ldc;
    int epi;
    // EPI_PATCH: input row r = b * remap_in + t  ->  output row b * remap_out + 1 + t, and
    // pos[(1 + t) * N + col] is added (cls token occupies output row b * remap_out).
    int remap_in, remap_out;
    const float *pos;
    int *error_flag;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_STAGE_ROW_BYTES = 128;

template <int BN, int STAGES>
struct GemmSmem
{
    static constexpr int A_BYTES = GEMM_BM * GEMM_STAGE_ROW_BYTES;
    static constexpr int B_BYTES = BN * GEMM_STAGE_ROW_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGING_WORDS_PER_WARP = 32 * 33;
    static constexpr int OFF_STAGING = STAGES * STAGE_BYTES;
    static constexpr int OFF_BIAS = OFF_STAGING + 4 * STAGING_WORDS_PER_WARP * 4;
    static constexpr int OFF_BARS = OFF_BIAS + 4 * BN * 4;
    static constexpr int NUM_BARS = 2 * STAGES + 4;
    static constexpr int OFF_TMEM_PTR = OFF_BARS + NUM_BARS * 8;
    static constexpr int TOTAL = OFF_TMEM_PTR + 16 + 1024; // +1024: manual alignment slack
};

template <int KIND>
struct KindTraits;
template <>
struct KindTraits<KIND_BF16>
{
    static constexpr int ELEM = 2;
    __host__ __device__ static constexpr uint32_t idesc(int bn) { return umma_idesc(1, 1, GEMM_BM, bn); }
};
template <>
struct KindTraits<KIND_TF32>
{
    static constexpr int ELEM = 4;
    __host__ __device__ static constexpr uint32_t idesc(int bn) { return umma_idesc(1, 2, GEMM_BM, bn); }
};
template <>
struct KindTraits<KIND_I8>
{
    static constexpr int ELEM = 1;
    __host__ __device__ static constexpr uint32_t idesc(int bn) { return umma_idesc(2, 1, GEMM_BM, bn); }
};

// ---- epilogue helpers --------------------------------------------------------------------------

__device__ __forceinline__ float epi_act_f32(float v, int epi)
{
    if (epi == EPI_RELU) return fmaxf(v, 0.0f);
    if (epi == EPI_GELU) return gelu_erf(v);
    return v;
}

template <int KIND, int BN, int OUT, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w,
                       const GemmParams p)
{
    using L = GemmSmem<BN, STAGES>;
    constexpr int ELEM = KindTraits<KIND>::ELEM;
    constexpr int BK = GEMM_STAGE_ROW_BYTES / ELEM; // elements of K per stage
    constexpr uint32_t TMEM_COLS = 2 * BN;          // two accumulator stages (power of two >= 32)
    constexpr uint32_t IDESC = KindTraits<KIND>::idesc(BN);
    static_assert(BN == 128 || BN == 256, "BN must be 128 or 256");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u; // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t *smem = smem_raw + (base - raw_addr);

    const uint32_t bars = base + L::OFF_BARS;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(smem + L::OFF_TMEM_PTR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM;
    const int tiles_n = (p.N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0)
    {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_w);
    }
    if (warp == 1 && lane == 0)
    {
        for (int s = 0; s < STAGES; s++)
        {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; a++)
        {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4); // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2)
    {
        tmem_alloc(base + L::OFF_TMEM_PTR, TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0)
    {
        // ===================== TMA producer =====================
        if (lane == 0)
        {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
            {
                const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
                for (int kb = 0; kb < num_kb; kb++)
                {
                    mbar_wait(empty_bar(stage), phase ^ 1u, p.error_flag, KERR_PRODUCER_EMPTY);
                    mbar_arrive_expect_tx(full_bar(stage), L::STAGE_BYTES);
                    const uint32_t a_dst = base + stage * L::STAGE_BYTES;
                    tma_load_2d(a_dst, &tma_a, full_bar(stage), kb * BK, m_blk * GEMM_BM);
                    tma_load_2d(a_dst + L::A_BYTES, &tma_w, full_bar(stage), kb * BK, n_blk * BN);
                    if (++stage == STAGES)
                    {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    }
    else if (warp == 1)
    {
        // ===================== MMA issuer =====================
        if (lane == 0)
        {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
            {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.error_flag, KERR_MMA_TMEM_EMPTY);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; kb++)
                {
                    mbar_wait(full_bar(stage), phase, p.error_flag, KERR_MMA_FULL);
                    tcgen05_fence_after();
                    const uint32_t a_addr = base + stage * L::STAGE_BYTES;
                    const uint64_t a_desc = umma_smem_desc_sw128(a_addr);
                    const uint64_t b_desc = umma_smem_desc_sw128(a_addr + L::A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; k++) // 4 x 32 bytes of K per stage; +2 = 32 B >> 4
                        umma_ss<KIND>(d_tmem, a_desc + 2u * k, b_desc + 2u * k, IDESC, (kb | k) != 0 ? 1u : 0u);
                    tcgen05_commit(empty_bar(stage)); // frees the smem slot once these MMAs retire
                    if (++stage == STAGES)
                    {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tcgen05_commit(tfull_bar(acc)); // accumulator complete -> epilogue
                if (++acc == 2)
                {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    }
    else if (warp >= 4)
    {
        // ===================== epilogue =====================
        const int q = warp & 3; // TMEM lane quarter this warp may read: lanes [32q, 32q+32)
        uint32_t *stg = reinterpret_cast<uint32_t *>(smem + L::OFF_STAGING) + q * L::STAGING_WORDS_PER_WARP;
        uint32_t *bias_s = reinterpret_cast<uint32_t *>(smem + L::OFF_BIAS) + q * BN;
        uint32_t acc = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        {
            const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
            const int col0 = n_blk * BN;
            const int row0 = m_blk * GEMM_BM + q * 32;

            // warp-private copy of this tile's bias slice (bit pattern: float or int32)
            for (int i = lane; i < BN; i += 32)
            {
                const int c = col0 + i;
                bias_s[i] = (p.bias != nullptr && c < p.N) ? reinterpret_cast<const uint32_t *>(p.bias)[c] : 0u;
            }
            // per-lane output row (identity, or patch-embedding remap) and position-embedding row
            long long my_orow = row0 + lane;
            int my_prow = 0;
            if (p.epi == EPI_PATCH)
            {
                const int r = row0 + lane;
                const int b = r / p.remap_in, t = r - b * p.remap_in;
                my_orow = (long long)b * p.remap_out + 1 + t;
                my_prow = 1 + t;
            }
            __syncwarp();

            mbar_wait(tfull_bar(acc), acc_phase, p.error_flag, KERR_EPI_TMEM_FULL);
            tcgen05_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;

            if constexpr (OUT == OUT_BF16)
            {
                // 64 columns per pass: 32 packed bf16x2 words per row -> one 128-byte row segment per store
                __nv_bfloat16 *out = reinterpret_cast<__nv_bfloat16 *>(p.out);
                for (int c = 0; c < BN / 64; c++)
                {
                    if (col0 + c * 64 >= p.N) break;
                    uint32_t v[64];
                    tmem_ld_32x32(t_base + c * 64, v);
                    tmem_ld_32x32(t_base + c * 64 + 32, v + 32);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j++)
                    {
                        const float lo = epi_act_f32(__uint_as_float(v[2 * j]) + __uint_as_float(bias_s[c * 64 + 2 * j]), p.epi);
                        const float hi = epi_act_f32(__uint_as_float(v[2 * j + 1]) + __uint_as_float(bias_s[c * 64 + 2 * j + 1]), p.epi);
                        stg[lane * 33 + j] = pack_bf16x2(lo, hi);
                    }
                    __syncwarp();
                    const int gcol = col0 + c * 64 + 2 * lane;
#pragma unroll 8
                    for (int r = 0; r < 32; r++)
                    {
                        const uint32_t w = stg[r * 33 + lane];
                        const long long grow = row0 + r;
                        if (grow < p.M)
                        {
                            __nv_bfloat16 *dst = out + grow * p.ldc + gcol;
                            if (gcol + 1 < p.N)
                                *reinterpret_cast<uint32_t *>(dst) = w;
                            else if (gcol < p.N)
                                *reinterpret_cast<uint16_t *>(dst) = (uint16_t)(w & 0xFFFFu);
                        }
                    }
                    __syncwarp();
                }
            }
            else if constexpr (OUT == OUT_F32 || OUT == OUT_S32)
            {
                uint32_t *out = reinterpret_cast<uint32_t *>(p.out);
                for (int c = 0; c < BN / 32; c++)
                {
                    if (col0 + c * 32 >= p.N) break;
                    uint32_t v[32];
                    tmem_ld_32x32(t_base + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j++)
                    {
                        uint32_t w;
                        if constexpr (OUT == OUT_F32)
                            w = __float_as_uint(epi_act_f32(__uint_as_float(v[j]) + __uint_as_float(bias_s[c * 32 + j]), p.epi));
                        else
                        {
                            int a = (int)v[j] + (int)bias_s[c * 32 + j];
                            if (p.epi == EPI_RELU) a = max(a, 0);
                            w = (uint32_t)a;
                        }
                        stg[lane * 33 + j] = w;
                    }
                    __syncwarp();
                    const int gcol = col0 + c * 32 + lane;
#pragma unroll 8
                    for (int r = 0; r < 32; r++)
                    {
                        uint32_t w = stg[r * 33 + lane];
                        const long long orow = __shfl_sync(0xffffffffu, my_orow, r);
                        const int prow = __shfl_sync(0xffffffffu, my_prow, r);
                        if (row0 + r < p.M && gcol < p.N)
                        {
                            uint32_t *dst = out + orow * p.ldc + gcol;
                            if constexpr (OUT == OUT_F32)
                            {
                                if (p.epi == EPI_RESIDUAL)
                                    w = __float_as_uint(__uint_as_float(*dst) + __uint_as_float(w));
                                else if (p.epi == EPI_PATCH)
                                    w = __float_as_uint(__uint_as_float(w) + p.pos[(long long)prow * p.N + gcol]);
                            }
                            *dst = w;
                        }
                    }
                    __syncwarp();
                }
            }
            else
            {
                // OUT_S8 requantisation: q = clamp((relu?)(acc + bias) >> 7, -128, 127); 8 packed words per row
                int8_t *out = reinterpret_cast<int8_t *>(p.out);
                for (int c = 0; c < BN / 32; c++)
                {
                    if (col0 + c * 32 >= p.N) break;
                    uint32_t v[32];
                    tmem_ld_32x32(t_base + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; j++)
                    {
                        uint32_t w = 0;
#pragma unroll
                        for (int e = 0; e < 4; e++)
                        {
                            int a = (int)v[4 * j + e] + (int)bias_s[c * 32 + 4 * j + e];
                            if (p.epi == EPI_REQUANT_RELU) a = max(a, 0);
                            a = min(127, max(-128, a >> 7));
                            w |= ((uint32_t)a & 0xFFu) << (8 * e);
                        }
                        stg[lane * 9 + j] = w;
                    }
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < 8; it++)
                    {
                        const int r = it * 4 + (lane >> 3), j = lane & 7;
                        const uint32_t w = stg[r * 9 + j];
                        const long long grow = row0 + r;
                        const int gcol = col0 + c * 32 + 4 * j;
                        if (grow < p.M)
                        {
                            int8_t *dst = out + grow * p.ldc + gcol;
                            if (gcol + 3 < p.N)
                                *reinterpret_cast<uint32_t *>(dst) = w;
                            else
                                for (int e = 0; e < 4; e++)
                                    if (gcol + e < p.N) dst[e] = (int8_t)((w >> (8 * e)) & 0xFFu);
                        }
                    }
                    __syncwarp();
                }
            }

            // all of this warp's TMEM reads of the tile are complete -> hand the accumulator back
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2)
            {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2)
    {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

} // namespace nc

// gemm.cu -- host launcher + tensor-map construction for the tcgen05 dense kernel, plus the two
// CUDA-core dense kernels:
//   * gemm_fp32_ordered_kernel: NETCUDA_PREC_FP32.  acc = bias, then fmaf in ascending k -- the exact
//     operation sequence of oracle_mlp_forward_one (oracle/oracle_mlp.c), hence bit-equal results.
//   * gemm_ref_kernel: same operand types and epilogues as the tcgen05 kernel, one thread per output.
//     Selected with variant = 1; exists so the tensor-core path can be cross-checked on the GPU at
//     full problem sizes.  It is a CUDA kernel, not a CPU fallback.
#include "gemm_tcgen05.cuh"
#include "kernels.h"

#include <cstdlib>
#include <mutex>

namespace nc
{

// ---- driver entry point for cuTensorMapEncodeTiled (no link-time dependency on libcuda) --------

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;

cudaError_t encode_tma_2d(void *map_out, int dtype_bytes, const void *ptr, long long dim0, long long dim1, long long pitch_bytes,
                          int box0, int box1, bool swizzle128)
{
    if (!g_encode_tiled) return cudaErrorNotReady;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (pitch_bytes & 15) != 0 || dim0 <= 0 || dim1 <= 0) return cudaErrorInvalidValue;
    const CUtensorMapDataType dt = dtype_bytes == 2   ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                   : dtype_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                                      : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    cuuint64_t gdim[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
    cuuint64_t gstride[1] = {(cuuint64_t)pitch_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = g_encode_tiled(reinterpret_cast<CUtensorMap *>(map_out), dt, 2, const_cast<void *>(ptr), gdim, gstride, box, estride,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// 3-D map over a [dim2][dim1][dim0] array (dim0 contiguous), box = box0 x box1 x 1: a TMA store clips the rows of a box that lie
// beyond dim1, so tiles that straddle the end of one image never touch the next one.
cudaError_t encode_tma_3d(void *map_out, int dtype_bytes, const void *ptr, long long dim0, long long dim1, long long dim2, long long stride1_bytes,
                          long long stride2_bytes, int box0, int box1, bool swizzle128)
{
    if (!g_encode_tiled) return cudaErrorNotReady;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (stride1_bytes & 15) != 0 || (stride2_bytes & 15) != 0 || dim0 <= 0 || dim1 <= 0 || dim2 <= 0)
        return cudaErrorInvalidValue;
    const CUtensorMapDataType dt = dtype_bytes == 2   ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                   : dtype_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                                      : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    cuuint64_t gdim[3] = {(cuuint64_t)dim0, (cuuint64_t)dim1, (cuuint64_t)dim2};
    cuuint64_t gstride[2] = {(cuuint64_t)stride1_bytes, (cuuint64_t)stride2_bytes};
    cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
    cuuint32_t estride[3] = {1, 1, 1};
    CUresult r = g_encode_tiled(reinterpret_cast<CUtensorMap *>(map_out), dt, 3, const_cast<void *>(ptr), gdim, gstride, box, estride,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int KIND, int BN, int OUT, int STAGES, int CG, int EW, int DS = 0>
static cudaError_t opt_in_smem()
{
    return cudaFuncSetAttribute(gemm_tn_tcgen05_kernel<KIND, BN, OUT, STAGES, CG, EW, DS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GemmSmem<BN, STAGES, CG, EW, DS>::TOTAL);
}

// smem ring depth: what fits next to the epilogue slabs (EW x 4 KB) in 227 KB
constexpr int STAGES_256 = 4;      // 1 CTA per tile, BN = 256: 48 KB per stage
constexpr int STAGES_128 = 6;      // 1 CTA per tile, BN = 128: 32 KB per stage
constexpr int STAGES_256_PAIR = 6; // CTA pair, BN = 256: 16 KB of A + 16 KB of W per CTA and stage
constexpr int STAGES_256_PAIR_EW16 = 5; // the same with 16 epilogue warps (64 KB of slabs)
constexpr int STAGES_256_PAIR_DS = 5;   // the same with two output slabs per epilogue warp (64 KB of slabs)

// the latency-oriented ordered fp32 kernel (defined below, next to the 64 x 64 tile kernel)
constexpr int OS_KC = 64, OS_PITCH = OS_KC + 4, OS_STAGES = 8, OS_ROWS = 64, OS_THREADS = 256;
__host__ __device__ constexpr int os_smem_bytes(int ch) { return OS_STAGES * (OS_ROWS + 4 * ch) * OS_PITCH * 4; }
template <int CH>
__global__ void gemm_fp32_ordered_small_kernel(const float *__restrict__ a, long long lda, const float *__restrict__ w, long long ldw,
                                               const float *__restrict__ bias, float *__restrict__ out, long long ldc, int relu, int M, int N,
                                               int K, int vec_ok);

// cudaFuncSetAttribute is per device, so the opt-in runs once for every device that is used.
cudaError_t gemm_global_init()
{
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    if (!g_encode_tiled)
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    }
    if (dev < 64 && done[dev]) return cudaSuccess;
#define NC_OPT(K, O)                                                                    \
    if ((e = opt_in_smem<K, 256, O, STAGES_256_PAIR, 2, 8>()) != cudaSuccess) return e; \
    if ((e = opt_in_smem<K, 256, O, STAGES_256, 1, 8>()) != cudaSuccess) return e;      \
    if ((e = opt_in_smem<K, 128, O, STAGES_128, 1, 8>()) != cudaSuccess) return e;
    NC_OPT(KIND_BF16, OUT_BF16)
    NC_OPT(KIND_BF16, OUT_F32)
    NC_OPT(KIND_TF32, OUT_F32)
    NC_OPT(KIND_TF32, OUT_BF16) // tf32 ViTs: the qkv projection feeds the bf16 attention core
    NC_OPT(KIND_I8, OUT_S8)
    NC_OPT(KIND_I8, OUT_S32)
#undef NC_OPT
    if ((e = opt_in_smem<KIND_BF16, 256, OUT_BF16, STAGES_256_PAIR_EW16, 2, 16>()) != cudaSuccess) return e;
    if ((e = opt_in_smem<KIND_BF16, 256, OUT_F32, STAGES_256_PAIR_EW16, 2, 16>()) != cudaSuccess) return e;
    if ((e = opt_in_smem<KIND_TF32, 256, OUT_F32, STAGES_256_PAIR_EW16, 2, 16>()) != cudaSuccess) return e;
    if ((e = opt_in_smem<KIND_TF32, 256, OUT_BF16, STAGES_256_PAIR_EW16, 2, 16>()) != cudaSuccess) return e;
    if ((e = opt_in_smem<KIND_BF16, 256, OUT_BF16, STAGES_256_PAIR_DS, 2, 8, 1>()) != cudaSuccess) return e;
    if ((e = opt_in_smem<KIND_BF16, 256, OUT_F32, STAGES_256_PAIR_DS, 2, 8, 1>()) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(gemm_fp32_ordered_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, os_smem_bytes(1))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(gemm_fp32_ordered_small_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, os_smem_bytes(2))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(gemm_fp32_ordered_small_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, os_smem_bytes(4))) != cudaSuccess) return e;
    if (dev < 64) done[dev] = true;
    return cudaSuccess;
}

static int elem_size(int kind) { return kind == GK_BF16 ? 2 : kind == GK_I8 ? 1 : 4; }

// 2-D K-major operand: dims {K, rows}, row pitch `pitch_bytes`, box {128 B of K, box_rows}, 128B swizzle.
static cudaError_t make_operand_map(CUtensorMap *map, int kind, const void *ptr, long long k, long long rows,
                                    long long pitch_bytes, int box_rows)
{
    if (!g_encode_tiled) return cudaErrorNotReady;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (pitch_bytes & 15) != 0 || k <= 0 || rows <= 0)
        return cudaErrorInvalidValue;
    const CUtensorMapDataType dt = kind == GK_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                   : kind == GK_I8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                                   : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)pitch_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(GEMM_STAGE_ROW_BYTES / elem_size(kind)), (cuuint32_t)box_rows};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = g_encode_tiled(map, dt, 2, const_cast<void *>(ptr), gdim, gstride, box, estride,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// Output matrix as a 2-D tensor {N, M} for the epilogue's TMA stores: box = one epilogue slab
// (slab_cols columns x 32 rows), 128B-swizzled when a slab row is 128 bytes.
template <int BN, int OUT, int EW>
static cudaError_t make_out_map(CUtensorMap *map, void *ptr, long long n, long long m, long long pitch_bytes)
{
    if (!g_encode_tiled) return cudaErrorNotReady;
    constexpr int cols = slab_cols<BN, OUT, EW>();
    constexpr int row_bytes = cols * OutTraits<OUT>::ELEM;
    const CUtensorMapDataType dt = OUT == OUT_BF16  ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                   : OUT == OUT_S8  ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                   : OUT == OUT_S32 ? CU_TENSOR_MAP_DATA_TYPE_INT32
                                                    : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)m};
    cuuint64_t gstride[1] = {(cuuint64_t)pitch_bytes};
    cuuint32_t box[2] = {(cuuint32_t)cols, 32};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = g_encode_tiled(map, dt, 2, ptr, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int KIND, int BN, int OUT, int STAGES, int CG, int EW = 8, int DS = 0>
static cudaError_t launch_tc(const GemmCall &c, cudaStream_t stream)
{
    CUtensorMap map_a, map_w, map_out;
    cudaError_t e = make_operand_map(&map_a, c.kind, c.a, c.k, c.a_rows > c.m ? c.a_rows : c.m, c.lda * elem_size(c.kind), GEMM_BM);
    if (e != cudaSuccess) return e;
    e = make_operand_map(&map_w, c.kind, c.w, c.k, c.n, c.ldw * elem_size(c.kind), BN / CG);
    if (e != cudaSuccess) return e;
    GemmParams p;
    p.M = c.m, p.N = c.n, p.K = c.k;
    p.bias = c.bias, p.out = c.out, p.ldc = c.ldc, p.epi = c.epi;
    static const bool gelu_scalar = getenv("NETCUDA_GELU_SCALAR") != nullptr; // A/B switch: one element per FFMA chain
    if (OUT == OUT_BF16 && c.epi == EPI_GELU && !gelu_scalar) p.epi = EPI_GELU_X2;
    p.remap_in = c.remap_in, p.remap_out = c.remap_out, p.pos = c.pos;
    p.error_flag = c.error_flag;
    p.k_splits = c.k_splits > 1 ? c.k_splits : 1;
    p.debug = nullptr;
#ifdef NETCUDA_DEBUG_TIMELINE // clock-stamp hooks of tools/gemm_timeline.py: compiled out of release builds
    if (const char *dbg = getenv("NETCUDA_GEMM_DEBUG_PTR")) p.debug = reinterpret_cast<long long *>(strtoull(dbg, nullptr, 0));
    if (const char *de = getenv("NETCUDA_GEMM_DEBUG_EPI")) // timeline of the launches with one epilogue only (the last such launch wins)
        if (atoi(de) != c.epi) p.debug = nullptr;
#endif
    // TMA-store epilogue whenever the output is addressable by a tensor map; otherwise direct stores
    const long long pitch_bytes = c.ldc * OutTraits<OUT>::ELEM;
    // (TMA bounds the innermost dimension in 16-byte units, so N must be a whole number of them as well)
    p.tma_store = (c.epi != EPI_PATCH && (reinterpret_cast<uintptr_t>(c.out) & 15u) == 0 && (pitch_bytes & 15) == 0 &&
                   ((c.n * (long long)OutTraits<OUT>::ELEM) & 15) == 0)
                      ? 1
                      : 0;
    if (p.tma_store)
    {
        e = make_out_map<BN, OUT, EW>(&map_out, c.out, c.n, c.m, pitch_bytes);
        if (e != cudaSuccess) return e;
    }
    else
        map_out = map_a; // never dereferenced
    if (p.k_splits > 1 && !(p.tma_store && OUT == OUT_S32 && c.epi == EPI_SPLITK)) return cudaErrorInvalidValue;
    const int tiles = ((c.m + GEMM_BM * CG - 1) / (GEMM_BM * CG)) * ((c.n + BN - 1) / BN) * p.k_splits;
    const int sms = c.num_sms > 0 ? c.num_sms : 148;
    const int slots = sms / CG; // tiles in flight: one per CTA, or one per CTA pair
    return launch_pdl(gemm_tn_tcgen05_kernel<KIND, BN, OUT, STAGES, CG, EW, DS>, dim3((unsigned)(CG * (tiles < slots ? tiles : slots))),
                      dim3(gemm_threads(EW)), (size_t)GemmSmem<BN, STAGES, CG, EW, DS>::TOTAL, stream, CG, map_a, map_w, map_out, p);
}

template <int KIND, int OUT>
static cudaError_t launch_tc_bn(const GemmCall &c, cudaStream_t stream)
{
    if (c.n <= 128) return launch_tc<KIND, 128, OUT, STAGES_128, 1>(c, stream);
    // A few hundred rows (a single image of a ViT: M = 197) make a handful of 256 x 256 pair tiles, each a long serial K loop on two
    // SMs while the rest of the GPU idles (ViT-B fc2, one image: 3 tiles, 23 us).  128 x 128 single-CTA tiles spread the same work
    // over four times as many SMs; same k order per output element, same bits.
    {
        const int sms = c.num_sms > 0 ? c.num_sms : 148;
        const long long pair_tiles = (long long)((c.m + 255) / 256) * ((c.n + 255) / 256);
        if (c.variant == 0 && c.k_splits <= 1 && pair_tiles * 4 <= sms) return launch_tc<KIND, 128, OUT, STAGES_128, 1>(c, stream);
    }
    // more than one 128-row block: pair the SMs (256-row tiles, half the W traffic per SM); variant 2 forces single CTAs
    if (c.m > GEMM_BM && c.variant != 2)
    {
        // variant 3: 16 epilogue warps for the GELU epilogue.  Measured on ViT-B fc1 (ncu, round 1): 212.5 us vs 203.5 us with
        // 8 warps -- the epilogue is bound by MUFU/FMA work per element, not by the number of warps -- so it is not the default.
        // ... but it is the default where K is so short that the GELU epilogue is all there is (ViT-Tiny fc1, K = 192: 40.7 -> 34.7 us).
        if constexpr ((KIND == KIND_BF16 || KIND == KIND_TF32) && (OUT == OUT_BF16 || OUT == OUT_F32))
        {
            // ... and of every other GEMM that is all epilogue, K of at most four k-blocks (ViT-Tiny, 256 images, inside the step: qkv
            // 0.323 -> 0.300 ms, proj 0.322 -> 0.307 ms, step 139.8 k -> 141.8 k images/s; alone, tools/gemm_k192_ab.py: 23.8 -> 22.8 us and
            // 18.7 -> 16.6 us.  Not for a single column of tiles with a long K -- fc2, N = 192: faster alone, 28.8 -> 27.1 us, no
            // faster in the step.  tf32 operands alike: TF32 ViT-Tiny step 92.6 k -> 98.7 k images/s, fc1 (fp32-output GELU) 0.88 -> 0.62 ms.
            // NETCUDA_GEMM_EW16=0: A/B)
            static const bool ew16_short = !getenv("NETCUDA_GEMM_EW16") || atoi(getenv("NETCUDA_GEMM_EW16")) != 0;
            if (c.variant == 3 || (c.variant == 0 && ((OUT == OUT_BF16 && c.epi == EPI_GELU && c.k <= 256) || (ew16_short && c.k <= 256))))
                return launch_tc<KIND, 256, OUT, STAGES_256_PAIR_EW16, 2, 16>(c, stream);
        }
        // Two output slabs per epilogue warp (slab i + 1 is filled while the TMA store of slab i drains) at the price of one
        // pipeline stage: pays where the epilogue or the store path sets the pace -- the GELU epilogue (ViT-B fc1: 10.3 -> 9.9 ms
        // per step) and the short-K residual update that is bound by the L2 reduce-add (proj: 3.95 -> 3.6 ms) -- and costs where the
        // mainloop does (qkv 6.75 -> 7.0 ms, fc2 8.4 -> 8.7 ms).  Variant 4 forces it everywhere, variant 5 nowhere (A/B).
        if constexpr (KIND == KIND_BF16 && (OUT == OUT_BF16 || OUT == OUT_F32))
        {
            // (... and any GEMM whose K is so short that the epilogue is all there is: ViT-Tiny qkv, K = 192: 29.3 -> 27.2 us)
            const bool pays = c.epi == EPI_GELU || (c.epi == EPI_RESIDUAL && c.k <= 1024) || c.k <= 256;
            if (c.variant == 4 || (c.variant == 0 && pays)) return launch_tc<KIND, 256, OUT, STAGES_256_PAIR_DS, 2, 8, 1>(c, stream);
        }
        return launch_tc<KIND, 256, OUT, STAGES_256_PAIR, 2>(c, stream);
    }
    return launch_tc<KIND, 256, OUT, STAGES_256, 1>(c, stream);
}

// ---- CUDA-core reference with identical operand types / epilogues ----------------------------------

template <int KIND>
__global__ void gemm_ref_kernel(const void *__restrict__ a_, long long lda, const void *__restrict__ w_, long long ldw,
                                const void *__restrict__ bias_, void *out_, long long ldc, int out_type, int epi, int M, int N,
                                int K, int remap_in, int remap_out, const float *__restrict__ pos)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y * blockDim.y + threadIdx.y;
    if (row >= M || col >= N) return;
    long long orow = row;
    int prow = 0;
    if (epi == EPI_PATCH)
    {
        const int b = row / remap_in, t = row - b * remap_in;
        orow = (long long)b * remap_out + 1 + t;
        prow = 1 + t;
    }
    if constexpr (KIND == KIND_I8)
    {
        const int8_t *a = reinterpret_cast<const int8_t *>(a_) + (long long)row * lda;
        const int8_t *w = reinterpret_cast<const int8_t *>(w_) + (long long)col * ldw;
        int acc = bias_ ? reinterpret_cast<const int *>(bias_)[col] : 0;
        for (int k = 0; k < K; k++) acc += (int)a[k] * (int)w[k];
        if (out_type == OUT_S32)
        {
            if (epi == EPI_RELU) acc = max(acc, 0);
            reinterpret_cast<int *>(out_)[orow * ldc + col] = acc;
        }
        else
        {
            if (epi == EPI_REQUANT_RELU) acc = max(acc, 0);
            reinterpret_cast<int8_t *>(out_)[orow * ldc + col] = (int8_t)min(127, max(-128, acc >> 7));
        }
    }
    else
    {
        float acc = 0.0f;
        if constexpr (KIND == KIND_BF16)
        {
            const __nv_bfloat16 *a = reinterpret_cast<const __nv_bfloat16 *>(a_) + (long long)row * lda;
            const __nv_bfloat16 *w = reinterpret_cast<const __nv_bfloat16 *>(w_) + (long long)col * ldw;
            for (int k = 0; k < K; k++) acc = fmaf(__bfloat162float(a[k]), __bfloat162float(w[k]), acc);
        }
        else
        {
            const float *a = reinterpret_cast<const float *>(a_) + (long long)row * lda;
            const float *w = reinterpret_cast<const float *>(w_) + (long long)col * ldw;
            for (int k = 0; k < K; k++) // tf32 operands: the tensor core ignores the low 13 mantissa bits
                acc = fmaf(__uint_as_float(__float_as_uint(a[k]) & 0xFFFFE000u), __uint_as_float(__float_as_uint(w[k]) & 0xFFFFE000u), acc);
        }
        float v = acc + (bias_ ? reinterpret_cast<const float *>(bias_)[col] : 0.0f);
        v = epi_act_f32(v, epi);
        if (out_type == OUT_BF16)
            reinterpret_cast<__nv_bfloat16 *>(out_)[orow * ldc + col] = __float2bfloat16_rn(v);
        else
        {
            float *dst = reinterpret_cast<float *>(out_) + orow * ldc + col;
            if (epi == EPI_RESIDUAL)
                v += *dst;
            else if (epi == EPI_PATCH)
                v += pos[(long long)prow * N + col];
            *dst = v;
        }
    }
}

template <int KIND>
static cudaError_t launch_ref(const GemmCall &c, cudaStream_t stream)
{
    dim3 block(32, 8), grid((c.n + 31) / 32, (c.m + 7) / 8);
    gemm_ref_kernel<KIND><<<grid, block, 0, stream>>>(c.a, c.lda, c.w, c.ldw, c.bias, c.out, c.ldc, c.out_type, c.epi, c.m, c.n,
                                                      c.k, c.remap_in, c.remap_out, c.pos);
    return cudaGetLastError();
}

// ---- NETCUDA_PREC_FP32: ordered fp32 on CUDA cores ------------------------------------------------

__global__ void __launch_bounds__(256)
gemm_fp32_ordered_kernel(const float *__restrict__ a, long long lda, const float *__restrict__ w, long long ldw,
                         const float *__restrict__ bias, float *__restrict__ out, long long ldc, int relu, int M, int N, int K)
{
    __shared__ float As[16][65];
    __shared__ float Ws[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int row0 = blockIdx.y * 64, col0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            const int c = col0 + tx * 4 + j;
            acc[i][j] = (bias != nullptr && c < N) ? bias[c] : 0.0f;
        }
    for (int k0 = 0; k0 < K; k0 += 16)
    {
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            const int e = threadIdx.x + i * 256;
            const int r = e >> 4, kk = e & 15;
            const bool kin = (k0 + kk) < K;
            As[kk][r] = (kin && row0 + r < M) ? a[(long long)(row0 + r) * lda + k0 + kk] : 0.0f;
            Ws[kk][r] = (kin && col0 + r < N) ? w[(long long)(col0 + r) * ldw + k0 + kk] : 0.0f;
        }
        __syncthreads();
        const int kmax = (K - k0) < 16 ? (K - k0) : 16;
        for (int kk = 0; kk < kmax; kk++) // strictly ascending k: same rounding sequence as the oracle
        {
            float av[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) wv[j] = Ws[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(wv[j], av[i], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const int r = row0 + ty * 4 + i;
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            const int c = col0 + tx * 4 + j;
            if (c >= N) continue;
            float v = acc[i][j];
            if (relu && v < 0.0f) v = 0.0f;
            out[(long long)r * ldc + c] = v;
        }
    }
}

// The same arithmetic for problems that do not fill the GPU with 64 x 64 tiles (the reference's own contract: ONE sample per
// launch_forward call, src/netFPGA.cpp:266-289; BASELINE config 1: 784-128-64-10 at 1..64 samples).  There every output is a chain
// of K dependent FMAs (4 cycles each: 784 + 128 + 64 of them are ~2 us for config 1) and everything else is latency to be hidden:
//   * a CTA computes 64 samples x 4 CH neurons, one thread = one sample x CH neurons, so a layer of 128 neurons is 32 CTAs
//     (the 64 x 64 kernel above: 2 CTAs, 49 unpipelined load -> barrier -> 16 FMAs -> barrier rounds, ~60 us);
//   * operands arrive through an 8-deep cp.async ring of 64-wide k chunks (17 KB each; 4-deep measured the same on config C1:
//     the ring is not what bounds it), rows kept k-contiguous at a pitch of 68 words: a
//     thread reads ITS sample's row as float4 (pitch = 4 mod 32 words: the eight lanes of a quarter warp hit 32 distinct banks) and
//     the neuron rows as float4 broadcasts -- (1 + CH) LDS.128 per 4 CH FMAs, no transposition on the way in;
//   * programmatic dependent launch: the next layer's CTAs are resident (and past their set-up) when this layer's last store lands.
// Same operation sequence per output as the oracle (acc = bias, fmaf in ascending k): bit-equal, tests/test_gpu_nets.py.

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src, int src_bytes) // bytes past src_bytes are zero-filled
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int CH>
__global__ void __launch_bounds__(OS_THREADS)
gemm_fp32_ordered_small_kernel(const float *__restrict__ a, long long lda, const float *__restrict__ w, long long ldw,
                               const float *__restrict__ bias, float *__restrict__ out, long long ldc, int relu, int M, int N, int K,
                               int vec_ok)
{
    extern __shared__ __align__(16) float os_smem[];
    constexpr int NT = 4 * CH;                                // neurons per CTA
    constexpr int STAGE_WORDS = (OS_ROWS + NT) * OS_PITCH;    // 64 sample rows, then NT neuron rows
    const int tid = threadIdx.x;
    const int row0 = blockIdx.y * OS_ROWS, col0 = blockIdx.x * NT;
    const int mrows = min(OS_ROWS, M - row0), nrows = min(NT, N - col0);
    const int nchunks = (K + OS_KC - 1) / OS_KC;
    const uint32_t smem_base = smem_u32(os_smem);

    griddep_launch_dependents(); // the next layer may take its seats; it reads nothing before its own griddep_wait()
    griddep_wait();              // the previous layer's activations are complete and visible from here on

    auto load_stage = [&](int kc)
    {
        const uint32_t st = smem_base + (uint32_t)((kc % OS_STAGES) * STAGE_WORDS * 4);
        const int k0 = kc * OS_KC, kn = min(OS_KC, K - k0);
        if (vec_ok) // 16-byte pieces: both base pointers and both row pitches are multiples of 16 bytes
        {
            const int total = (mrows + nrows) * (OS_KC / 4);
            for (int p = tid; p < total; p += OS_THREADS)
            {
                const int r = p >> 4, j = p & 15;
                if (4 * j >= kn) continue;
                const float *src = r < mrows ? a + (long long)(row0 + r) * lda : w + (long long)(col0 + r - mrows) * ldw;
                const int srow = r < mrows ? r : OS_ROWS + (r - mrows);
                cp_async_16(st + (uint32_t)((srow * OS_PITCH + 4 * j) * 4), src + k0 + 4 * j, min(16, 4 * (kn - 4 * j)));
            }
        }
        else
        {
            const int total = (mrows + nrows) * OS_KC;
            for (int p = tid; p < total; p += OS_THREADS)
            {
                const int r = p >> 6, kk = p & 63;
                if (kk >= kn) continue;
                const float *src = r < mrows ? a + (long long)(row0 + r) * lda : w + (long long)(col0 + r - mrows) * ldw;
                const int srow = r < mrows ? r : OS_ROWS + (r - mrows);
                cp_async_4(st + (uint32_t)((srow * OS_PITCH + kk) * 4), src + k0 + kk);
            }
        }
    };

    const int m = tid & (OS_ROWS - 1), ng = tid >> 6; // sample of the tile; group of CH neurons (uniform over a warp)
    const bool active = m < mrows && ng * CH < nrows;
    float acc[CH];
#pragma unroll
    for (int c = 0; c < CH; c++)
    {
        const int n = col0 + ng * CH + c;
        acc[c] = (bias != nullptr && n < N) ? bias[n] : 0.0f;
    }

    for (int s = 0; s < OS_STAGES - 1; s++)
    {
        if (s < nchunks) load_stage(s);
        cp_async_commit();
    }
    for (int kc = 0; kc < nchunks; kc++)
    {
        cp_async_wait<OS_STAGES - 2>(); // chunk kc has landed (this thread's pieces) ...
        __syncthreads();                // ... everybody's; and everybody is done reading the slot refilled below (chunk kc - 1)
        if (kc + OS_STAGES - 1 < nchunks) load_stage(kc + OS_STAGES - 1);
        cp_async_commit();
        if (!active) continue;
        const float *st = os_smem + (kc % OS_STAGES) * STAGE_WORDS;
        const float *as = st + m * OS_PITCH;
        const float *ws = st + (OS_ROWS + ng * CH) * OS_PITCH;
        const int kn = min(OS_KC, K - kc * OS_KC);
        int k = 0;
#pragma unroll 8
        for (; k + 4 <= kn; k += 4) // strictly ascending k: the oracle's rounding sequence
        {
            const float4 av = *reinterpret_cast<const float4 *>(as + k);
            float4 wv[CH];
#pragma unroll
            for (int c = 0; c < CH; c++) wv[c] = *reinterpret_cast<const float4 *>(ws + c * OS_PITCH + k);
#pragma unroll
            for (int c = 0; c < CH; c++) acc[c] = fmaf(wv[c].x, av.x, acc[c]);
#pragma unroll
            for (int c = 0; c < CH; c++) acc[c] = fmaf(wv[c].y, av.y, acc[c]);
#pragma unroll
            for (int c = 0; c < CH; c++) acc[c] = fmaf(wv[c].z, av.z, acc[c]);
#pragma unroll
            for (int c = 0; c < CH; c++) acc[c] = fmaf(wv[c].w, av.w, acc[c]);
        }
        for (; k < kn; k++)
        {
            const float a1 = as[k];
#pragma unroll
            for (int c = 0; c < CH; c++) acc[c] = fmaf(ws[c * OS_PITCH + k], a1, acc[c]);
        }
    }
    cp_async_wait<0>();
    if (!active) return;
#pragma unroll
    for (int c = 0; c < CH; c++)
    {
        const int n = col0 + ng * CH + c;
        if (n >= N) continue;
        float v = acc[c];
        if (relu && v < 0.0f) v = 0.0f;
        out[(long long)(row0 + m) * ldc + n] = v;
    }
}

template <int CH>
static cudaError_t launch_fp32_small(const GemmCall &c, cudaStream_t stream)
{
    const int vec_ok = ((reinterpret_cast<uintptr_t>(c.a) | reinterpret_cast<uintptr_t>(c.w)) & 15u) == 0 && (c.lda & 3) == 0 && (c.ldw & 3) == 0;
    const dim3 grid((unsigned)((c.n + 4 * CH - 1) / (4 * CH)), (unsigned)((c.m + OS_ROWS - 1) / OS_ROWS));
    return launch_pdl(gemm_fp32_ordered_small_kernel<CH>, grid, dim3(OS_THREADS), (size_t)os_smem_bytes(CH), stream, 1, (const float *)c.a, c.lda,
                      (const float *)c.w, c.ldw, (const float *)c.bias, (float *)c.out, c.ldc, c.epi == EPI_RELU ? 1 : 0, c.m, c.n, c.k, vec_ok);
}

// ---- dispatcher -------------------------------------------------------------------------------------

cudaError_t launch_gemm(const GemmCall &c, cudaStream_t stream)
{
    if (c.m <= 0 || c.n <= 0 || c.k <= 0) return cudaErrorInvalidValue;
    if (c.kind == GK_FP32_SIMT)
    {
        if (c.out_type != OUT_F32 || (c.epi != EPI_NONE && c.epi != EPI_RELU)) return cudaErrorInvalidValue;
        dim3 grid((c.n + 63) / 64, (c.m + 63) / 64);
        // fewer 64 x 64 tiles than SMs: the latency-oriented kernel, with as few neurons per CTA as keeps the grid within one CTA
        // per SM (its ring takes 145..170 KB of shared memory; variant 1 forces the tile kernel: A/B and cross-check)
        const int sms = c.num_sms > 0 ? c.num_sms : 148;
        if (c.variant == 0 && (long long)grid.x * grid.y < sms)
        {
            const long long mt = grid.y;
            if (mt * ((c.n + 3) / 4) <= (long long)sms) return launch_fp32_small<1>(c, stream);
            if (mt * ((c.n + 7) / 8) <= (long long)sms) return launch_fp32_small<2>(c, stream);
            return launch_fp32_small<4>(c, stream);
        }
        gemm_fp32_ordered_kernel<<<grid, 256, 0, stream>>>((const float *)c.a, c.lda, (const float *)c.w, c.ldw,
                                                         (const float *)c.bias, (float *)c.out, c.ldc, c.epi == EPI_RELU, c.m,
                                                         c.n, c.k);
        return cudaGetLastError();
    }
    if (c.variant == 1)
    {
        if (c.kind == GK_BF16) return launch_ref<KIND_BF16>(c, stream);
        if (c.kind == GK_TF32) return launch_ref<KIND_TF32>(c, stream);
        return launch_ref<KIND_I8>(c, stream);
    }
    if (c.kind == GK_BF16)
    {
        if (c.out_type == OUT_BF16) return launch_tc_bn<KIND_BF16, OUT_BF16>(c, stream);
        if (c.out_type == OUT_F32) return launch_tc_bn<KIND_BF16, OUT_F32>(c, stream);
    }
    else if (c.kind == GK_TF32)
    {
        if (c.out_type == OUT_F32) return launch_tc_bn<KIND_TF32, OUT_F32>(c, stream);
        if (c.out_type == OUT_BF16) return launch_tc_bn<KIND_TF32, OUT_BF16>(c, stream);
    }
    else if (c.kind == GK_I8)
    {
        if (c.out_type == OUT_S8) return launch_tc_bn<KIND_I8, OUT_S8>(c, stream);
        if (c.out_type == OUT_S32) return launch_tc_bn<KIND_I8, OUT_S32>(c, stream);
    }
    return cudaErrorInvalidValue;
}

} // namespace nc

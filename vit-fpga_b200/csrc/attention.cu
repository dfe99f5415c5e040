// attention.cu -- multi-head attention core, softmax(q k^T / sqrt(64)) v, per (image, head).
//
// No reference counterpart (the reference has no ViT); semantics follow
// oracle_attention (oracle/oracle_vit.c).  Input is the packed bf16 output of the QKV projection,
// rows [q | k | v] of 3*heads*64 columns per token, so no re-layout kernel runs between the GEMM and
// attention, and the output [token][heads*64] is directly the A operand of the output projection.
//
// Three kernels:
//  (1) attention_tc_kernel (tokens <= 256, i.e. every 224-pixel ViT): tcgen05.  One persistent CTA per SM
//      walks over (image, head) items.  Per item: TMA loads Q (two 128-row tiles), K and V of the head
//      (128B-swizzled, double-buffered across items); S = Q.K^T with UMMA 128 x n_pad x 16 into tensor
//      memory; two softmax warpgroups (one per query tile, thread = query row) read S from TMEM, take the
//      exact row max, write the un-normalised bf16 P back into the same TMEM columns; O = P.V is a UMMA
//      with the A operand in TMEM and V consumed MN-major straight from its row-major [key][64] tile;
//      the warpgroup scales O by 1/rowsum and stores its rows with TMA.  S and P never touch smem or HBM.
//      One MMA-issuing thread per query tile (blocking waits); the two softmax warps of an SM sub-partition
//      take turns in their exp2 pass (template parameter MODE).  The PRODUCT kernel for these sequence lengths is
//      attention_tc16_kernel (sixteen softmax warps, further below).  Measured alternatives of round 1 (256 x 12 x 197, us per launch): two-pass softmax 126; a
//      single-read softmax that keeps the row as fp16 differences in registers 134; P through shared memory
//      with an SS-mode P.V 171; P.V issued in 64-key groups while the softmax is still running 147.
//  (2) attention_tc_long_kernel (tokens > 256, e.g. 577 tokens at 384 pixels): the key-blocked variant of the
//      same building blocks (online softmax across key blocks of <= 256 keys, O accumulated in registers).
//  (3) attention_kernel: flash-style single pass with mma.sync m16n8k16 (bf16 in, fp32 accumulate) -- NOT on
//      the product path; it is the cross-check variant (netcuda_set_gemm_variant(1)) the tests compare with:
//   - one CTA per (query block, head, image); K and V of the head live in shared memory once
//     (197 keys: 2 x 26 KB; 577 keys: 2 x 74 KB), XOR-swizzled in 16-byte chunks so ldmatrix is
//     bank-conflict free; rows >= tokens are zero-filled and masked to -inf before the softmax;
//   - each warp owns 16 query rows; S and P never leave registers (online softmax over 64-key
//     chunks, exp2 with the 1/sqrt(64)*log2(e) scale folded in, warp-shuffle row reductions).
#include "kernels.h"
#include "ptx.cuh"

#include <cstdlib>

namespace nc
{

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float *d, const uint32_t *a, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int ATT_HD = 64;

template <int NW, typename OutT>
__global__ void __launch_bounds__(NW * 32)
attention_kernel(const __nv_bfloat16 *__restrict__ qkv, OutT *__restrict__ out, int tokens, int heads, int tpad)
{
    extern __shared__ __align__(128) uint8_t att_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y;
    const long long b = blockIdx.z;
    const int D = heads * ATT_HD;
    const long long ld = 3LL * D;
    const __nv_bfloat16 *base = qkv + b * tokens * ld;
    const uint32_t ks = smem_u32(att_smem);
    const uint32_t vs = ks + (uint32_t)tpad * 128u;

    // ---- stage K and V of this head: row r, 16-byte chunk c -> r*128 + ((c ^ (r & 7)) << 4)
    for (int idx = threadIdx.x; idx < tpad * 8; idx += NW * 32)
    {
        const int r = idx >> 3, c = idx & 7;
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
        if (r < tokens)
        {
            const __nv_bfloat16 *src = base + r * ld + h * ATT_HD + c * 8;
            cp_async_16(ks + off, src + D);
            cp_async_16(vs + off, src + 2 * D);
        }
        else
        {
            *reinterpret_cast<uint4 *>(att_smem + off) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4 *>(att_smem + (size_t)tpad * 128 + off) = make_uint4(0, 0, 0, 0);
        }
    }

    // ---- Q fragments straight from global (rows past the end re-read the last token; never stored)
    const int q0 = (blockIdx.x * NW + warp) * 16;
    const int g = lane >> 2, tg = lane & 3;
    const int qr0 = min(q0 + g, tokens - 1), qr1 = min(q0 + g + 8, tokens - 1);
    uint32_t qa[4][4];
    {
        const __nv_bfloat16 *q_lo = base + qr0 * ld + h * ATT_HD + tg * 2;
        const __nv_bfloat16 *q_hi = base + qr1 * ld + h * ATT_HD + tg * 2;
#pragma unroll
        for (int s = 0; s < 4; s++)
        {
            qa[s][0] = *reinterpret_cast<const uint32_t *>(q_lo + s * 16);
            qa[s][1] = *reinterpret_cast<const uint32_t *>(q_hi + s * 16);
            qa[s][2] = *reinterpret_cast<const uint32_t *>(q_lo + s * 16 + 8);
            qa[s][3] = *reinterpret_cast<const uint32_t *>(q_hi + s * 16 + 8);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    if (q0 >= tokens) return; // whole warp has no query rows (after the only block-wide barrier)

    const float sl = 0.125f * 1.4426950408889634f; // 1/sqrt(64) * log2(e)
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    const int lrow = lane & 7, lmat = lane >> 3;

    for (int kc = 0; kc < tpad; kc += 64)
    {
        const int nkb = min(64, tpad - kc) >> 3; // 8-key blocks in this chunk (even)
        float s[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; nb++)
        {
            s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.0f;
            if (nb < nkb)
            {
                const int key = kc + nb * 8 + lrow;
                const uint32_t rowaddr = ks + (uint32_t)key * 128u;
#pragma unroll
                for (int h2 = 0; h2 < 2; h2++)
                {
                    uint32_t r0, r1, r2, r3;
                    ldmatrix_x4(rowaddr + (uint32_t)(((4 * h2 + lmat) ^ (key & 7)) << 4), r0, r1, r2, r3);
                    mma_bf16_16816(s[nb], qa[2 * h2], r0, r1);
                    mma_bf16_16816(s[nb], qa[2 * h2 + 1], r2, r3);
                }
            }
        }
        // scale, mask, chunk row-max
        float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
        for (int nb = 0; nb < 8; nb++)
        {
            if (nb < nkb)
            {
                const int key = kc + nb * 8 + tg * 2;
                const bool v0 = key < tokens, v1 = key + 1 < tokens;
                s[nb][0] = v0 ? s[nb][0] * sl : -INFINITY;
                s[nb][1] = v1 ? s[nb][1] * sl : -INFINITY;
                s[nb][2] = v0 ? s[nb][2] * sl : -INFINITY;
                s[nb][3] = v1 ? s[nb][3] * sl : -INFINITY;
                cm0 = fmaxf(cm0, fmaxf(s[nb][0], s[nb][1]));
                cm1 = fmaxf(cm1, fmaxf(s[nb][2], s[nb][3]));
            }
        }
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
        const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1); // finite: every chunk holds a valid key
        const float al0 = ex2_approx(m0 - mn0), al1 = ex2_approx(m1 - mn1);
        m0 = mn0, m1 = mn1;
        l0 *= al0, l1 *= al1;
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            o[i][0] *= al0, o[i][1] *= al0;
            o[i][2] *= al1, o[i][3] *= al1;
        }
        // probabilities + P.V, 16 keys per step
#pragma unroll
        for (int kk = 0; kk < 4; kk++)
        {
            if (2 * kk < nkb)
            {
                uint32_t pa[4];
                {
                    const float p00 = ex2_approx(s[2 * kk][0] - mn0), p01 = ex2_approx(s[2 * kk][1] - mn0);
                    const float p02 = ex2_approx(s[2 * kk][2] - mn1), p03 = ex2_approx(s[2 * kk][3] - mn1);
                    const float p10 = ex2_approx(s[2 * kk + 1][0] - mn0), p11 = ex2_approx(s[2 * kk + 1][1] - mn0);
                    const float p12 = ex2_approx(s[2 * kk + 1][2] - mn1), p13 = ex2_approx(s[2 * kk + 1][3] - mn1);
                    l0 += (p00 + p01) + (p10 + p11);
                    l1 += (p02 + p03) + (p12 + p13);
                    pa[0] = pack_bf16x2(p00, p01);
                    pa[1] = pack_bf16x2(p02, p03);
                    pa[2] = pack_bf16x2(p10, p11);
                    pa[3] = pack_bf16x2(p12, p13);
                }
                const int key = kc + kk * 16 + (lmat & 1) * 8 + lrow;
                const uint32_t rowaddr = vs + (uint32_t)key * 128u;
#pragma unroll
                for (int dp = 0; dp < 4; dp++)
                {
                    uint32_t r0, r1, r2, r3;
                    ldmatrix_x4_trans(rowaddr + (uint32_t)(((2 * dp + (lmat >> 1)) ^ (key & 7)) << 4), r0, r1, r2, r3);
                    mma_bf16_16816(o[2 * dp], pa, r0, r1);
                    mma_bf16_16816(o[2 * dp + 1], pa, r2, r3);
                }
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    OutT *obase = out + b * tokens * (long long)D + h * ATT_HD + tg * 2;
    const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        if constexpr (sizeof(OutT) == 2)
        {
            if (r0 < tokens) *reinterpret_cast<uint32_t *>(obase + (long long)r0 * D + i * 8) = pack_bf16x2(o[i][0] * i0, o[i][1] * i0);
            if (r1 < tokens) *reinterpret_cast<uint32_t *>(obase + (long long)r1 * D + i * 8) = pack_bf16x2(o[i][2] * i1, o[i][3] * i1);
        }
        else
        {
            if (r0 < tokens) *reinterpret_cast<float2 *>(obase + (long long)r0 * D + i * 8) = make_float2(o[i][0] * i0, o[i][1] * i0);
            if (r1 < tokens) *reinterpret_cast<float2 *>(obase + (long long)r1 * D + i * 8) = make_float2(o[i][2] * i1, o[i][3] * i1);
        }
    }
}

template <int NW, typename OutT>
static cudaError_t launch_attention_nw(const __nv_bfloat16 *q, OutT *o, int batch, int tokens, int heads, int tpad, size_t smem, cudaStream_t stream)
{
    // per-device attribute; setting it on every launch costs ~1 us and keeps multi-GPU processes correct
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<NW, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    dim3 grid((tokens + 16 * NW - 1) / (16 * NW), heads, batch);
    attention_kernel<NW, OutT><<<grid, NW * 32, smem, stream>>>(q, o, tokens, heads, tpad);
    return cudaGetLastError();
}

// =====================================================================================================
// tcgen05 kernel
// =====================================================================================================

constexpr int ATC_THREADS = 384;
// Warp roles: producer = warp 0; MMA issuers = warp 1 (query tile 0) and warp 3 (query tile 1); TMEM allocator = warp 2; softmax
// warpgroups = warps 4-7 (query tile 0) and 8-11 (tile 1).
// (The opposite order -- softmax on the low warp ids so that the scheduler favours the MMA / TMA issue -- was measured: 129.5
// vs 126 us per launch, 7.9 vs 7.0 ms per ViT-B step.)
constexpr int ATC_W_PRODUCER = 0, ATC_W_MMA = 1, ATC_W_ALLOC = 2, ATC_W_MMA1 = 3;
constexpr int ATC_Q_BYTES = 128 * 128;  // one query tile: 128 rows x 64 bf16
constexpr int ATC_KV_BYTES = 256 * 128; // up to 256 keys x 64 bf16
constexpr int ATC_BUF_BYTES = 2 * ATC_Q_BYTES + 2 * ATC_KV_BYTES;
constexpr int ATC_OSLAB_BYTES = 32 * 128;         // one softmax warp's 32 output rows x 128 bytes (128B-swizzled TMA-store source)
constexpr int ATC_OFF_OSLAB = 2 * ATC_BUF_BYTES; // 1024-byte aligned
constexpr int ATC_OFF_BARS = ATC_OFF_OSLAB + 8 * ATC_OSLAB_BYTES;
constexpr int ATC_NUM_BARS = 4 + 8 + 8;
constexpr int ATC_OFF_TMEM_PTR = ATC_OFF_BARS + ATC_NUM_BARS * 8;
constexpr int ATC_SMEM = ATC_OFF_TMEM_PTR + 16;
constexpr int ATC_REGION_COLS = 256; // TMEM columns per query tile: S at [0, n_pad), P over [0, n_pad/2), O at [128, 192)
constexpr int ATC_O_COL = 128;

enum : int
{
    KERR_ATT_PRODUCER = 11,
    KERR_ATT_MMA_FULL = 12,
    KERR_ATT_MMA_SFREE = 13,
    KERR_ATT_MMA_PFULL = 14,
    KERR_ATT_WG_SFULL = 15,
    KERR_ATT_WG_OFULL = 16,
    KERR_ATT_WG_TURN = 17,
};

struct AttnTcParams
{
    void *out;    // [batch * tokens][heads * 64], bf16 or (out_f32) fp32
    int batch, tokens, heads;
    int n_pad;    // keys padded to a multiple of 16 (UMMA N of the S tile, UMMA K extent of P.V)
    int n_mtiles; // 1 or 2 query tiles of 128 rows
    int stagger;  // MODE 0: hold the first S of tile 1 back until tile 0 has finished its first softmax
    int out_f32;  // output rows are fp32 (tf32 nets): two 32-column store boxes per warp instead of one 64-column bf16 box
    int split, p1_col, o_col; // attention_tc16_kernel: first key chunk of half 1; TMEM columns of half 1's P and of O | row sums (80 columns)
    int *error_flag;
    long long *debug; // optional [32 items][12 warps][8] clock64 stamps of CTA 0 (NETCUDA_DEBUG_TIMELINE builds; null otherwise)
};

__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); // FMNMX3: one issue slot for two comparisons
    return d;
}

// O rows of one softmax warp (lane = row, 64 fp32 columns each, scaled by inv) -> global memory through the warp's 128B-swizzled
// slab and 3-D TMA stores (rows past the image's last token are clipped by the tensor map): one 64-column bf16 box, or two
// 32-column fp32 boxes.  Per-lane-row global stores -- 8 x 16 bytes to 32 different lines per instruction -- kept the warp in
// this phase for ~1700 cycles of its ~8700-cycle chain per item.
__device__ __forceinline__ void store_o_rows(const uint32_t *o, float inv, uint8_t *oslab, uint32_t oslab_addr, const CUtensorMap *map, bool f32,
                                             int col0, int row0, int img, int lane)
{
    uint8_t *orow = oslab + lane * 128;
    if (!f32)
    {
        if (lane == 0) tma_store_wait_read(); // the previous item's store has finished reading the slab
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            uint4 pk;
            pk.x = pack_bf16x2(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
            pk.y = pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
            pk.z = pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
            pk.w = pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
            *reinterpret_cast<uint4 *>(orow + ((j ^ (lane & 7)) << 4)) = pk;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0)
        {
            tma_store_3d(map, oslab_addr, col0, row0, img);
            tma_store_commit();
        }
        return;
    }
#pragma unroll
    for (int half = 0; half < 2; half++)
    {
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            const uint32_t *v = o + half * 32 + 4 * j;
            const uint4 pk = make_uint4(__float_as_uint(__uint_as_float(v[0]) * inv), __float_as_uint(__uint_as_float(v[1]) * inv),
                                        __float_as_uint(__uint_as_float(v[2]) * inv), __float_as_uint(__uint_as_float(v[3]) * inv));
            *reinterpret_cast<uint4 *>(orow + ((j ^ (lane & 7)) << 4)) = pk;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0)
        {
            tma_store_3d(map, oslab_addr, col0 + half * 32, row0, img);
            tma_store_commit();
        }
    }
}

// MODE 0: one MMA-issuing thread polls both query tiles (round-1 kernel).
// MODE 1: one issuing thread per query tile (warps 1 and 3), each in blocking mbarrier waits: no polling round (six test_waits of
//         ~150 cycles each) between "P is ready" and the P.V issue, or between "O has been read" and the next S.
// MODE 2: MODE 1 + the two softmax warps of one SM sub-partition (query tile 0 / tile 1, same TMEM lane quarter) take turns in
//         their exponential pass.
// Measured (512 x 12 x 197, us per launch in isolation): MODE 0 200, MODE 1 194, MODE 2 187.  Tried on top of MODE 1 and dropped
// (tools/attn_timeline.py shows why): a deeper tcgen05.ld pipeline with setmaxnreg 232 / 40 and the 32 MUFUs of a chunk issued
// before anything consumes them (193); a quarter or half of the exponentials as a degree-3 Cody-Waite polynomial on the FMA pipe
// (192 / 200: a polynomial costs ~7 FMA-pipe issue slots of 2 cycles against the MUFU's 8 cycles, and a single warp cannot overlap
// the two pipes anyway -- see attention_tc16_kernel below, which is the product kernel for these sequence lengths).
template <int MODE>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                    const __grid_constant__ CUtensorMap tma_out, const AttnTcParams p)
{
    extern __shared__ __align__(1024) uint8_t atc_smem[];
    const uint32_t base = smem_u32(atc_smem);
    if ((base & 1023u) != 0)
    {
        if (threadIdx.x == 0 && p.error_flag) atomicExch(p.error_flag, KERR_SMEM_ALIGN);
        return;
    }
    const uint32_t bars = base + ATC_OFF_BARS;
    auto full_bar = [&](int b) { return bars + 8u * b; };
    auto empty_bar = [&](int b) { return bars + 8u * (2 + b); };
    auto sfull_bar = [&](int t) { return bars + 8u * (4 + t); };
    auto pfull_bar = [&](int t) { return bars + 8u * (6 + t); };
    auto ofull_bar = [&](int t) { return bars + 8u * (8 + t); };
    auto sfree_bar = [&](int t) { return bars + 8u * (10 + t); };
    auto turn_bar = [&](int t, int q) { return bars + 8u * (12 + t * 4 + q); }; // MODE 2: "tile t's warp of lane quarter q may run its exp2 pass"
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(atc_smem + ATC_OFF_TMEM_PTR);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31; // (uniform: the MMA issuers' descriptors stay in uniform registers, see ptx.cuh)
    const int items = p.batch * p.heads;
    const int D = p.heads * ATT_HD;
#ifdef NETCUDA_DEBUG_TIMELINE
    long long *const dbg = p.debug;
#else
    constexpr long long *dbg = nullptr;
#endif

    if (warp == ATC_W_PRODUCER && lane == 0)
    {
        tma_prefetch_desc(&tma_q);
        tma_prefetch_desc(&tma_kv);
    }
    if (warp == ATC_W_MMA && lane == 0)
    {
        for (int b = 0; b < 2; b++)
        {
            mbar_init(full_bar(b), 1);
            mbar_init(empty_bar(b), MODE == 0 ? 1 : p.n_mtiles); // MODE >= 1: one commit per tile's issuer
            mbar_init(sfull_bar(b), 1);
            // one arrive per softmax warp that owns at least one real query row (197 tokens: 4 warps for tile 0, 3 for tile 1)
            const int rows_b = min(max(p.tokens - b * 128, 1), 128);
            mbar_init(pfull_bar(b), (rows_b + 31) / 32);
            mbar_init(ofull_bar(b), 1);
            mbar_init(sfree_bar(b), (rows_b + 31) / 32);
            for (int q = 0; q < 4; q++) mbar_init(turn_bar(b, q), 1);
        }
        fence_barrier_init();
    }
    if (warp == ATC_W_ALLOC)
    {
        tmem_alloc(base + ATC_OFF_TMEM_PTR, 512);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    griddep_launch_dependents();
    griddep_wait(); // the qkv matrix is the previous kernel's output
    if (warp < 4)
    {
    if (warp == ATC_W_PRODUCER)
    {
        // ===================== TMA producer =====================
        if (lane == 0)
        {
            const uint32_t tx = (uint32_t)(p.n_mtiles * ATC_Q_BYTES + 2 * p.n_pad * 128);
            int it = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, it++)
            {
                const int buf = it & 1;
                const int b = item / p.heads, h = item - b * p.heads;
                mbar_wait(empty_bar(buf), ((uint32_t)(it >> 1) & 1u) ^ 1u, p.error_flag, KERR_ATT_PRODUCER);
                if (dbg && blockIdx.x == 0 && it < 32) dbg[(it * 12 + ATC_W_PRODUCER) * 8] = clock64();
                mbar_arrive_expect_tx(full_bar(buf), tx);
                const uint32_t dst = base + buf * ATC_BUF_BYTES;
                const int row = b * p.tokens;
                for (int t = 0; t < p.n_mtiles; t++) tma_load_2d(dst + t * ATC_Q_BYTES, &tma_q, full_bar(buf), h * ATT_HD, row + t * 128);
                tma_load_2d(dst + 2 * ATC_Q_BYTES, &tma_kv, full_bar(buf), D + h * ATT_HD, row);
                tma_load_2d(dst + 2 * ATC_Q_BYTES + ATC_KV_BYTES, &tma_kv, full_bar(buf), 2 * D + h * ATT_HD, row);
            }
        }
    }
    else if (MODE >= 1 && (warp == ATC_W_MMA || warp == ATC_W_MMA1))
    {
        // ===================== MMA issuer of one query tile (blocking waits) =====================
        const int t = warp == ATC_W_MMA ? 0 : 1;
        if (lane == 0 && t < p.n_mtiles)
        {
            const uint32_t idesc_s = umma_idesc(1, 1, 128, (uint32_t)p.n_pad);
            const uint32_t idesc_o = umma_idesc(1, 1, 128, ATT_HD) | UMMA_IDESC_B_MN_MAJOR;
            const int ksteps = p.n_pad / 16;
            const uint32_t region = tmem_base + t * ATC_REGION_COLS;
            int it = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, it++)
            {
                const int buf = it & 1;
                const uint32_t sm = base + buf * ATC_BUF_BYTES;
                // S_t = Q_t . K^T: K / Q landed, and the previous item's O (same TMEM region) has been read
                mbar_wait(full_bar(buf), (uint32_t)(it >> 1) & 1u, p.error_flag, KERR_ATT_MMA_FULL);
                mbar_wait(sfree_bar(t), ((uint32_t)it & 1u) ^ 1u, p.error_flag, KERR_ATT_MMA_SFREE);
                tcgen05_fence_after();
                if (dbg && blockIdx.x == 0 && it < 32) dbg[(it * 12 + warp) * 8 + 3] = clock64();
                {
                    const uint64_t q_desc = umma_smem_desc_sw128(sm + t * ATC_Q_BYTES);
                    const uint64_t k_desc = umma_smem_desc_sw128(sm + 2 * ATC_Q_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; k++) umma_ss<KIND_BF16>(region, q_desc + 2u * k, k_desc + 2u * k, idesc_s, k != 0 ? 1u : 0u);
                    tcgen05_commit(sfull_bar(t));
                }
                if (dbg && blockIdx.x == 0 && it < 32) dbg[(it * 12 + warp) * 8 + 0] = clock64();
                // O_t = P_t . V.  V tile [key][64] read MN-major: one UMMA K step = 16 keys = 2048 bytes
                mbar_wait(pfull_bar(t), (uint32_t)it & 1u, p.error_flag, KERR_ATT_MMA_PFULL);
                tcgen05_fence_after();
                if (dbg && blockIdx.x == 0 && it < 32) dbg[(it * 12 + warp) * 8 + 1] = clock64();
                {
                    const uint64_t v_desc = umma_smem_desc_sw128(sm + 2 * ATC_Q_BYTES + ATC_KV_BYTES);
                    for (int k = 0; k < ksteps; k++)
                        umma_ts_bf16(region + ATC_O_COL, region + 8u * k, v_desc + (uint64_t)(128u * k), idesc_o, k != 0 ? 1u : 0u);
                    tcgen05_commit(ofull_bar(t));
                    tcgen05_commit(empty_bar(buf)); // this tile is done with the smem buffer (the producer waits for every tile's commit)
                }
                if (dbg && blockIdx.x == 0 && it < 32) dbg[(it * 12 + warp) * 8 + 2] = clock64();
            }
        }
    }
    else if (MODE == 0 && warp == ATC_W_MMA)
    {
        // ===================== MMA issuer (polling both tiles) =====================
        // One thread serves both query tiles.  The two softmax warpgroups run independently of each other, so the
        // issuer does not follow a fixed order: it polls, per tile, "S of the next item may start" (K/Q landed,
        // previous O of this tile drained) and "P of the current item is ready" and issues whichever is.
        if (lane == 0)
        {
            const uint32_t idesc_s = umma_idesc(1, 1, 128, (uint32_t)p.n_pad);
            const uint32_t idesc_o = umma_idesc(1, 1, 128, ATT_HD) | UMMA_IDESC_B_MN_MAJOR;
            const int ksteps = p.n_pad / 16;
            const int n_it = blockIdx.x < items ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
            int it_s[2] = {0, 0}, it_pv[2] = {0, 0};
            if (p.n_mtiles == 1) it_s[1] = it_pv[1] = n_it;
            long long t0 = clock64();
            while (it_pv[0] < n_it || it_pv[1] < n_it)
            {
                bool progress = false;
#pragma unroll
                for (int t = 0; t < 2; t++)
                {
                    // S_t = Q_t . K^T of item it_s[t]: the P of the previous item (same columns) must have been consumed.
                    // The very first S_1 is held back until tile 0 has finished its first softmax: the two warpgroups then
                    // stay half a period apart, so that one is in its exp2 (MUFU-bound) pass while the other one waits for
                    // its MMAs, reduces the row max or stores O, instead of both fighting for the MUFU at the same time.
                    if (it_s[t] < n_it && it_pv[t] == it_s[t] && !(p.stagger && t == 1 && it_s[1] == 0 && it_pv[0] == 0))
                    {
                        const int i = it_s[t], buf = i & 1;
                        if (mbar_test_wait(full_bar(buf), (uint32_t)(i >> 1) & 1u) && mbar_test_wait(sfree_bar(t), ((uint32_t)i & 1u) ^ 1u))
                        {
                            tcgen05_fence_after();
                            const uint32_t sm = base + buf * ATC_BUF_BYTES;
                            const uint64_t q_desc = umma_smem_desc_sw128(sm + t * ATC_Q_BYTES);
                            const uint64_t k_desc = umma_smem_desc_sw128(sm + 2 * ATC_Q_BYTES);
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                umma_ss<KIND_BF16>(tmem_base + t * ATC_REGION_COLS, q_desc + 2u * k, k_desc + 2u * k, idesc_s, k != 0 ? 1u : 0u);
                            tcgen05_commit(sfull_bar(t));
                            if (dbg && blockIdx.x == 0 && i < 32) dbg[(i * 12 + ATC_W_MMA) * 8 + t] = clock64();
                            it_s[t]++;
                            progress = true;
                        }
                    }
                    // O_t = P_t . V of item it_pv[t].  V tile [key][64] read MN-major: one UMMA K step = 16 keys = 2048 bytes
                    if (it_pv[t] < it_s[t])
                    {
                        const int i = it_pv[t], buf = i & 1;
                        if (mbar_test_wait(pfull_bar(t), (uint32_t)i & 1u))
                        {
                            tcgen05_fence_after();
                            const uint64_t v_desc = umma_smem_desc_sw128(base + buf * ATC_BUF_BYTES + 2 * ATC_Q_BYTES + ATC_KV_BYTES);
                            const uint32_t region = tmem_base + t * ATC_REGION_COLS;
                            for (int k = 0; k < ksteps; k++)
                                umma_ts_bf16(region + ATC_O_COL, region + 8u * k, v_desc + (uint64_t)(128u * k), idesc_o, k != 0 ? 1u : 0u);
                            tcgen05_commit(ofull_bar(t));
                            if (dbg && blockIdx.x == 0 && i < 32) dbg[(i * 12 + ATC_W_MMA) * 8 + 2 + t] = clock64();
                            it_pv[t]++;
                            // both tiles are past item i: every MMA that reads its smem buffer has been issued
                            if (it_pv[t ^ 1] > i) tcgen05_commit(empty_bar(buf));
                            progress = true;
                        }
                    }
                }
                if (progress)
                    t0 = clock64();
                else if (clock64() - t0 > 8000000000LL)
                {
                    if (p.error_flag) atomicExch(p.error_flag, KERR_ATT_MMA_FULL);
                    __threadfence_system();
                    __trap();
                }
            }
        }
    }
    }
    else
    {
    if (((warp - 4) >> 2) < p.n_mtiles && ((warp - 4) >> 2) * 128 + (warp & 3) * 32 < p.tokens)
    {
        // ===================== softmax + output warpgroups (one per query tile) =====================
        const int t = (warp - 4) >> 2; // query tile
        const int q = warp & 3;        // TMEM lane quarter
        const uint32_t region = tmem_base + ((uint32_t)(q * 32) << 16) + t * ATC_REGION_COLS;
        const float sl = 0.125f * 1.4426950408889634f; // 1/sqrt(64) * log2(e)
        uint8_t *oslab = atc_smem + ATC_OFF_OSLAB + (warp - 4) * ATC_OSLAB_BYTES;
        const uint32_t oslab_addr = base + ATC_OFF_OSLAB + (warp - 4) * ATC_OSLAB_BYTES;
        const int nfull = p.tokens >> 5, tail = p.tokens & 31; // full 32-key chunks, keys in the ragged last chunk
        const int nchunks = nfull + (tail ? 1 : 0);
        // MODE 2: the other tile's warp on this SM sub-partition exists (it owns at least one query row)
        const bool partner = MODE == 2 && p.n_mtiles == 2 && (t ^ 1) * 128 + q * 32 < p.tokens;
        int it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, it++)
        {
            const uint32_t ph = (uint32_t)it & 1u;
            const int b = item / p.heads, h = item - b * p.heads;
            auto stamp = [&](int slot) {
                if (dbg && blockIdx.x == 0 && lane == 0 && it < 32) dbg[(it * 12 + warp) * 8 + slot] = clock64();
            };
            stamp(0);
            mbar_wait(sfull_bar(t), ph, p.error_flag, KERR_ATT_WG_SFULL);
            tcgen05_fence_after();
            stamp(1);

            // ---- pass 1: exact row maximum over the valid keys ----
            float mx = -INFINITY;
            auto reduce = [&](const uint32_t *v, int c) {
                if (c < nfull)
                {
                    float m0 = fmax3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
                    float m1 = fmax3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
                    float m2 = fmax3(__uint_as_float(v[6]), __uint_as_float(v[7]), __uint_as_float(v[8]));
                    float m3 = fmax3(__uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
#pragma unroll
                    for (int j = 12; j < 28; j += 8)
                    {
                        m0 = fmax3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                        m1 = fmax3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        m2 = fmax3(m2, __uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
                        m3 = fmax3(m3, __uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                    }
                    m0 = fmax3(m0, __uint_as_float(v[28]), __uint_as_float(v[29]));
                    m1 = fmax3(m1, __uint_as_float(v[30]), __uint_as_float(v[31]));
                    mx = fmax3(mx, fmaxf(m0, m1), fmaxf(m2, m3));
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if (j < tail) mx = fmaxf(mx, __uint_as_float(v[j]));
                }
            };
            {
                // (chunk c + 1 is in flight while c is reduced)
                uint32_t va[32], vb[32];
                tmem_ld_32x32(region, va);
                for (int c = 0; c < nchunks; c += 2)
                {
                    tmem_ld_wait();
                    if (c + 1 < nchunks) tmem_ld_32x32(region + (c + 1) * 32, vb);
                    reduce(va, c);
                    if (c + 1 < nchunks)
                    {
                        tmem_ld_wait();
                        if (c + 2 < nchunks) tmem_ld_32x32(region + (c + 2) * 32, va);
                        reduce(vb, c + 1);
                    }
                }
            }

            stamp(2);
            // MODE 2: wait for this warp's turn on the MUFU (tile 0 goes first; the very first wait of a fresh barrier passes)
            if (partner) mbar_wait(turn_bar(t, q), t == 0 ? (ph ^ 1u) : ph, p.error_flag, KERR_ATT_WG_TURN);
            stamp(7);
            // ---- pass 2: p = 2^((s - max) * scale); P (bf16) overwrites the first half of the S columns it came from ----
            const float msc = mx * sl;
            float sum0 = 0.0f, sum1 = 0.0f;
            {
                uint32_t va[32], vb[32];
                auto expo = [&](const uint32_t *v, int c) {
                    uint32_t w[16];
                    if (c < nfull)
                    {
#pragma unroll
                        for (int j = 0; j < 16; j++)
                        {
                            const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), sl, -msc));
                            const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sl, -msc));
                            sum0 += p0, sum1 += p1;
                            w[j] = pack_bf16x2(p0, p1);
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int j = 0; j < 16; j++)
                        {
                            const float p0 = (2 * j < tail) ? ex2_approx(fmaf(__uint_as_float(v[2 * j]), sl, -msc)) : 0.0f;
                            const float p1 = (2 * j + 1 < tail) ? ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sl, -msc)) : 0.0f;
                            sum0 += p0, sum1 += p1;
                            w[j] = pack_bf16x2(p0, p1);
                        }
                    }
                    tmem_st_32x16(region + c * 16, w);
                };
                tmem_ld_32x32(region, va);
                for (int c = 0; c < nchunks; c += 2)
                {
                    tmem_ld_wait();
                    if (c + 1 < nchunks) tmem_ld_32x32(region + (c + 1) * 32, vb);
                    expo(va, c);
                    if (c + 1 < nchunks)
                    {
                        tmem_ld_wait();
                        if (c + 2 < nchunks) tmem_ld_32x32(region + (c + 2) * 32, va);
                        expo(vb, c + 1);
                    }
                }
            }
            const float sum = sum0 + sum1;
            if (partner)
            {
                __syncwarp();
                if (lane == 0) mbar_arrive(turn_bar(t ^ 1, q)); // the partner's turn
            }
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pfull_bar(t));
            stamp(3);

            // ---- O = P.V / rowsum ----
            mbar_wait(ofull_bar(t), ph, p.error_flag, KERR_ATT_WG_OFULL);
            tcgen05_fence_after();
            stamp(4);
            uint32_t o[64];
            tmem_ld_32x32(region + ATC_O_COL, o);
            tmem_ld_32x32(region + ATC_O_COL + 32, o + 32);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sfree_bar(t)); // region t may be overwritten by the next item's S
            stamp(6);
            store_o_rows(o, 1.0f / sum, oslab, oslab_addr, &tma_out, p.out_f32 != 0, h * ATT_HD, t * 128 + q * 32, b, lane);
            stamp(5);
        }
        if (lane == 0) tma_store_wait_all(); // the slab is read, and the rows are written, before the CTA goes away
    }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == ATC_W_ALLOC)
    {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// =====================================================================================================
// tcgen05 kernel, tokens <= 256, SIXTEEN softmax warps (attention_tc16_kernel)
// =====================================================================================================
// Measured on the kernel above (tools/attn_timeline.py, tools/micro/mufu_bw.cu, B200): a CTA's period (~7.3 k cycles per item) is
// the instruction time of ONE softmax warp per query row block, and an SM sub-partition runs only two such warps.  A single warp
// cannot overlap its MUFU work with its FMA-pipe work -- 32 EX2 alone take 298 cycles, the 112 FFMA / FADD / F2FP of a 32-key
// chunk 215, both together 510 -- while two warps of a sub-partition in the same phase get through a chunk each in 616 (308 per
// chunk and sub-partition), three in 276, four in 267: the MUFU's 256.  So the softmax needs more warps per sub-partition, not
// fewer instructions per warp:
//   * 20 warps: producer, one MMA issuer per query tile, TMEM allocator, and 16 softmax warps -- two per (query tile, TMEM lane
//     quarter): "half 0" owns keys [0, 128), "half 1" keys [128, n_pad).  Four softmax warps per sub-partition.
//   * The two warps of a row block exchange their partial row maxima once per item through 256 bytes of shared memory and a
//     64-thread named barrier; nothing else is exchanged:
//   * the row sums come out of the tensor core: next to O = P.V (N = 64) the issuer runs P.1 (N = 16) against a tile of bf16 ones,
//     so the 32 FADDs per chunk disappear and the sum is the sum of exactly the bf16 values that multiply V;
//   * the half-1 warp (three key chunks of a 197-token image against four) also reads O, divides by the row sum and stores the
//     rows; the half-0 warp goes straight on to the next item;
//   * the two query tiles take turns in the exponential pass (per lane quarter): with four warps of a sub-partition in it at once
//     the pass takes 4 x 256 cycles per chunk for everybody, and every other phase of BOTH tiles -- S, the max pass, P.V, O -- is then
//     exposed (measured: period 8.7 k cycles); alternating, one tile's MMAs and max pass run under the other tile's exponentials.
// TMEM per query tile (256 columns): S at [0, n_pad); P of half 0 over [0, 64), of half 1 over [128, 128 + 16 chunks) -- each over
// S columns its own warp has already consumed; O at [64, 128) and the row sums at [192, 208): written by the MMAs only after both
// halves have arrived.
constexpr int A16_THREADS = 640;
constexpr int A16_OFF_OSLAB = 2 * ATC_BUF_BYTES;    // 1024-byte aligned; one 32 x 128-byte slab per half-1 warp (as in attention_tc_kernel)
constexpr int A16_OFF_ONES = A16_OFF_OSLAB + 8 * ATC_OSLAB_BYTES; // 16 rows x 128 bytes of bf16 1.0 (1024-byte aligned)
constexpr int A16_OFF_BARS = A16_OFF_ONES + 2048;
constexpr int A16_NUM_BARS = 12 + 8;
constexpr int A16_OFF_TMEM_PTR = A16_OFF_BARS + A16_NUM_BARS * 8;
constexpr int A16_SMEM = A16_OFF_TMEM_PTR + 16;
static_assert(A16_SMEM <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
// Key split and TMEM columns (AttnTcParams::split / p1_col / o_col, chosen by the launcher): P of half 0 over [0, 16 split), P of half 1
// from p1_col on, O at [o_col, o_col + 64) and the row sums at [o_col + 64, o_col + 80).  Up to 224 keys: split 3 (three full chunks
// against three full chunks + the ragged one for 197 tokens), p1_col 112, o_col 176; 225..256 keys: split 4, p1_col 144, o_col 64.

template <bool DB> // DB: keep the next chunk's tcgen05.ld in flight while the current one is processed
__global__ void __launch_bounds__(A16_THREADS, 1)
attention_tc16_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                      const __grid_constant__ CUtensorMap tma_out, const AttnTcParams p)
{
    extern __shared__ __align__(1024) uint8_t a16_smem[];
    const uint32_t base = smem_u32(a16_smem);
    if ((base & 1023u) != 0)
    {
        if (threadIdx.x == 0 && p.error_flag) atomicExch(p.error_flag, KERR_SMEM_ALIGN);
        return;
    }
    const uint32_t bars = base + A16_OFF_BARS;
    auto full_bar = [&](int b) { return bars + 8u * b; };
    auto empty_bar = [&](int b) { return bars + 8u * (2 + b); };
    auto sfull_bar = [&](int t) { return bars + 8u * (4 + t); };
    auto pfull_bar = [&](int t) { return bars + 8u * (6 + t); };
    auto ofull_bar = [&](int t) { return bars + 8u * (8 + t); };
    auto sfree_bar = [&](int t) { return bars + 8u * (10 + t); };
    auto turn_bar = [&](int t, int q) { return bars + 8u * (12 + t * 4 + q); }; // "tile t's warps of lane quarter q may run their exp2 pass"
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(a16_smem + A16_OFF_TMEM_PTR);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31; // (uniform: the MMA issuers' descriptors stay in uniform registers, see ptx.cuh)
    const int items = p.batch * p.heads;
    const int D = p.heads * ATT_HD;
    const int nfull = p.tokens >> 5, tail = p.tokens & 31; // full 32-key chunks, keys in the ragged last chunk
    const int nchunks = nfull + (tail ? 1 : 0);
    const bool two_halves = nchunks > p.split;              // does half 1 own any keys?
#ifdef NETCUDA_DEBUG_TIMELINE
    long long *const dbg = p.debug; // [32 items][20 warps][8] clock64 stamps of CTA 0
#else
    constexpr long long *dbg = nullptr;
#endif

    if (warp == ATC_W_PRODUCER && lane == 0)
    {
        tma_prefetch_desc(&tma_q);
        tma_prefetch_desc(&tma_kv);
    }
    if (warp == ATC_W_MMA && lane == 0)
    {
        for (int b = 0; b < 2; b++)
        {
            mbar_init(full_bar(b), 1);
            mbar_init(empty_bar(b), p.n_mtiles); // one commit per tile's issuer
            mbar_init(sfull_bar(b), 1);
            // one arrive per softmax warp that owns at least one real query row (and, for half 1, at least one key)
            const int rows_b = min(max(p.tokens - b * 128, 1), 128);
            const uint32_t blocks = (uint32_t)((rows_b + 31) / 32);
            mbar_init(pfull_bar(b), blocks * (two_halves ? 2u : 1u));
            mbar_init(ofull_bar(b), 1);
            mbar_init(sfree_bar(b), blocks); // the half-1 warps read O
            for (int q = 0; q < 4; q++) mbar_init(turn_bar(b, q), two_halves ? 2u : 1u);
        }
        fence_barrier_init();
    }
    if (warp == ATC_W_ALLOC)
    {
        tmem_alloc(base + A16_OFF_TMEM_PTR, 512);
        tmem_relinquish();
    }
    // the tile of ones behind the row-sum MMA: every element is 1.0, so swizzle and majorness of the descriptor do not matter
    for (int i = threadIdx.x; i < 2048 / 4; i += A16_THREADS) reinterpret_cast<uint32_t *>(a16_smem + A16_OFF_ONES)[i] = 0x3F803F80u;
    fence_proxy_async_smem(); // generic-proxy writes -> visible to the tensor core's async-proxy reads
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    griddep_launch_dependents();
    griddep_wait(); // the qkv matrix is the previous kernel's output

    if (warp < 4)
    {
        // 640 x 96 registers at launch; the four service warps keep 40 each, the 16 softmax warps grow to 104
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == ATC_W_PRODUCER)
        {
            // ===================== TMA producer =====================
            if (lane == 0)
            {
                const uint32_t tx = (uint32_t)(p.n_mtiles * ATC_Q_BYTES + 2 * p.n_pad * 128);
                int it = 0;
                for (int item = blockIdx.x; item < items; item += gridDim.x, it++)
                {
                    const int buf = it & 1;
                    const int b = item / p.heads, h = item - b * p.heads;
                    mbar_wait(empty_bar(buf), ((uint32_t)(it >> 1) & 1u) ^ 1u, p.error_flag, KERR_ATT_PRODUCER);
                    mbar_arrive_expect_tx(full_bar(buf), tx);
                    const uint32_t dst = base + buf * ATC_BUF_BYTES;
                    const int row = b * p.tokens;
                    for (int t = 0; t < p.n_mtiles; t++) tma_load_2d(dst + t * ATC_Q_BYTES, &tma_q, full_bar(buf), h * ATT_HD, row + t * 128);
                    tma_load_2d(dst + 2 * ATC_Q_BYTES, &tma_kv, full_bar(buf), D + h * ATT_HD, row);
                    tma_load_2d(dst + 2 * ATC_Q_BYTES + ATC_KV_BYTES, &tma_kv, full_bar(buf), 2 * D + h * ATT_HD, row);
                }
            }
        }
        else if (warp == ATC_W_MMA || warp == ATC_W_MMA1)
        {
            // ===================== MMA issuer of one query tile (blocking waits) =====================
            const int t = warp == ATC_W_MMA ? 0 : 1;
            if (lane == 0 && t < p.n_mtiles)
            {
                const uint32_t idesc_s = umma_idesc(1, 1, 128, (uint32_t)p.n_pad);
                // O and the row sums in ONE MMA per 16 keys (the A operand, P, is read from tensor memory once): B is MN-major with
                // N = 80 = the 64 columns of V plus a second MN atom whose 16 used columns are ones.  The second atom is addressed through
                // the descriptor's leading-dimension byte offset, set per step so that it always lands on the tile of ones.
                const uint32_t idesc_o = umma_idesc(1, 1, 128, ATT_HD + 16) | UMMA_IDESC_B_MN_MAJOR;
                const int ksteps = p.n_pad / 16;
                const uint32_t region = tmem_base + t * ATC_REGION_COLS;
                int it = 0;
                for (int item = blockIdx.x; item < items; item += gridDim.x, it++)
                {
                    const int buf = it & 1;
                    const uint32_t sm = base + buf * ATC_BUF_BYTES;
                    // S_t = Q_t . K^T: K / Q landed, and the previous item's O and row sums (same TMEM region) have been read
                    mbar_wait(full_bar(buf), (uint32_t)(it >> 1) & 1u, p.error_flag, KERR_ATT_MMA_FULL);
                    mbar_wait(sfree_bar(t), ((uint32_t)it & 1u) ^ 1u, p.error_flag, KERR_ATT_MMA_SFREE);
                    tcgen05_fence_after();
                    if (dbg && blockIdx.x == 0 && it < 32) dbg[(it * 20 + warp) * 8 + 3] = clock64();
                    {
                        const uint64_t q_desc = umma_smem_desc_sw128(sm + t * ATC_Q_BYTES);
                        const uint64_t k_desc = umma_smem_desc_sw128(sm + 2 * ATC_Q_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; k++) umma_ss<KIND_BF16>(region, q_desc + 2u * k, k_desc + 2u * k, idesc_s, k != 0 ? 1u : 0u);
                        tcgen05_commit(sfull_bar(t));
                        if (dbg && blockIdx.x == 0 && it < 32) dbg[(it * 20 + warp) * 8 + 0] = clock64();
                    }
                    // O_t = P_t . V and rowsum_t = P_t . 1, 16 keys per step; P of keys >= 128 lives at column A16_P1_COL
                    mbar_wait(pfull_bar(t), (uint32_t)it & 1u, p.error_flag, KERR_ATT_MMA_PFULL);
                    tcgen05_fence_after();
                    {
                        const uint32_t v_addr = sm + 2 * ATC_Q_BYTES + ATC_KV_BYTES;
                        if (dbg && blockIdx.x == 0 && it < 32) dbg[(it * 20 + warp) * 8 + 1] = clock64();
                        for (int k = 0; k < ksteps; k++)
                        {
                            const uint32_t a_col = region + (k < 2 * p.split ? 8u * k : (uint32_t)p.p1_col + 8u * (k - 2 * p.split));
                            const uint32_t vk = v_addr + 2048u * k; // 16 keys x 128 bytes per step
                            uint64_t v_desc = umma_smem_desc_sw128(vk) & ~((uint64_t)0x3FFF << 16);
                            v_desc |= (uint64_t)(((base + A16_OFF_ONES - vk) >> 4) & 0x3FFFu) << 16; // LBO: V's MN atom -> the atom of ones
                            umma_ts_bf16(region + p.o_col, a_col, v_desc, idesc_o, k != 0 ? 1u : 0u);
                        }
                        tcgen05_commit(ofull_bar(t));
                        tcgen05_commit(empty_bar(buf)); // this tile is done with the smem buffer
                        if (dbg && blockIdx.x == 0 && it < 32) dbg[(it * 20 + warp) * 8 + 2] = clock64();
                    }
                }
            }
        }
    }
    else
    {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int idx = warp - 4;
        const int q = idx & 3;           // TMEM lane quarter (= warp % 4)
        const int half = (idx >> 2) & 1; // key half
        const int t = idx >> 3;          // query tile
        if (t < p.n_mtiles && t * 128 + q * 32 < p.tokens)
        {
            // ===================== softmax + output: one warp per (query tile, lane quarter, key half) =====================
            const uint32_t region = tmem_base + ((uint32_t)(q * 32) << 16) + t * ATC_REGION_COLS;
            const float sl = 0.125f * 1.4426950408889634f; // 1/sqrt(64) * log2(e)
            const int slab = t * 4 + q; // (used by the half-1 warp)
            uint8_t *oslab = a16_smem + A16_OFF_OSLAB + slab * ATC_OSLAB_BYTES;
            const uint32_t oslab_addr = base + A16_OFF_OSLAB + slab * ATC_OSLAB_BYTES;
            const bool works = half == 0 || two_halves; // does this warp own keys?
            const bool partner = p.n_mtiles == 2 && (t ^ 1) * 128 + q * 32 < p.tokens; // the other tile has warps on this sub-partition
            const int c_lo = half ? p.split : 0;
            const int c_hi = half ? nchunks : min(nchunks, p.split);
            const uint32_t p_col = half ? (uint32_t)p.p1_col : 0u; // where this half's P goes: chunk c -> p_col + 16 (c - c_lo)
            const uint32_t pair_bar = 1u + (uint32_t)(t * 4 + q);    // named barrier of this row block's two warps
            int it = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, it++)
            {
                const uint32_t ph = (uint32_t)it & 1u;
                const int buf = it & 1;
                const int b = item / p.heads, h = item - b * p.heads;
                auto stamp = [&](int slot) {
                    if (dbg && blockIdx.x == 0 && lane == 0 && it < 32) dbg[(it * 20 + warp) * 8 + slot] = clock64();
                };
                stamp(0);
                if (works)
                {
                mbar_wait(sfull_bar(t), ph, p.error_flag, KERR_ATT_WG_SFULL);
                tcgen05_fence_after();
                stamp(1);

                // ---- pass 1: row maximum over this half's keys (chunk c + 1 in flight while c is reduced), then the exact maximum
                // through the partner ----
                float mx = -INFINITY;
                {
                    uint32_t va[32], vb[32];
                    auto reduce = [&](const uint32_t *v, int c) {
                        if (c < nfull)
                        {
                            float m0 = fmax3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
                            float m1 = fmax3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
                            float m2 = fmax3(__uint_as_float(v[6]), __uint_as_float(v[7]), __uint_as_float(v[8]));
                            float m3 = fmax3(__uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
#pragma unroll
                            for (int j = 12; j < 28; j += 8)
                            {
                                m0 = fmax3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                                m1 = fmax3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                                m2 = fmax3(m2, __uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
                                m3 = fmax3(m3, __uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                            }
                            m0 = fmax3(m0, __uint_as_float(v[28]), __uint_as_float(v[29]));
                            m1 = fmax3(m1, __uint_as_float(v[30]), __uint_as_float(v[31]));
                            mx = fmax3(mx, fmaxf(m0, m1), fmaxf(m2, m3));
                        }
                        else
                        {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (j < tail) mx = fmaxf(mx, __uint_as_float(v[j]));
                        }
                    };
                    // (loads are unconditional -- chunk c + 1 <= 7 stays inside the tile's 256 columns -- so that no buffer is
                    // written on one path only: ptxas spills conditionally loaded tcgen05.ld destinations)
                    if constexpr (DB)
                    {
                        tmem_ld_32x32(region + c_lo * 32, va);
                        for (int c = c_lo; c < c_hi; c += 2)
                        {
                            tmem_ld_wait();
                            tmem_ld_32x32(region + min(c + 1, 7) * 32, vb);
                            reduce(va, c);
                            if (c + 1 >= c_hi) break;
                            tmem_ld_wait();
                            tmem_ld_32x32(region + min(c + 2, 7) * 32, va);
                            reduce(vb, c + 1);
                        }
                        tmem_ld_wait(); // (the last prefetch)
                    }
                    else
                    {
                        for (int c = c_lo; c < c_hi; c++)
                        {
                            tmem_ld_32x32(region + c * 32, va);
                            tmem_ld_wait();
                            reduce(va, c);
                        }
                    }
                }
                stamp(2);
                if (two_halves)
                {
                    // the Q tile of this item is dead once S has been computed (and is not reloaded before both issuers have passed
                    // this item): its first kilobyte carries the 2 x 128 partial maxima of the tile
                    float *ex = reinterpret_cast<float *>(a16_smem + buf * ATC_BUF_BYTES + t * ATC_Q_BYTES);
                    ex[half * 128 + q * 32 + lane] = mx;
                    named_bar_sync(pair_bar, 64);
                    mx = fmaxf(mx, ex[(half ^ 1) * 128 + q * 32 + lane]);
                    fence_proxy_async_smem(); // these generic-proxy accesses precede the TMA reload of the buffer (ordered through pfull -> empty)
                }

                // the exponential pass belongs to one query tile at a time (tile 0 first; the very first wait of a fresh barrier passes)
                if (partner && p.stagger) mbar_wait(turn_bar(t, q), t == 0 ? (ph ^ 1u) : ph, p.error_flag, KERR_ATT_WG_TURN);
                stamp(7);
                // ---- pass 2: p = 2^((s - max) * scale) as bf16, over S columns this warp has already consumed ----
                const float msc = mx * sl;
                {
                    uint32_t va[32], vb[32];
                    auto expo = [&](const uint32_t *v, int c) {
                        uint32_t w[16];
                        if (c < nfull)
                        {
#pragma unroll
                            for (int j = 0; j < 16; j++)
                            {
                                const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), sl, -msc));
                                const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sl, -msc));
                                w[j] = pack_bf16x2(p0, p1);
                            }
                        }
                        else
                        {
#pragma unroll
                            for (int j = 0; j < 16; j++)
                            {
                                const float p0 = (2 * j < tail) ? ex2_approx(fmaf(__uint_as_float(v[2 * j]), sl, -msc)) : 0.0f;
                                const float p1 = (2 * j + 1 < tail) ? ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sl, -msc)) : 0.0f;
                                w[j] = pack_bf16x2(p0, p1);
                            }
                        }
                        tmem_st_32x16(region + p_col + (c - c_lo) * 16, w);
                    };
                    // P of chunk c lands on S columns of chunks <= c of this half: the prefetched chunk c + 1 is never overwritten early
                    if constexpr (DB)
                    {
                        tmem_ld_32x32(region + c_lo * 32, va);
                        for (int c = c_lo; c < c_hi; c += 2)
                        {
                            tmem_ld_wait();
                            tmem_ld_32x32(region + min(c + 1, 7) * 32, vb);
                            expo(va, c);
                            if (c + 1 >= c_hi) break;
                            tmem_ld_wait();
                            tmem_ld_32x32(region + min(c + 2, 7) * 32, va);
                            expo(vb, c + 1);
                        }
                        tmem_ld_wait(); // (the last prefetch)
                    }
                    else
                    {
                        for (int c = c_lo; c < c_hi; c++)
                        {
                            tmem_ld_32x32(region + c * 32, va);
                            tmem_ld_wait();
                            expo(va, c);
                        }
                    }
                }
                if (partner && p.stagger)
                {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(turn_bar(t ^ 1, q)); // the other tile's turn
                }
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(pfull_bar(t));
                }
                stamp(3);
                if (half == 0) continue; // the half-1 warp finishes the row block

                // ---- O = P.V divided by the row sum P.1 ----
                mbar_wait(ofull_bar(t), ph, p.error_flag, KERR_ATT_WG_OFULL);
                tcgen05_fence_after();
                stamp(4);
                uint32_t o[64], rs[4];
                tmem_ld_32x32(region + p.o_col, o);
                tmem_ld_32x32(region + p.o_col + 32, o + 32);
                tmem_ld_32xN<4>(region + p.o_col + 64, rs);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(sfree_bar(t)); // region t may be overwritten by the next item's S
                stamp(6);
                store_o_rows(o, 1.0f / __uint_as_float(rs[0]), oslab, oslab_addr, &tma_out, p.out_f32 != 0, h * ATT_HD, t * 128 + q * 32, b, lane);
                stamp(5);
            }
            if (lane == 0) tma_store_wait_all(); // the slab is read, and the rows are written, before the CTA goes away
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == ATC_W_ALLOC)
    {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// =====================================================================================================
// tcgen05 kernel for long sequences (tokens > 256, e.g. 577 tokens at 384 pixels): key-blocked variant
// =====================================================================================================
// Same building blocks as attention_tc_kernel; a work item is (image, head, PAIR of 128-row query tiles) and walks over the
// keys in n_kb blocks of kb <= 256 keys.  Per block: S = Q.K_blk^T into TMEM, the softmax warpgroup folds the block into a
// running row maximum m and row sum l (online softmax), writes bf16 P over S, O_blk = P.V_blk comes back through TMEM and is
// accumulated in REGISTERS: O = O * 2^((m_old - m_new) * scale) + O_blk.  Q tiles stay resident for the whole item
// (double-buffered across items), K / V blocks stream through a two-slot ring.

constexpr int ATL_OFF_KV = 4 * ATC_Q_BYTES;                  // after 2 item slots x 2 query tiles
constexpr int ATL_OFF_OSLAB = ATL_OFF_KV + 4 * ATC_KV_BYTES; // 2 step slots x (K + V); then one output slab per softmax warp
constexpr int ATL_OFF_BARS = ATL_OFF_OSLAB + 8 * ATC_OSLAB_BYTES;
constexpr int ATL_NUM_BARS = 16;
constexpr int ATL_OFF_TMEM_PTR = ATL_OFF_BARS + ATL_NUM_BARS * 8;
constexpr int ATL_SMEM = ATL_OFF_TMEM_PTR + 16;
static_assert(ATL_SMEM <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");

struct AttnLongParams
{
    void *out; // [batch * tokens][heads * 64], bf16 or (out_f32) fp32
    int out_f32;
    int batch, tokens, heads;
    int kb;       // keys per block: multiple of 16, <= 256
    int n_kb;     // key blocks per item
    int n_qtiles; // 128-row query tiles per (image, head)
    int n_qpairs; // work items per (image, head)
    int *error_flag;
};

__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_long_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                         const __grid_constant__ CUtensorMap tma_out, const AttnLongParams p)
{
    extern __shared__ __align__(1024) uint8_t atl_smem[];
    const uint32_t base = smem_u32(atl_smem);
    if ((base & 1023u) != 0)
    {
        if (threadIdx.x == 0 && p.error_flag) atomicExch(p.error_flag, KERR_SMEM_ALIGN);
        return;
    }
    const uint32_t bars = base + ATL_OFF_BARS;
    auto qfull_bar = [&](int b) { return bars + 8u * b; };
    auto qempty_bar = [&](int b) { return bars + 8u * (2 + b); };
    auto kvfull_bar = [&](int b) { return bars + 8u * (4 + b); };
    auto kvempty_bar = [&](int b) { return bars + 8u * (6 + b); };
    auto sfull_bar = [&](int t) { return bars + 8u * (8 + t); };
    auto pfull_bar = [&](int t) { return bars + 8u * (10 + t); };
    auto ofull_bar = [&](int t) { return bars + 8u * (12 + t); };
    auto sfree_bar = [&](int t) { return bars + 8u * (14 + t); };
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(atl_smem + ATL_OFF_TMEM_PTR);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31; // (uniform: the MMA issuers' descriptors stay in uniform registers, see ptx.cuh)
    const int items = p.batch * p.heads * p.n_qpairs;
    const int D = p.heads * ATT_HD;
    const int n_it = blockIdx.x < items ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0; // items of this CTA
    // query tiles of this CTA's ii-th item (2, or 1 for the last pair of an odd tile count)
    auto tiles_of = [&](int ii) { return min(2, p.n_qtiles - 2 * ((blockIdx.x + ii * (int)gridDim.x) % p.n_qpairs)); };

    if (warp == ATC_W_PRODUCER && lane == 0)
    {
        tma_prefetch_desc(&tma_q);
        tma_prefetch_desc(&tma_kv);
    }
    if (warp == ATC_W_MMA && lane == 0)
    {
        for (int b = 0; b < 2; b++)
        {
            mbar_init(qfull_bar(b), 1);
            mbar_init(qempty_bar(b), 1);
            mbar_init(kvfull_bar(b), 1);
            mbar_init(kvempty_bar(b), 1);
            mbar_init(sfull_bar(b), 1);
            mbar_init(pfull_bar(b), 4); // one arrive per softmax warp
            mbar_init(ofull_bar(b), 1);
            mbar_init(sfree_bar(b), 4);
        }
        fence_barrier_init();
    }
    if (warp == ATC_W_ALLOC)
    {
        tmem_alloc(base + ATL_OFF_TMEM_PTR, 512);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    griddep_launch_dependents();
    griddep_wait();

    if (warp < 4)
    {
        // the softmax warps keep 64 output accumulators per thread on top of a score chunk: they take the registers
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == ATC_W_PRODUCER)
        {
            if (lane == 0)
            {
                const uint32_t tx_kv = (uint32_t)(2 * p.kb * 128);
                int st = 0;
                for (int ii = 0; ii < n_it; ii++)
                {
                    const int item = blockIdx.x + ii * gridDim.x;
                    const int bh = item / p.n_qpairs, qp = item - bh * p.n_qpairs;
                    const int b = bh / p.heads, h = bh - b * p.heads;
                    const int row = b * p.tokens, nt = tiles_of(ii), qb = ii & 1;
                    mbar_wait(qempty_bar(qb), ((uint32_t)(ii >> 1) & 1u) ^ 1u, p.error_flag, KERR_ATT_PRODUCER);
                    mbar_arrive_expect_tx(qfull_bar(qb), (uint32_t)(nt * ATC_Q_BYTES));
                    for (int t = 0; t < nt; t++)
                        tma_load_2d(base + (qb * 2 + t) * ATC_Q_BYTES, &tma_q, qfull_bar(qb), h * ATT_HD, row + (2 * qp + t) * 128);
                    for (int kbi = 0; kbi < p.n_kb; kbi++, st++)
                    {
                        const int sb = st & 1;
                        mbar_wait(kvempty_bar(sb), ((uint32_t)(st >> 1) & 1u) ^ 1u, p.error_flag, KERR_ATT_PRODUCER);
                        mbar_arrive_expect_tx(kvfull_bar(sb), tx_kv);
                        const uint32_t dst = base + ATL_OFF_KV + sb * 2 * ATC_KV_BYTES;
                        tma_load_2d(dst, &tma_kv, kvfull_bar(sb), D + h * ATT_HD, row + kbi * p.kb);
                        tma_load_2d(dst + ATC_KV_BYTES, &tma_kv, kvfull_bar(sb), 2 * D + h * ATT_HD, row + kbi * p.kb);
                    }
                }
            }
        }
        else if (warp == ATC_W_MMA)
        {
            // one thread serves both query tiles and polls, per tile, "S of the next key block may start" / "P is ready"
            if (lane == 0)
            {
                const uint32_t idesc_s = umma_idesc(1, 1, 128, (uint32_t)p.kb);
                const uint32_t idesc_o = umma_idesc(1, 1, 128, ATT_HD) | UMMA_IDESC_B_MN_MAJOR;
                const int ksteps = p.kb / 16, n_steps = n_it * p.n_kb;
                int s_step[2] = {0, 0}, pv_step[2] = {0, 0}; // next step (item-major, key block minor) per tile
                uint32_t cnt_s[2] = {0, 0}, cnt_pv[2] = {0, 0}; // steps the tile really took part in (barrier phases)
                long long t0 = clock64();
                while (pv_step[0] < n_steps || pv_step[1] < n_steps)
                {
                    bool progress = false;
#pragma unroll
                    for (int t = 0; t < 2; t++)
                    {
                        if (s_step[t] < n_steps && pv_step[t] == s_step[t])
                        {
                            const int step = s_step[t], ii = step / p.n_kb, kbi = step - ii * p.n_kb, sb = step & 1, qb = ii & 1;
                            if (t == 1 && tiles_of(ii) < 2)
                            {
                                // tile 1 does not exist in this item: pass, and release what only waited for it
                                s_step[1]++, pv_step[1]++;
                                if (kbi == p.n_kb - 1 && s_step[0] > step) tcgen05_commit(qempty_bar(qb));
                                if (pv_step[0] > step) tcgen05_commit(kvempty_bar(sb));
                                progress = true;
                            }
                            else if (mbar_test_wait(kvfull_bar(sb), (uint32_t)(step >> 1) & 1u) && mbar_test_wait(qfull_bar(qb), (uint32_t)(ii >> 1) & 1u) &&
                                     mbar_test_wait(sfree_bar(t), (cnt_s[t] & 1u) ^ 1u))
                            {
                                tcgen05_fence_after();
                                const uint64_t q_desc = umma_smem_desc_sw128(base + (qb * 2 + t) * ATC_Q_BYTES);
                                const uint64_t k_desc = umma_smem_desc_sw128(base + ATL_OFF_KV + sb * 2 * ATC_KV_BYTES);
#pragma unroll
                                for (int k = 0; k < 4; k++)
                                    umma_ss<KIND_BF16>(tmem_base + t * ATC_REGION_COLS, q_desc + 2u * k, k_desc + 2u * k, idesc_s, k != 0 ? 1u : 0u);
                                tcgen05_commit(sfull_bar(t));
                                s_step[t]++, cnt_s[t]++;
                                if (kbi == p.n_kb - 1 && s_step[t ^ 1] > step) tcgen05_commit(qempty_bar(qb)); // last use of the Q tiles
                                progress = true;
                            }
                        }
                        if (pv_step[t] < s_step[t])
                        {
                            const int step = pv_step[t], sb = step & 1;
                            if (mbar_test_wait(pfull_bar(t), cnt_pv[t] & 1u))
                            {
                                tcgen05_fence_after();
                                const uint64_t v_desc = umma_smem_desc_sw128(base + ATL_OFF_KV + sb * 2 * ATC_KV_BYTES + ATC_KV_BYTES);
                                const uint32_t region = tmem_base + t * ATC_REGION_COLS;
                                for (int k = 0; k < ksteps; k++)
                                    umma_ts_bf16(region + ATC_O_COL, region + 8u * k, v_desc + (uint64_t)(128u * k), idesc_o, k != 0 ? 1u : 0u);
                                tcgen05_commit(ofull_bar(t));
                                pv_step[t]++, cnt_pv[t]++;
                                if (pv_step[t ^ 1] > step) tcgen05_commit(kvempty_bar(sb)); // both tiles are done with this K / V block
                                progress = true;
                            }
                        }
                    }
                    if (progress)
                        t0 = clock64();
                    else if (clock64() - t0 > 8000000000LL)
                    {
                        if (p.error_flag) atomicExch(p.error_flag, KERR_ATT_MMA_FULL);
                        __threadfence_system();
                        __trap();
                    }
                }
            }
        }
    }
    else
    {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ===================== softmax + output warpgroups (one per query tile of the pair) =====================
        const int t = (warp - 4) >> 2;
        const int q = warp & 3;
        const uint32_t region = tmem_base + ((uint32_t)(q * 32) << 16) + t * ATC_REGION_COLS;
        const float sl = 0.125f * 1.4426950408889634f; // 1/sqrt(64) * log2(e)
        const int kchunks = (p.kb + 31) >> 5;          // 32-key chunks of a block (P is written for all of them)
        uint32_t cnt = 0;                              // steps this tile took part in (barrier phase)
        for (int ii = 0; ii < n_it; ii++)
        {
            if (t >= tiles_of(ii)) continue;
            const int item = blockIdx.x + ii * gridDim.x;
            const int bh = item / p.n_qpairs, qp = item - bh * p.n_qpairs;
            const int b = bh / p.heads, h = bh - b * p.heads;
            const int qrow = (2 * qp + t) * 128 + q * 32 + lane; // query row within the image
            float o[64];
#pragma unroll
            for (int j = 0; j < 64; j++) o[j] = 0.0f;
            float m_run = -INFINITY, l_run = 0.0f;
            for (int kbi = 0; kbi < p.n_kb; kbi++, cnt++)
            {
                const uint32_t ph = cnt & 1u;
                const int valid_keys = min(p.kb, p.tokens - kbi * p.kb); // >= 1 by construction of n_kb
                const int nfull = valid_keys >> 5, tail = valid_keys & 31;
                const int nchunks = nfull + (tail ? 1 : 0);
                mbar_wait(sfull_bar(t), ph, p.error_flag, KERR_ATT_WG_SFULL);
                tcgen05_fence_after();
                // pass 1: block maximum over the valid keys
                float bmax = -INFINITY;
                for (int c = 0; c < nchunks; c++)
                {
                    uint32_t v[32];
                    tmem_ld_32x32(region + c * 32, v);
                    tmem_ld_wait();
                    const int lim = c < nfull ? 32 : tail;
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if (j < lim) bmax = fmaxf(bmax, __uint_as_float(v[j]));
                }
                const float m_new = fmaxf(m_run, bmax);
                const float alpha = ex2_approx((m_run - m_new) * sl); // 0 for the first block (m_run = -inf)
                const float msc = m_new * sl;
                m_run = m_new;
                // pass 2: p = 2^((s - m) * scale), bf16 P over the S columns; fully masked chunks are written as zeros
                float sum0 = 0.0f, sum1 = 0.0f;
                for (int c = 0; c < kchunks; c++)
                {
                    uint32_t w[16];
                    if (c < nchunks)
                    {
                        uint32_t v[32];
                        tmem_ld_32x32(region + c * 32, v);
                        tmem_ld_wait();
                        const int lim = c < nfull ? 32 : tail;
#pragma unroll
                        for (int j = 0; j < 16; j++)
                        {
                            const float p0 = (2 * j < lim) ? ex2_approx(fmaf(__uint_as_float(v[2 * j]), sl, -msc)) : 0.0f;
                            const float p1 = (2 * j + 1 < lim) ? ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sl, -msc)) : 0.0f;
                            sum0 += p0, sum1 += p1;
                            w[j] = pack_bf16x2(p0, p1);
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int j = 0; j < 16; j++) w[j] = 0u;
                    }
                    tmem_st_32x16(region + c * 16, w);
                }
                l_run = fmaf(l_run, alpha, sum0 + sum1);
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(pfull_bar(t));

                // O = O * alpha + P.V_blk
                mbar_wait(ofull_bar(t), ph, p.error_flag, KERR_ATT_WG_OFULL);
                tcgen05_fence_after();
                uint32_t ob[64];
                tmem_ld_32x32(region + ATC_O_COL, ob);
                tmem_ld_32x32(region + ATC_O_COL + 32, ob + 32);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(sfree_bar(t)); // region t may be overwritten by the next block's S
#pragma unroll
                for (int j = 0; j < 64; j++) o[j] = fmaf(o[j], alpha, __uint_as_float(ob[j]));
            }
            // O / rowsum through this warp's swizzled slab and 3-D TMA stores, as in attention_tc_kernel
            if (qrow - lane < p.tokens) // (warp-uniform: the box starts inside the image)
            {
                uint8_t *oslab = atl_smem + ATL_OFF_OSLAB + (warp - 4) * ATC_OSLAB_BYTES;
                store_o_rows(reinterpret_cast<const uint32_t *>(o), 1.0f / l_run, oslab, base + ATL_OFF_OSLAB + (warp - 4) * ATC_OSLAB_BYTES, &tma_out,
                             p.out_f32 != 0, h * ATT_HD, qrow - lane, b, lane);
            }
        }
        if (lane == 0) tma_store_wait_all(); // the slab is read, and the rows are written, before the CTA goes away
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == ATC_W_ALLOC)
    {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static cudaError_t launch_attention_tc_long(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t stream, int *error_flag,
                                            int num_sms, bool out_f32)
{
    cudaError_t e = cudaFuncSetAttribute(attention_tc_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATL_SMEM);
    if (e != cudaSuccess) return e;
    AttnLongParams p;
    p.out = out, p.out_f32 = out_f32 ? 1 : 0;
    p.batch = batch, p.tokens = tokens, p.heads = heads;
    p.n_kb = (tokens + 255) / 256;                            // fewest blocks of at most 256 keys ...
    p.kb = (((tokens + p.n_kb - 1) / p.n_kb) + 15) & ~15;     // ... of equal size, rounded up to the UMMA granularity
    p.n_kb = (tokens + p.kb - 1) / p.kb;                      // (every block then holds at least one valid key)
    p.n_qtiles = (tokens + 127) / 128;
    p.n_qpairs = (p.n_qtiles + 1) / 2;
    p.error_flag = error_flag;
    const long long D = (long long)heads * ATT_HD, rows = (long long)batch * tokens;
    CUtensorMap map_q, map_kv;
    e = encode_tma_2d(&map_q, 2, qkv, 3 * D, rows, 3 * D * 2, ATT_HD, 128, true);
    if (e != cudaSuccess) return e;
    e = encode_tma_2d(&map_kv, 2, qkv, 3 * D, rows, 3 * D * 2, ATT_HD, p.kb, true);
    if (e != cudaSuccess) return e;
    const long long items = (long long)batch * heads * p.n_qpairs;
    const int sms = num_sms > 0 ? num_sms : 148;
    CUtensorMap map_out; // [batch][tokens][D]: a 32-row store box never spills into the next image
    const int oe = out_f32 ? 4 : 2;
    e = encode_tma_3d(&map_out, oe, out, D, tokens, batch, D * oe, (long long)tokens * D * oe, out_f32 ? 32 : ATT_HD, 32, true);
    if (e != cudaSuccess) return e;
    return launch_pdl(attention_tc_long_kernel, dim3((unsigned)(items < sms ? items : sms)), dim3(ATC_THREADS), (size_t)ATL_SMEM, stream, 1, map_q,
                      map_kv, map_out, p);
}

template <int MODE>
static cudaError_t launch_attention_tc_v(const CUtensorMap &map_q, const CUtensorMap &map_kv, const CUtensorMap &map_out, const AttnTcParams &p,
                                         int grid, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM);
    if (e != cudaSuccess) return e;
    return launch_pdl(attention_tc_kernel<MODE>, dim3(grid), dim3(ATC_THREADS), (size_t)ATC_SMEM, stream, 1, map_q, map_kv, map_out, p);
}

// Kernel variant: 0..2 = MODE of attention_tc_kernel (12 warps); 4 / 14 / 24 / 34 (+ 100) = attention_tc16_kernel with its A/B switches.
// The default is the measured best; NETCUDA_ATT_TC_VARIANT (read once) selects another one for A/B runs.
constexpr int ATT_TC_DEFAULT_VARIANT = 4; // the 16-softmax-warp kernel (177 us per 512 x 12 x 197 launch in isolation; round-1 kernel = variant 0: 200 us)
static int att_tc_variant()
{
    static const int v = getenv("NETCUDA_ATT_TC_VARIANT") ? atoi(getenv("NETCUDA_ATT_TC_VARIANT")) : ATT_TC_DEFAULT_VARIANT;
    return v;
}

static cudaError_t launch_attention_tc(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t stream, int *error_flag,
                                       int num_sms, bool out_f32, int tc_variant)
{
    AttnTcParams p;
    p.out = out, p.out_f32 = out_f32 ? 1 : 0;
    p.batch = batch, p.tokens = tokens, p.heads = heads;
    p.n_pad = (tokens + 15) & ~15;
    p.n_mtiles = tokens > 128 ? 2 : 1;
    p.stagger = 1;
    p.error_flag = error_flag;
    p.debug = nullptr;
#ifdef NETCUDA_DEBUG_TIMELINE
    if (const char *dbg = getenv("NETCUDA_ATTENTION_DEBUG_PTR")) p.debug = reinterpret_cast<long long *>(strtoull(dbg, nullptr, 0));
#endif
    const long long D = (long long)heads * ATT_HD, rows = (long long)batch * tokens;
    // rows past the last token of the last image are zero-filled by TMA; a tile that runs into the next image reads
    // that image's (finite) rows: extra keys are masked in the softmax, extra query rows are never stored
    CUtensorMap map_q, map_kv;
    cudaError_t e = encode_tma_2d(&map_q, 2, qkv, 3 * D, rows, 3 * D * 2, ATT_HD, 128, true);
    if (e != cudaSuccess) return e;
    e = encode_tma_2d(&map_kv, 2, qkv, 3 * D, rows, 3 * D * 2, ATT_HD, p.n_pad, true);
    if (e != cudaSuccess) return e;
    // output [batch][tokens][D] as a 3-D map: a 32-row store box never spills into the next image
    CUtensorMap map_out;
    const int oe = out_f32 ? 4 : 2;
    e = encode_tma_3d(&map_out, oe, out, D, tokens, batch, D * oe, (long long)tokens * D * oe, out_f32 ? 32 : ATT_HD, 32, true);
    if (e != cudaSuccess) return e;
    const int items = batch * heads;
    const int sms = num_sms > 0 ? num_sms : 148;
    const int grid = items < sms ? items : sms;
    const int vcode = tc_variant >= 0 ? tc_variant : att_tc_variant();
    const int variant = vcode % 100;
    if (variant % 10 == 4)
    {
        // 4: single-buffered loads, keys split 128 | rest;  14: double-buffered;  24 / 34: the same with the balanced split (96 | rest,
        // up to 224 keys);  + 100: no turn-taking between the two query tiles
        const int nchunks = (tokens + 31) / 32;
        const bool balanced = variant >= 20 && nchunks <= 7;
        p.split = balanced ? 3 : 4;
        p.p1_col = balanced ? 112 : 144;
        p.o_col = balanced ? 176 : 64;
        p.stagger = vcode < 100 ? 1 : 0;
        const bool db = (variant / 10) & 1;
        auto kern = db ? attention_tc16_kernel<true> : attention_tc16_kernel<false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A16_SMEM);
        if (e != cudaSuccess) return e;
        return launch_pdl(kern, dim3(grid), dim3(A16_THREADS), (size_t)A16_SMEM, stream, 1, map_q, map_kv, map_out, p);
    }
    switch (variant)
    {
    case 0: return launch_attention_tc_v<0>(map_q, map_kv, map_out, p, grid, stream);
    case 1: return launch_attention_tc_v<1>(map_q, map_kv, map_out, p, grid, stream);
    case 2: return launch_attention_tc_v<2>(map_q, map_kv, map_out, p, grid, stream);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_attention(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t stream, int *error_flag, int num_sms,
                             int variant, bool out_f32, int tc_variant)
{
    if (batch <= 0) return cudaSuccess;
    if (tokens <= 0 || heads <= 0 || heads > 65535 || batch > 65535) return cudaErrorInvalidValue;
    if (tokens <= 256 && variant == 0) return launch_attention_tc(qkv, out, batch, tokens, heads, stream, error_flag, num_sms, out_f32, tc_variant);
    if (variant == 0) return launch_attention_tc_long(qkv, out, batch, tokens, heads, stream, error_flag, num_sms, out_f32);
    const int tpad = (tokens + 15) & ~15;
    const size_t smem = (size_t)tpad * 128 * 2;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    // 197 tokens: 7 warps x 2 CTAs = 224 rows (12 % padding); general case: 4 warps per CTA.
    const __nv_bfloat16 *q = reinterpret_cast<const __nv_bfloat16 *>(qkv);
    if (tokens > 112 && tokens <= 224)
        return out_f32 ? launch_attention_nw<7, float>(q, reinterpret_cast<float *>(out), batch, tokens, heads, tpad, smem, stream)
                       : launch_attention_nw<7, __nv_bfloat16>(q, reinterpret_cast<__nv_bfloat16 *>(out), batch, tokens, heads, tpad, smem, stream);
    return out_f32 ? launch_attention_nw<4, float>(q, reinterpret_cast<float *>(out), batch, tokens, heads, tpad, smem, stream)
                   : launch_attention_nw<4, __nv_bfloat16>(q, reinterpret_cast<__nv_bfloat16 *>(out), batch, tokens, heads, tpad, smem, stream);
}

} // namespace nc

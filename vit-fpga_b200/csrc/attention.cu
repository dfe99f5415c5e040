// attention.cu -- multi-head attention core, softmax(q k^T / sqrt(64)) v, per (image, head).
//
// No reference counterpart (the reference has no ViT); semantics follow
// oracle_attention (oracle/oracle_vit.c).  Input is the packed bf16 output of the QKV projection,
// rows [q | k | v] of 3*heads*64 columns per token, so no re-layout kernel runs between the GEMM and
// attention, and the output [token][heads*64] is directly the A operand of the output projection.
//
// Round-1 kernel: flash-style single pass with mma.sync m16n8k16 (bf16 in, fp32 accumulate).
//   - one CTA per (query block, head, image); K and V of the head live in shared memory once
//     (197 keys: 2 x 26 KB; 577 keys: 2 x 74 KB), XOR-swizzled in 16-byte chunks so ldmatrix is
//     bank-conflict free; rows >= tokens are zero-filled and masked to -inf before the softmax;
//   - each warp owns 16 query rows; S and P never leave registers (online softmax over 64-key
//     chunks, exp2 with the 1/sqrt(64)*log2(e) scale folded in, warp-shuffle row reductions);
//   - the attention core is ~4 % of ViT-B's FLOPs; moving it to tcgen05 is listed in DESIGN.md.
#include "kernels.h"
#include "ptx.cuh"

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

template <int NW>
__global__ void __launch_bounds__(NW * 32)
attention_kernel(const __nv_bfloat16 *__restrict__ qkv, __nv_bfloat16 *__restrict__ out, int tokens, int heads, int tpad)
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
    __nv_bfloat16 *obase = out + b * tokens * (long long)D + h * ATT_HD + tg * 2;
    const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        if (r0 < tokens) *reinterpret_cast<uint32_t *>(obase + (long long)r0 * D + i * 8) = pack_bf16x2(o[i][0] * i0, o[i][1] * i0);
        if (r1 < tokens) *reinterpret_cast<uint32_t *>(obase + (long long)r1 * D + i * 8) = pack_bf16x2(o[i][2] * i1, o[i][3] * i1);
    }
}

template <int NW>
static cudaError_t launch_attention_nw(const __nv_bfloat16 *q, __nv_bfloat16 *o, int batch, int tokens, int heads, int tpad,
                                       size_t smem, cudaStream_t stream)
{
    // per-device attribute; setting it on every launch costs ~1 us and keeps multi-GPU processes correct
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    dim3 grid((tokens + 16 * NW - 1) / (16 * NW), heads, batch);
    attention_kernel<NW><<<grid, NW * 32, smem, stream>>>(q, o, tokens, heads, tpad);
    return cudaGetLastError();
}

cudaError_t launch_attention(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t stream)
{
    if (batch <= 0) return cudaSuccess;
    if (tokens <= 0 || heads <= 0 || heads > 65535 || batch > 65535) return cudaErrorInvalidValue;
    const int tpad = (tokens + 15) & ~15;
    const size_t smem = (size_t)tpad * 128 * 2;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    // 197 tokens: 7 warps x 2 CTAs = 224 rows (12 % padding); general case: 4 warps per CTA.
    const __nv_bfloat16 *q = reinterpret_cast<const __nv_bfloat16 *>(qkv);
    __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(out);
    if (tokens > 112 && tokens <= 224)
        return launch_attention_nw<7>(q, o, batch, tokens, heads, tpad, smem, stream);
    return launch_attention_nw<4>(q, o, batch, tokens, heads, tpad, smem, stream);
}

} // namespace nc

// elementwise.cu -- the HBM-bound kernels of the forward path: LayerNorm, patch extraction, class
// token rows, and the fp32 <-> operand-type conversions at the API boundary
// (replacing the host scalar copy loops of src/netFPGA.cpp:266-267,285-289 for batched input).
// All of them are one pass over their data with 128-bit accesses where alignment allows.
#include "kernels.h"

#include <algorithm>
#include <cstdlib>
#include "ptx.cuh"

namespace nc
{

// ---- LayerNorm: one warp per row, row cached in registers, two-pass (mean, then variance) --------

template <int MAXV, typename OutT> // MAXV float4 per lane -> dim <= 128 * MAXV; OutT = bf16 (bf16 nets) or float (tf32 nets)
__global__ void __launch_bounds__(256)
layernorm_kernel(const float *__restrict__ x, long long ldx, const float *__restrict__ gamma, const float *__restrict__ beta,
                 OutT *__restrict__ y, long long ldy, int rows, int dim, float eps)
{
    griddep_launch_dependents();
    griddep_wait(); // x is the previous kernel's output, y the previous-but-one's input
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int nv = dim >> 2; // float4 per row
    const float4 *xr = reinterpret_cast<const float4 *>(x + (long long)warp * ldx);
    float4 v[MAXV];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; i++)
    {
        const int idx = i * 32 + lane;
        if (idx < nv)
        {
            v[i] = xr[idx];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)dim;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; i++)
    {
        const int idx = i * 32 + lane;
        if (idx < nv)
        {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)dim + eps);
    const float4 *g4 = reinterpret_cast<const float4 *>(gamma);
    const float4 *b4 = reinterpret_cast<const float4 *>(beta);
    OutT *yrow = y + (long long)warp * ldy;
#pragma unroll
    for (int i = 0; i < MAXV; i++)
    {
        const int idx = i * 32 + lane;
        if (idx < nv)
        {
            const float4 g = g4[idx], b = b4[idx];
            const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
            const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
            if constexpr (sizeof(OutT) == 2)
                reinterpret_cast<uint2 *>(yrow)[idx] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
            else
                reinterpret_cast<float4 *>(yrow)[idx] = make_float4(o0, o1, o2, o3);
        }
    }
}

// Rows of at most 192 elements (ViT-Tiny: D = 192 = 48 float4): one HALF warp per row, three float4 per lane.  With a whole warp per
// row a third of the lanes' second load is empty and a warp has 768 bytes in flight; here every lane carries three loads (1536 bytes per
// warp) and there are half as many warps to schedule.  Same two-pass arithmetic; the reductions stay inside the half warp (xor 8..1).
template <typename OutT>
__global__ void __launch_bounds__(256)
layernorm_halfwarp_kernel(const float *__restrict__ x, long long ldx, const float *__restrict__ gamma, const float *__restrict__ beta,
                          OutT *__restrict__ y, long long ldy, int rows, int dim, float eps)
{
    griddep_launch_dependents();
    griddep_wait();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int l = threadIdx.x & 15;
    const bool valid = row < rows; // (the other half of the warp may still hold a row: no early return before the shuffles)
    const int nv = dim >> 2;
    const float4 *xr = reinterpret_cast<const float4 *>(x + (long long)(valid ? row : 0) * ldx);
    float4 v[3];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        const int idx = i * 16 + l;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && idx < nv)
        {
            v[i] = xr[idx];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)dim;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        const int idx = i * 16 + l;
        if (valid && idx < nv)
        {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (!valid) return;
    const float rstd = rsqrtf(q / (float)dim + eps);
    const float4 *g4 = reinterpret_cast<const float4 *>(gamma);
    const float4 *b4 = reinterpret_cast<const float4 *>(beta);
    OutT *yrow = y + (long long)row * ldy;
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        const int idx = i * 16 + l;
        if (idx < nv)
        {
            const float4 g = g4[idx], b = b4[idx];
            const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
            const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
            if constexpr (sizeof(OutT) == 2)
                reinterpret_cast<uint2 *>(yrow)[idx] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
            else
                reinterpret_cast<float4 *>(yrow)[idx] = make_float4(o0, o1, o2, o3);
        }
    }
}

template <typename OutT>
static cudaError_t launch_layernorm_t(const float *x, long long ldx, const float *gamma, const float *beta, OutT *y, long long ldy, int rows,
                                      int dim, float eps, cudaStream_t stream)
{
    const int threads = 256, rows_per_block = threads / 32;
    const int grid = (rows + rows_per_block - 1) / rows_per_block;
    static const bool halfwarp = !getenv("NETCUDA_LN_HALFWARP") || atoi(getenv("NETCUDA_LN_HALFWARP")) != 0; // (=0: a warp per row, A/B)
    if (dim <= 192 && halfwarp)
        return launch_pdl(layernorm_halfwarp_kernel<OutT>, dim3((rows + 15) / 16), dim3(threads), 0, stream, 1, x, ldx, gamma, beta, y, ldy, rows, dim, eps);
    if (dim <= 256)
        return launch_pdl(layernorm_kernel<2, OutT>, dim3(grid), dim3(threads), 0, stream, 1, x, ldx, gamma, beta, y, ldy, rows, dim, eps);
    else if (dim <= 1024)
        return launch_pdl(layernorm_kernel<8, OutT>, dim3(grid), dim3(threads), 0, stream, 1, x, ldx, gamma, beta, y, ldy, rows, dim, eps);
    return launch_pdl(layernorm_kernel<32, OutT>, dim3(grid), dim3(threads), 0, stream, 1, x, ldx, gamma, beta, y, ldy, rows, dim, eps);
}

cudaError_t launch_layernorm(const float *x, long long ldx, const float *gamma, const float *beta, void *y, long long ldy,
                             int rows, int dim, float eps, cudaStream_t stream, bool out_f32)
{
    if (rows <= 0) return cudaSuccess;
    if (dim <= 0 || (dim & 3) || (ldx & 3) || (ldy & 3) || dim > 4096) return cudaErrorInvalidValue;
    if (out_f32) return launch_layernorm_t(x, ldx, gamma, beta, reinterpret_cast<float *>(y), ldy, rows, dim, eps, stream);
    return launch_layernorm_t(x, ldx, gamma, beta, reinterpret_cast<__nv_bfloat16 *>(y), ldy, rows, dim, eps, stream);
}

// ---- patch extraction: fp32 NCHW image -> bf16 patch rows (the im2col of a stride==kernel conv) ----
// One thread converts 8 horizontally adjacent pixels (two float4 loads -> one 16-byte store).
// Threads walk the image in its own memory order, so loads are fully coalesced; the two threads that
// cover one 16-pixel patch row write adjacent 16-byte halves of the same 32-byte sector.

// 8 values of one patch row, as bf16 (one 16-byte store) or unchanged as fp32 (tf32 nets: two 16-byte stores)
template <typename OutT>
__device__ __forceinline__ void store_patch8(OutT *dst, const float *v)
{
    if constexpr (sizeof(OutT) == 2)
        *reinterpret_cast<uint4 *>(dst) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    else
    {
        reinterpret_cast<float4 *>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4 *>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

template <typename OutT>
__global__ void __launch_bounds__(256)
patchify_kernel(const float *__restrict__ img, OutT *__restrict__ patches, long long total8, int S, int P, int g)
{
    griddep_launch_dependents();
    griddep_wait();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total8) return;
    const int s8 = S >> 3;
    const int x8 = (int)(t % s8);
    long long r = t / s8;
    const int yy = (int)(r % S);
    r /= S;
    const int ch = (int)(r % 3);
    const long long b = r / 3;
    const int x = x8 << 3;
    const int gy = yy / P, py = yy - gy * P, gx = x / P, px = x - gx * P;
    const float4 *src = reinterpret_cast<const float4 *>(img + (((b * 3 + ch) * S + yy) * (long long)S + x));
    const float4 a = src[0], c = src[1];
    const float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    const long long prow = (b * g + gy) * g + gx;
    const int pcol = (ch * P + py) * P + px;
    store_patch8(patches + prow * (long long)(3 * P * P) + pcol, v);
}

cudaError_t launch_patchify(const float *img, void *patches, int batch, int image_size, int patch_size, cudaStream_t stream, bool out_f32)
{
    if (batch <= 0) return cudaSuccess;
    if ((patch_size & 7) || (image_size % patch_size)) return cudaErrorInvalidValue;
    const long long total8 = (long long)batch * 3 * image_size * (image_size >> 3);
    const int threads = 256;
    const long long grid = (total8 + threads - 1) / threads;
    if (out_f32)
        return launch_pdl(patchify_kernel<float>, dim3((unsigned)grid), dim3(threads), 0, stream, 1, img, reinterpret_cast<float *>(patches), total8,
                          image_size, patch_size, image_size / patch_size);
    return launch_pdl(patchify_kernel<__nv_bfloat16>, dim3((unsigned)grid), dim3(threads), 0, stream, 1, img,
                      reinterpret_cast<__nv_bfloat16 *>(patches), total8, image_size, patch_size, image_size / patch_size);
}

// ---- camera frames: u8 HWC -> normalised bf16 patch rows --------------------------------------------------
// The reference's image side channel carries frames as unsigned char vectors (net::image_set::resized_image_data,
// def/defines.h:31-38; staged byte by byte at src/netFPGA.cpp:314-315).  Feeding the ViT from such frames directly moves
// a quarter of the bytes over PCIe and skips the fp32 image in HBM: one thread takes 8 pixels (24 interleaved bytes) and
// writes 8 bf16 of each colour plane of the patch row.  v = (u8 * (1/255) - mean[c]) * inv_std[c], every step rounded to
// fp32 on its own (no FMA contraction) so that a host can reproduce the patch matrix bit for bit.
template <typename OutT>
__global__ void __launch_bounds__(256)
patchify_u8_kernel(const uint8_t *__restrict__ img, OutT *__restrict__ patches, long long total8, int S, int P, int g, float3 mean,
                   float3 inv_std)
{
    griddep_launch_dependents();
    griddep_wait();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total8) return;
    const int s8 = S >> 3;
    const int x = (int)(t % s8) << 3;
    const long long r = t / s8;
    const int yy = (int)(r % S);
    const long long b = r / S;
    const uint2 *src = reinterpret_cast<const uint2 *>(img + ((b * S + yy) * (long long)S + x) * 3); // 24-byte pixel groups: 8-byte aligned
    const uint2 q0 = src[0], q1 = src[1], q2 = src[2];
    const uint32_t w[6] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y};
    const int gy = yy / P, py = yy - gy * P, gx = x / P, px = x - gx * P;
    const long long prow = (b * g + gy) * g + gx;
    const float mu[3] = {mean.x, mean.y, mean.z}, is[3] = {inv_std.x, inv_std.y, inv_std.z};
#pragma unroll
    for (int ch = 0; ch < 3; ch++)
    {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            const int byte = i * 3 + ch;
            const float u = (float)((w[byte >> 2] >> (8 * (byte & 3))) & 0xFFu);
            v[i] = __fmul_rn(__fsub_rn(__fmul_rn(u, 1.0f / 255.0f), mu[ch]), is[ch]);
        }
        const int pcol = (ch * P + py) * P + px;
        store_patch8(patches + prow * (long long)(3 * P * P) + pcol, v);
    }
}

cudaError_t launch_patchify_u8(const uint8_t *img, void *patches, int batch, int image_size, int patch_size, const float *mean,
                               const float *inv_std, cudaStream_t stream, bool out_f32)
{
    if (batch <= 0) return cudaSuccess;
    if ((patch_size & 7) || (image_size % patch_size)) return cudaErrorInvalidValue;
    const long long total8 = (long long)batch * image_size * (image_size >> 3);
    const int threads = 256;
    const long long grid = (total8 + threads - 1) / threads;
    const float3 mu = make_float3(mean[0], mean[1], mean[2]), is = make_float3(inv_std[0], inv_std[1], inv_std[2]);
    if (out_f32)
        return launch_pdl(patchify_u8_kernel<float>, dim3((unsigned)grid), dim3(threads), 0, stream, 1, img, reinterpret_cast<float *>(patches),
                          total8, image_size, patch_size, image_size / patch_size, mu, is);
    return launch_pdl(patchify_u8_kernel<__nv_bfloat16>, dim3((unsigned)grid), dim3(threads), 0, stream, 1, img,
                      reinterpret_cast<__nv_bfloat16 *>(patches), total8, image_size, patch_size, image_size / patch_size, mu, is);
}

// ---- class-token rows of the residual stream ---------------------------------------------------------

__global__ void cls_rows_kernel(float *__restrict__ x, const float *__restrict__ cls, const float *__restrict__ pos, int batch,
                                int tokens, int dim)
{
    griddep_launch_dependents();
    griddep_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)batch * dim) return;
    const int b = (int)(i / dim), c = (int)(i % dim);
    x[(long long)b * tokens * dim + c] = cls[c] + pos[c];
}

cudaError_t launch_cls_rows(float *x, const float *cls, const float *pos, int batch, int tokens, int dim, cudaStream_t stream)
{
    if (batch <= 0) return cudaSuccess;
    const long long n = (long long)batch * dim;
    return launch_pdl(cls_rows_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, 1, x, cls, pos, batch, tokens, dim);
}

// ---- API-boundary conversions -------------------------------------------------------------------------

template <typename OutT, int MODE> // MODE 0: bf16, 1: fp32 copy, 2: Q1.7 quantise, 3: int8 copy
__global__ void __launch_bounds__(256)
convert_rows_kernel(const void *__restrict__ in_, OutT *__restrict__ out, long long rows, int n, int ld)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ld) return;
    const long long r = i / ld;
    const int c = (int)(i - r * ld);
    if constexpr (MODE == 3)
    {
        out[i] = c < n ? reinterpret_cast<const int8_t *>(in_)[r * n + c] : (int8_t)0;
    }
    else
    {
        const float v = c < n ? reinterpret_cast<const float *>(in_)[r * n + c] : 0.0f;
        if constexpr (MODE == 0)
            out[i] = __float2bfloat16_rn(v);
        else if constexpr (MODE == 1)
            out[i] = v;
        else
        {
            // clamp(rintf(x * 128), -128, 127): same expression as oracle_quantize_q17
            float qv = rintf(v * 128.0f);
            qv = fminf(127.0f, fmaxf(-128.0f, qv));
            out[i] = (int8_t)qv;
        }
    }
}

template <typename OutT, int MODE>
static cudaError_t launch_convert(const void *in, OutT *out, long long rows, int n, int ld, cudaStream_t stream)
{
    if (rows <= 0) return cudaSuccess;
    const long long total = rows * ld;
    convert_rows_kernel<OutT, MODE><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, rows, n, ld);
    return cudaGetLastError();
}

cudaError_t launch_convert_rows_bf16(const float *in, void *out, long long rows, int n, int ld, cudaStream_t stream)
{
    return launch_convert<__nv_bfloat16, 0>(in, reinterpret_cast<__nv_bfloat16 *>(out), rows, n, ld, stream);
}
cudaError_t launch_convert_rows_f32(const float *in, float *out, long long rows, int n, int ld, cudaStream_t stream)
{
    return launch_convert<float, 1>(in, out, rows, n, ld, stream);
}
cudaError_t launch_quantize_rows_q17(const float *in, int8_t *out, long long rows, int n, int ld, cudaStream_t stream)
{
    return launch_convert<int8_t, 2>(in, out, rows, n, ld, stream);
}
// INT8 weights W[fan_out][ld] -> the streaming layout of mlp_umma_stream.cu's cluster kernel: blocks of 128 neurons x 128 bytes of K
// (16 KB, zero-padded at the ragged edges), block (T, kb) at ((T * nkb) + kb) * 16 KB, and inside a block the byte order a TMA load with
// CU_TENSOR_MAP_SWIZZLE_128B would have produced in shared memory (16-byte chunk c of row r at r * 128 + ((c ^ (r & 7)) * 16)) -- so that
// a plain 1-D bulk copy of a block (or of its 64-row half) lands ready for tcgen05.mma, and the bytes a CTA streams are contiguous in HBM.
__global__ void retile_i8_weights_kernel(const int8_t *__restrict__ w, long long ld, int fan_out, int fan_in, int nkb, int8_t *__restrict__ out,
                                         long long n_chunks)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_chunks; i += (long long)gridDim.x * blockDim.x)
    {
        const int cs = (int)(i & 7), r = (int)((i >> 3) & 127); // chunk position and row inside the block
        const long long blk = i >> 10;
        const int kb = (int)(blk % nkb);
        const long long T = blk / nkb;
        const int c = cs ^ (r & 7); // the logical chunk stored at this position
        const long long row = T * 128 + r;
        const int col = kb * 128 + c * 16;
        int4 v = make_int4(0, 0, 0, 0);
        if (row < fan_out && col < fan_in) v = *reinterpret_cast<const int4 *>(w + row * ld + col); // (fan_in and ld are multiples of 16)
        reinterpret_cast<int4 *>(out)[i] = v;
    }
}

cudaError_t launch_retile_i8_weights(const int8_t *w, long long ld, int fan_out, int fan_in, int8_t *out, cudaStream_t stream)
{
    if ((fan_in & 15) || (ld & 15)) return cudaErrorInvalidValue;
    const int nkb = (fan_in + 127) / 128;
    const long long n_chunks = (long long)((fan_out + 127) / 128) * nkb * 1024;
    const int grid = (int)std::min<long long>((n_chunks + 255) / 256, 148 * 16);
    retile_i8_weights_kernel<<<grid, 256, 0, stream>>>(w, ld, fan_out, fan_in, nkb, out, n_chunks);
    return cudaGetLastError();
}

cudaError_t launch_pad_rows_i8(const int8_t *in, int8_t *out, long long rows, int n, int ld, cudaStream_t stream)
{
    return launch_convert<int8_t, 3>(in, out, rows, n, ld, stream);
}

__global__ void __launch_bounds__(256)
splitk_finalize_kernel(int32_t *__restrict__ ws, const int32_t *__restrict__ bias, void *__restrict__ out, long long ldo, int out_is_s8, int relu,
                       int m, int n)
{
    // 4 consecutive columns per thread: int4 in, one packed word (int8) or one int4 (int32) out
    const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n4 = n >> 2;
    if (i4 >= (long long)m * n4) return;
    const int row = (int)(i4 / n4), col = (int)(i4 - (long long)row * n4) * 4;
    int4 *wp = reinterpret_cast<int4 *>(ws + (long long)row * n + col);
    const int4 a = *wp, b = *reinterpret_cast<const int4 *>(bias + col);
    *wp = make_int4(0, 0, 0, 0);
    int v[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
#pragma unroll
    for (int e = 0; e < 4; e++)
        if (relu) v[e] = max(v[e], 0);
    if (out_is_s8)
    {
        uint32_t word = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) word |= ((uint32_t)min(127, max(-128, v[e] >> 7)) & 0xFFu) << (8 * e);
        *reinterpret_cast<uint32_t *>(reinterpret_cast<int8_t *>(out) + (long long)row * ldo + col) = word;
    }
    else
        *reinterpret_cast<int4 *>(reinterpret_cast<int32_t *>(out) + (long long)row * ldo + col) = make_int4(v[0], v[1], v[2], v[3]);
}

cudaError_t launch_splitk_finalize(int32_t *ws, const int32_t *bias, void *out, long long ldo, bool out_is_s8, bool relu, int m, int n,
                                   cudaStream_t stream)
{
    if (m <= 0 || n <= 0) return cudaSuccess;
    if ((n & 3) || (ldo & 3)) return cudaErrorInvalidValue;
    const long long total = (long long)m * (n >> 2);
    splitk_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(ws, bias, out, ldo, out_is_s8 ? 1 : 0, relu ? 1 : 0, m, n);
    return cudaGetLastError();
}

__global__ void dequant_q214_kernel(const int32_t *__restrict__ in, float *__restrict__ out, long long count)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (float)in[i] * (1.0f / 16384.0f);
}

cudaError_t launch_dequant_q214(const int32_t *in, float *out, long long count, cudaStream_t stream)
{
    if (count <= 0) return cudaSuccess;
    dequant_q214_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(in, out, count);
    return cudaGetLastError();
}

} // namespace nc

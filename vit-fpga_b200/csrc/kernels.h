// kernels.h -- host-side launchers of every device kernel in the library (internal header).
// Each launcher enqueues exactly one kernel on `stream` and returns the CUDA status of the launch.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <cstdarg>

namespace nc
{

// Programmatic dependent launch (ptx.cuh, griddep_wait): the next kernel's launch and prologue overlap the tail of the running one.
// Measured: a loss at the bench's pass size (ViT-B, 512 images: 24.8 -> 24.2 k images/s; kernel boundaries are not where the time goes)
// but 9 % on ViT-Tiny at 256 images and a quarter of the time of a single-sample ViT pass (831 -> 627 us, ~90 kernels of a few
// microseconds).  So the runtime switches it on for passes of short kernels (g_pdl_small_pass); NETCUDA_PDL=0 / 1 forces it off / on.
extern thread_local bool g_pdl_small_pass;
inline bool pdl_enabled()
{
    static const int forced = getenv("NETCUDA_PDL") ? atoi(getenv("NETCUDA_PDL")) : -1;
    return forced >= 0 ? forced != 0 : g_pdl_small_pass;
}

// Launch with an optional cluster size and, when pdl_enabled(), programmatic stream serialization.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    const int pdl = pdl_enabled() ? 1 : 0;
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = pdl;
    n++;
    if (cluster > 1)
    {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster, attr[n].val.clusterDim.y = 1, attr[n].val.clusterDim.z = 1;
        n++;
    }
    cfg.attrs = attr, cfg.numAttrs = (unsigned)n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Operand kinds of the dense kernel (match ptx.cuh KIND_*), plus the CUDA-core fp32 path.
enum : int
{
    GK_BF16 = 0,
    GK_TF32 = 1,
    GK_I8 = 2,
    GK_FP32_SIMT = 3
};

struct GemmCall
{
    int kind;                // GK_*
    int variant;             // 0 = tcgen05, 1 = CUDA-core reference with identical operand types
    const void *a;           // [M][lda] operand type
    long long lda;
    int a_rows;              // rows addressable behind `a` (>= M); TMA bounds
    const void *w;           // [N][ldw]
    long long ldw;
    const void *bias;        // float[N] or int32[N] (int8 kind) or null
    void *out;
    long long ldc;
    int out_type;            // OUT_* (gemm_tcgen05.cuh)
    int epi;                 // EPI_*
    int m, n, k;
    int k_splits;            // > 1: split-K with reduce-add into an int32 output (epi must be EPI_SPLITK), see gemm_tcgen05.cuh
    int remap_in, remap_out; // EPI_PATCH
    const float *pos;
    int *error_flag;
    int num_sms;
};

// out = epilogue(A . W^T + bias) -- tcgen05 path or reference path depending on call.variant / kind.
cudaError_t launch_gemm(const GemmCall &call, cudaStream_t stream);
// One-time per-process setup of the tcgen05 kernels (dynamic smem opt-in, driver entry point).
cudaError_t gemm_global_init();

// 2-D row-major tensor map {dim0 (contiguous), dim1} with a {box0, box1} box; 128B swizzle when box0 spans
// 128 bytes and `swizzle128` is set.  dtype_bytes selects the element type (1: u8, 2: bf16, 4: f32).
cudaError_t encode_tma_2d(void *map_out, int dtype_bytes, const void *ptr, long long dim0, long long dim1, long long pitch_bytes,
                          int box0, int box1, bool swizzle128);
cudaError_t encode_tma_3d(void *map_out, int dtype_bytes, const void *ptr, long long dim0, long long dim1, long long dim2, long long stride1_bytes,
                          long long stride2_bytes, int box0, int box1, bool swizzle128);

// ---- INT8 MLP forward for 1..32 samples in one persistent weight-streaming kernel (mlp_stream.cu) ----
constexpr int MLP_STREAM_MAX_LAYERS = 16;
constexpr int MLP_STREAM_MAX_BATCH = 32;
constexpr int MLP_STREAM_LL_MAX_BATCH = 4; // ... of which up to here without a grid barrier (tagged words)
struct MlpStreamLayer
{
    const int8_t *w;     // [fan_out][fan_in], fan_in a multiple of 16, at most 4096
    const int8_t *w_tiled = nullptr; // the same weights in 16 KB streaming blocks (launch_retile_i8_weights), or null
    const int32_t *bias; // [fan_out]
    int fan_in, fan_out;
};
struct MlpStreamParams
{
    MlpStreamLayer layers[MLP_STREAM_MAX_LAYERS];
    int n_layers, batch;
    int max_fan_in;      // widest fan-in of the net (sizes the ring slots)
    unsigned relu_mask;  // bit l: ReLU after layer l
    const int8_t *in;    // [batch][fan_in of layer 0]
    int8_t *act[2];      // hidden activations: layer l writes act[(l + 1) & 1] with pitch fan_out, layer l + 1 reads it
    int32_t *out;        // [batch][fan_out of the last layer]: raw accumulators (+ bias, ReLU if flagged)
    unsigned *barrier;   // four zero-initialised words in device memory: [0], [1] counters (left at zero by every launch), [2] launch epoch
    void *ll[2] = {nullptr, nullptr}; // tagged-word activation buffers (mlp_stream.cu, <= MLP_STREAM_LL_MAX_BATCH samples):
                                      // [samples][fan_out / 4] x {4 activations, tag}; zeroed at allocation; null = grid-barrier path
    int l2_prefetch_tiles = 0;  // mlp_stream.cu: 16-row weight tiles requested into L2 ahead of the shared-memory ring
    long long *debug = nullptr; // optional [n_layers][6] globaltimer stamps of CTA debug_cta (NETCUDA_STREAM_DEBUG_PTR)
    int debug_cta = 0;
    int *error_flag;
};
bool mlp_stream_supported(const MlpStreamParams &p, int grid);
cudaError_t launch_mlp_i8_stream(const MlpStreamParams &p, int grid, cudaStream_t stream);
// ---- the same for up to 128 samples on tcgen05 (mlp_umma_stream.cu): activations through shared memory, accumulators in TMEM ----
constexpr int MLP_UMMA_STREAM_MAX_BATCH = 128;
bool mlp_umma_stream_supported(const MlpStreamParams &p);
cudaError_t launch_mlp_i8_umma_stream(const MlpStreamParams &p, int num_sms, cudaStream_t stream);
// ... and with a cluster of two or four CTAs per neuron tile, each streaming its part of K (partial sums meet in distributed shared memory)
int mlp_umma_cluster_size(const MlpStreamParams &p, int num_sms); // 4, 2, or 0 = only the single-CTA kernel serves this net
cudaError_t launch_mlp_i8_umma_cluster(const MlpStreamParams &p, int num_sms, int mode, cudaStream_t stream); // mode: A/B codes, see the launcher

// y = LayerNorm(x) * gamma + beta; x fp32 rows (pitch ldx), y bf16 rows (pitch ldy) -- or fp32 rows for the tf32 nets.
cudaError_t launch_layernorm(const float *x, long long ldx, const float *gamma, const float *beta, void *y, long long ldy,
                             int rows, int dim, float eps, cudaStream_t stream, bool out_f32 = false);

// softmax(q k^T / sqrt(64)) v per (image, head) on packed bf16 qkv rows; head_dim fixed at 64.
// tokens <= 256: tcgen05 kernel (S and P in tensor memory); longer sequences: mma.sync flash kernel.
// `out` is bf16 [batch * tokens][heads * 64], or fp32 of the same shape when out_f32 is set (the A operand of a tf32 projection).
// variant 1 = the mma.sync cross-check kernel; tc_variant >= 0 picks a build variant of the short-sequence tcgen05 kernel (A/B, tests).
cudaError_t launch_attention(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t stream, int *error_flag = nullptr,
                             int num_sms = 0, int variant = 0, bool out_f32 = false, int tc_variant = -1);

// fp32 NCHW -> bf16 (or, out_f32, fp32) patch rows [batch * np][3 * p * p].
cudaError_t launch_patchify(const float *img, void *patches, int batch, int image_size, int patch_size, cudaStream_t stream, bool out_f32 = false);
// u8 HWC frames -> patch rows of (u8 / 255 - mean[c]) * inv_std[c]
cudaError_t launch_patchify_u8(const uint8_t *img, void *patches, int batch, int image_size, int patch_size, const float *mean,
                               const float *inv_std, cudaStream_t stream, bool out_f32 = false);

// x[b * tokens][:] = cls + pos[0]  (fp32 residual stream rows of the class token)
cudaError_t launch_cls_rows(float *x, const float *cls, const float *pos, int batch, int tokens, int dim, cudaStream_t stream);

// Row-wise conversions of the fp32 API inputs into the operand type, zero-padding [n, ld).
cudaError_t launch_convert_rows_bf16(const float *in, void *out, long long rows, int n, int ld, cudaStream_t stream);
cudaError_t launch_convert_rows_f32(const float *in, float *out, long long rows, int n, int ld, cudaStream_t stream);
cudaError_t launch_quantize_rows_q17(const float *in, int8_t *out, long long rows, int n, int ld, cudaStream_t stream);
cudaError_t launch_pad_rows_i8(const int8_t *in, int8_t *out, long long rows, int n, int ld, cudaStream_t stream);
// W[fan_out][ld] int8 -> 16 KB blocks of 128 neurons x 128 bytes of K in shared-memory (128B-swizzled) byte order: the layout the
// cluster streaming kernel copies with 1-D bulk copies.  `out` holds mlp_tiled_weight_bytes(fan_in, fan_out) bytes.
cudaError_t launch_retile_i8_weights(const int8_t *w, long long ld, int fan_out, int fan_in, int8_t *out, cudaStream_t stream);
inline size_t mlp_tiled_weight_bytes(int fan_in, int fan_out) { return (size_t)((fan_out + 127) / 128) * ((fan_in + 127) / 128) * 16384; }
// Second half of a split-K int8 layer: v = ws + bias (optionally max(v, 0)); out = int8 requantisation (clamp(v >> 7)) or the
// int32 itself; ws is set back to zero for the next layer.
cudaError_t launch_splitk_finalize(int32_t *ws, const int32_t *bias, void *out, long long ldo, bool out_is_s8, bool relu, int m, int n,
                                   cudaStream_t stream);
// out_f32 = (float)acc * 2^-14   (INT8 nets through the float API)
cudaError_t launch_dequant_q214(const int32_t *in, float *out, long long count, cudaStream_t stream);

// Host-side staging copy (staging.cpp): pageable caller memory -> page-locked slot, on the process-wide copy-thread pool with
// non-temporal stores.  Returns when every byte has landed.
void staging_copy(void *dst, const void *src, size_t bytes);
int staging_threads();

// Sets the calling thread's netcuda_last_error() text and returns `code` (runtime.cu; shared with weights_io.cu).
int set_last_error_v(int code, const char *fmt, va_list ap);

} // namespace nc

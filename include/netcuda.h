/* netcuda.h -- C ABI of the B200 (sm_100a) backend for VIT-FPGA's net forward path.
 *
 * This is the drop-in boundary as a shared library (libnetcuda.so): plain pointers and sizes,
 * an opaque handle, `int` status codes.  No C++ or torch types cross it.  Every entry point
 * replaces one step of the OpenCL life cycle in the reference's src/netFPGA.cpp (cited per
 * function, `file:line` relative to the reference tree); INTEGRATION.md shows the C++ class
 * (include/netCUDA.h) and a ctypes stub bound on top of it.
 *
 * Conventions
 *   - every function returns NETCUDA_OK (0) or a NETCUDA_ERR_* code and never calls exit()
 *     (the reference's AOCLUtils checkError prints + exit()s, e.g. src/netFPGA.cpp:274);
 *   - netcuda_last_error() returns a thread-local, human readable description of the last
 *     failure on the calling thread (role of aocl_utils::printError);
 *   - host buffers are borrowed for the duration of the call; device pointers must live on the
 *     handle's device; `stream` arguments are `cudaStream_t` passed as `void*` (NULL = the
 *     handle's own stream);
 *   - all weight matrices are row-major W[out][in], the reference's flat layout
 *     (src/netFPGA.cpp:91-106).
 */
#ifndef NETCUDA_H
#define NETCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NETCUDA_ABI_VERSION 1

/* status codes */
#define NETCUDA_OK 0
#define NETCUDA_ERR_INVALID 1     /* bad argument / shape / state            */
#define NETCUDA_ERR_CUDA 2        /* a CUDA runtime or driver call failed    */
#define NETCUDA_ERR_UNSUPPORTED 3 /* valid request this build cannot serve   */
#define NETCUDA_ERR_NO_DEVICE 4   /* no sm_100 device visible                */
#define NETCUDA_ERR_KERNEL 5      /* a kernel reported a pipeline time-out   */
#define NETCUDA_ERR_RING_FULL 6   /* frame ring: every slot is in flight     */
#define NETCUDA_ERR_RING_EMPTY 7  /* frame ring: nothing to pop              */

/* net kinds */
#define NETCUDA_KIND_MLP 0 /* what net::net_data describes (def/defines.h:14-23) */
#define NETCUDA_KIND_VIT 1 /* vision transformer; no reference counterpart       */

/* arithmetic */
#define NETCUDA_PREC_FP32 0 /* CUDA-core fp32, k-ascending fmaf: bit-equal to the oracle   */
#define NETCUDA_PREC_TF32 1 /* tcgen05 kind::tf32, fp32 accumulate (MLPs; ViT linear layers) */
#define NETCUDA_PREC_BF16 2 /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate           */
#define NETCUDA_PREC_INT8 3 /* tcgen05 kind::i8, Q1.7 operands, int32 accumulate (MLP only) */

/* MLP activation placement.  The reference hints at a single net-wide "RELU2"
 * (src/netFPGA.cpp:79) that never reaches the device; see DESIGN.md for the choice. */
#define NETCUDA_ACT_RELU_HIDDEN 0 /* max(0,.) on every layer but the last (default) */
#define NETCUDA_ACT_RELU_ALL 1    /* max(0,.) on every layer                        */
#define NETCUDA_ACT_NONE 2        /* purely linear                                  */

typedef struct netcuda_net netcuda_t; /* opaque */

/* Describes the net to build.  Unused fields for a kind must be 0. */
typedef struct netcuda_desc
{
    int32_t kind;       /* NETCUDA_KIND_*                                            */
    int32_t precision;  /* NETCUDA_PREC_*                                            */
    int32_t device;     /* CUDA ordinal                                              */
    int32_t activation; /* NETCUDA_ACT_* (MLP)                                       */
    int32_t max_batch;  /* samples processed per internal pass (0 = library default) */
    /* MLP: mirrors net_fpga's n_ins / n_layers / n_p_l (include/netFPGA.h:22-26) */
    int32_t n_ins;
    int32_t n_layers;
    const int32_t *n_p_l; /* n_layers entries, borrowed during netcuda_create only */
    /* ViT */
    int32_t image_size; /* square input, e.g. 224   */
    int32_t patch_size; /* e.g. 16                  */
    int32_t dim;        /* embedding width D        */
    int32_t depth;      /* encoder blocks L         */
    int32_t heads;      /* D / heads must be 64     */
    int32_t mlp_dim;    /* hidden width of the MLP  */
    int32_t n_classes;  /* head outputs             */
} netcuda_desc;

/* ---- life cycle ------------------------------------------------------------------------ */

/* Number of usable CUDA devices.  Replaces platform/device discovery,
 * src/netFPGA.cpp:372-377 (clGetPlatformIDs / clGetDeviceIDs). */
int netcuda_device_count(int *count);

/* Create context, stream set, device arenas and pinned staging for `desc`.
 * Replaces _init_program + _init_kernel, src/netFPGA.cpp:367-441. */
int netcuda_create(const netcuda_desc *desc, netcuda_t **out);

/* Release everything owned by the handle.  Replaces cleanup(), src/netFPGA.cpp:639-651. */
int netcuda_destroy(netcuda_t *h);

/* ---- weights ---------------------------------------------------------------------------- */

/* Upload an MLP in the reference's flat layout: `w_flat` = all layers' W[out][in] matrices back
 * to back, `b_flat` = all biases back to back (src/netFPGA.cpp:91-106).  Converted on upload to
 * the handle's precision (INT8: q = clamp(rint(w*128)), bias rint(b*16384), see oracle/).
 * Replaces _load_params, src/netFPGA.cpp:484-515. */
int netcuda_upload_mlp(netcuda_t *h, const float *w_flat, const float *b_flat);

/* INT8 handles only: upload already-quantised Q1.7 weights and Q2.14 int32 biases. */
int netcuda_upload_mlp_i8(netcuda_t *h, const int8_t *w_flat, const int32_t *b_flat);

/* Number of floats netcuda_upload_vit expects for `desc` (layout documented in DESIGN.md):
 * patch_w[D][3PP] patch_b[D] cls[D] pos[N][D] { ln1_g ln1_b qkv_w[3D][D] qkv_b proj_w[D][D]
 * proj_b ln2_g ln2_b fc1_w[F][D] fc1_b fc2_w[D][F] fc2_b } x depth, lnf_g lnf_b head_w[C][D] head_b. */
int netcuda_vit_param_count(const netcuda_desc *desc, size_t *count);

/* Upload a ViT from that flat fp32 layout. */
int netcuda_upload_vit(netcuda_t *h, const float *flat, size_t count);

/* ---- weight files ------------------------------------------------------- *
 * On-disk image of the flat in-memory layout the reference builds at src/netFPGA.cpp:91-106 (and that
 * get_net_data, :206-237, is meant to invert), so that a net survives the process:
 *
 *   offset  0  char[8]  "NETCUDAW"
 *           8  u32 version (1)      12  u32 kind (NETCUDA_KIND_*)   16  u32 dtype (NETCUDA_FILE_*)
 *          20  u32 activation       24  u32 n_ins                   28  u32 n_layers
 *          32  u32 image_size, patch_size, dim, depth, heads, mlp_dim, n_classes   (ViT; 0 for an MLP)
 *          60  u32 reserved (0)     64  u64 n_weights               72  u64 n_biases
 *          80  u32 crc32 (IEEE, of every byte after the 96-byte header)           84  12 bytes 0
 *          96  i32 n_p_l[n_layers], zero-padded to a multiple of 16 bytes
 *          ..  weights: n_weights elements (f32, or i8 for NETCUDA_FILE_Q17), zero-padded to 16 bytes;
 *              MLP: all W[out][in] back to back; ViT: the flat vector of netcuda_vit_param_count
 *          ..  biases: n_biases elements (f32, or i32 Q2.14 for NETCUDA_FILE_Q17); none for a ViT
 * Little endian.  The file functions are pure host code (no GPU needed) except netcuda_create_from_file. */
#define NETCUDA_FILE_F32 0 /* fp32 weights and biases                   */
#define NETCUDA_FILE_Q17 1 /* int8 Q1.7 weights, int32 Q2.14 biases     */
#define NETCUDA_FILE_MAX_LAYERS 64

typedef struct netcuda_file_info
{
    netcuda_desc desc; /* kind, activation, n_ins, n_layers, ViT dims; precision = the file's natural one
                          (INT8 for Q17, BF16 for a ViT, FP32 for an fp32 MLP); n_p_l points at n_p_l below */
    int32_t n_p_l[NETCUDA_FILE_MAX_LAYERS];
    int32_t dtype; /* NETCUDA_FILE_* */
    uint64_t n_weights, n_biases;
} netcuda_file_info;

int netcuda_file_write_mlp(const char *path, const int32_t *n_p_l, int n_layers, int n_ins, int activation,
                           const float *w_flat, const float *b_flat);
int netcuda_file_write_mlp_i8(const char *path, const int32_t *n_p_l, int n_layers, int n_ins, int activation,
                              const int8_t *w_flat, const int32_t *b_flat);
int netcuda_file_write_vit(const char *path, const netcuda_desc *desc, const float *flat, size_t count);
/* Header only (validates magic, version, sizes against the file length). */
int netcuda_file_info_read(const char *path, netcuda_file_info *info);
/* Payload into caller buffers of exactly n_weights / n_biases elements of the file's dtype; checks the CRC. */
int netcuda_file_read(const char *path, void *weights, size_t weight_bytes, void *biases, size_t bias_bytes);
/* netcuda_create + upload from a file.  precision < 0 = the file's natural precision (an fp32 MLP file can be
 * opened as FP32 / TF32 / BF16 / INT8, a Q17 file only as INT8, a ViT as BF16 or TF32); max_batch 0 = default. */
int netcuda_create_from_file(const char *path, int precision, int device, int max_batch, netcuda_t **out);

/* ---- forward ---------------------------------------------------------------------------- */

/* Host-to-host forward: `in` holds batch*n_in floats, `out` receives batch*n_out floats.
 * Synchronous like the reference (blocking read at src/netFPGA.cpp:277) but batched; inputs are
 * staged through pinned memory (or DMA'd in place when `in` is already page-locked) in
 * max_batch-sized passes, with the copy of pass i+1 overlapping the compute of pass i.
 * Replaces the H2D / task / D2H triple at src/netFPGA.cpp:266-277. */
int netcuda_forward(netcuda_t *h, const float *in, size_t batch, float *out);

/* Page-lock a caller-owned host range (cudaHostRegister, portable across devices) so that netcuda_forward / netcuda_submit DMA
 * straight from / into it instead of staging through the library's own pinned slots -- the reference keeps its host vectors for
 * the life of the net (src/netFPGA.cpp:78-107), and a caller that feeds the same buffer again and again pays the page-locking
 * once.  The range must stay allocated until netcuda_host_unregister(p) (same start address).  Registering a range that already
 * is page-locked is not an error. */
int netcuda_host_register(const void *p, size_t bytes);
int netcuda_host_unregister(const void *p);

/* Asynchronous form of netcuda_forward: the reference chains write -> task -> read with events
 * but then blocks in the read (src/netFPGA.cpp:273-277), so its caller can never overlap two samples.
 * netcuda_submit enqueues the same H2D / kernels / D2H sequence and returns a ticket at once; the H2D copy of call
 * i+1 runs on the copy engine while the kernels of call i execute.  `in` and `out` must stay valid until
 * netcuda_wait(ticket) returns; they should be page-locked (cudaHostAlloc / cudaHostRegister) -- pageable buffers
 * work but are staged with host memcpys inside submit / wait.  Up to 4 calls may be in flight per handle; a fifth
 * submit first retires the oldest.  Tickets start at 1.  netcuda_query sets *done without blocking.
 * One thread per handle (distinct handles may be driven from distinct threads). */
int netcuda_submit(netcuda_t *h, const float *in, size_t batch, float *out, uint64_t *ticket);
int netcuda_wait(netcuda_t *h, uint64_t ticket);
int netcuda_query(netcuda_t *h, uint64_t ticket, int *done);

/* Device-resident forward on `stream` (asynchronous; no host copies): the roofline path.
 * d_in: fp32 [batch][n_in], d_out: fp32 [batch][n_out]. */
int netcuda_forward_device(netcuda_t *h, const void *d_in, size_t batch, void *d_out, void *stream);

/* INT8 handles: raw quantised forward, int8 [batch][n_in] -> int32 [batch][n_out]
 * (accumulator of the last layer, Q2.14).  Used for the bit-exactness tests. */
int netcuda_forward_device_i8(netcuda_t *h, const int8_t *d_in, size_t batch, int32_t *d_out, void *stream);
int netcuda_forward_i8(netcuda_t *h, const int8_t *in, size_t batch, int32_t *out);

/* ---- u8 frames for vision transformers ---------------------------------- *
 * The reference moves camera frames as unsigned char vectors (net::image_set::resized_image_data, def/defines.h:31-38,
 * staged at src/netFPGA.cpp:314-318).  These entry points feed a ViT from such frames directly: `frames` is
 * batch x H x W x 3 bytes (interleaved RGB, row-major); the first kernel turns them into normalised bf16 patch rows,
 * v = (u8 / 255 - mean[c]) * (1 / stddev[c]) in fp32 -- a quarter of the PCIe bytes of the float path and no fp32 image
 * in HBM.  Default mean = stddev = 0.5 (0..255 -> [-1, 1], the reference's MIN_RANGE / MAX_RANGE). */
int netcuda_set_u8_normalization(netcuda_t *h, const float *mean, const float *stddev); /* 3 values each */
int netcuda_forward_u8(netcuda_t *h, const uint8_t *frames, size_t batch, float *out);
int netcuda_submit_u8(netcuda_t *h, const uint8_t *frames, size_t batch, float *out, uint64_t *ticket);
int netcuda_forward_device_u8(netcuda_t *h, const uint8_t *d_frames, size_t batch, float *d_out, void *stream);

/* ---- introspection ---------------------------------------------------------------------- */

int netcuda_n_in(const netcuda_t *h, size_t *n);  /* floats per input sample  */
int netcuda_n_out(const netcuda_t *h, size_t *n); /* floats per output sample */

/* Wall-clock microseconds of the last netcuda_forward call, transfers included.
 * Replaces get_forward_performance, src/netFPGA.cpp:603-611 (stopwatch at :262-284). */
int netcuda_last_forward_us(const netcuda_t *h, int64_t *us);

/* Kernels launched by this handle since creation (bench.py reports the per-step delta). */
int netcuda_launch_count(const netcuda_t *h, uint64_t *count);

/* Algorithmic multiply-accumulate FLOPs (2*MACs of the dense contractions) per sample. */
int netcuda_flops_per_sample(const netcuda_t *h, double *flops);

/* Per-kernel timing (replaces the reference's single PERFORMANCE stopwatch, src/netFPGA.cpp:262-284,
 * with one CUDA-event pair around every kernel this handle launches, recorded on the launching stream).
 * Off by default; while on, each launch costs two extra event records.  netcuda_profile_read waits for
 * the recorded work, adds the durations up per launch-site label ("qkv", "fc1", "layernorm", ...), writes
 * up to `cap` entries, returns the number of labels in *count and clears the records. */
typedef struct netcuda_kernel_stat
{
    char label[32];
    uint64_t launches;
    double ms;    /* summed device time of those launches (CUDA events)                     */
    double flops; /* summed algorithmic FLOPs / integer ops (2*MACs); 0 for HBM-bound kernels */
    double bytes; /* summed algorithmic bytes (minimum HBM traffic: operands read once + output written once) */
} netcuda_kernel_stat;
int netcuda_profile_enable(netcuda_t *h, int on);
int netcuda_profile_read(netcuda_t *h, netcuda_kernel_stat *stats, int cap, int *count);

/* Select a debugging/measurement variant of the dense kernel for this handle:
 * 0 = default (tcgen05, CTA pairs on large problems), 1 = CUDA-core reference GEMM with the same operand
 * rounding (also selects the mma.sync attention kernel), 2 = tcgen05 with one CTA per tile everywhere,
 * 3 = default but with 16 instead of 8 epilogue warps in the GELU GEMM, 4 / 5 = two output slabs per epilogue warp (and a 5-stage
 * ring) in every / in no CTA-pair GEMM -- the default picks per epilogue (A/B measurements; same bits as 0). */
int netcuda_set_gemm_variant(netcuda_t *h, int variant);

const char *netcuda_last_error(void);
int netcuda_abi_version(void);

/* ---- single-kernel entry points (tests/, profiles/) -------------------------------------- *
 * Each launches exactly one hot-path kernel on device pointers so that it can be checked
 * against the oracle in isolation.  `device` selects the GPU; stream NULL = legacy default. */

#define NETCUDA_OUT_F32 0
#define NETCUDA_OUT_BF16 1
#define NETCUDA_OUT_S8 2
#define NETCUDA_OUT_S32 3

#define NETCUDA_EPI_NONE 0
#define NETCUDA_EPI_RELU 1
#define NETCUDA_EPI_GELU 2     /* exact (erf) GELU                                    */
#define NETCUDA_EPI_RESIDUAL 3 /* out(fp32) = out + acc + bias (in-place residual add) */
#define NETCUDA_EPI_REQUANT 4  /* INT8: q = min(127, max(0,acc+bias) >> 7)            */

/* out[M][N] = epi(A[M][K] . W[N][K]^T + bias[N]).  precision selects operand type
 * (BF16: bf16, TF32: fp32, INT8: int8, FP32: fp32 on CUDA cores); lda/ldw/ldc in elements. */
int netcuda_op_gemm(int device, int precision, int variant,
                    const void *d_a, int lda, const void *d_w, int ldw, const void *d_bias,
                    void *d_out, int ldc, int out_type, int epilogue, int m, int n, int k, void *stream);

/* y[r][:] = (x[r][:] - mean) * rstd * gamma + beta ; x fp32 (row stride ldx), y bf16 (row stride ldy). */
int netcuda_op_layernorm(int device, const float *d_x, int ldx, const float *d_gamma, const float *d_beta,
                         void *d_y, int ldy, int rows, int dim, float eps, void *stream);

/* Multi-head attention core on packed qkv: d_qkv bf16 [batch*tokens][3*heads*64] (q|k|v blocks,
 * head-major inside each), d_out bf16 [batch*tokens][heads*64] = softmax(q k^T / 8) v per head. */
int netcuda_op_attention(int device, const void *d_qkv, void *d_out, int batch, int tokens, int heads, void *stream);
/* The same with an explicit kernel choice and output type (tests and A/B measurements):
 *   kernel < 0: the product default;  NETCUDA_ATT_KERNEL_MMA_SYNC: the mma.sync cross-check kernel;
 *   otherwise a build variant of the short-sequence tcgen05 kernels (csrc/attention.cu): 0..2 = the 12-warp kernel (0 = one polling
 *   MMA issuer, 1 = one blocking issuer per query tile, 2 = 1 + exp2 turn-taking); 4 = the 16-softmax-warp kernel (the default),
 *   14 / 24 / 34 its double-buffered / balanced-split forms, + 100 without turn-taking between the query tiles.
 *   out_f32 != 0: d_out is fp32 [batch*tokens][heads*64] (what a TF32 ViT's output projection reads). */
#define NETCUDA_ATT_KERNEL_MMA_SYNC 100
int netcuda_op_attention_ex(int device, const void *d_qkv, void *d_out, int batch, int tokens, int heads, int kernel, int out_f32,
                            void *stream);

/* fp32 NCHW images -> bf16 patch matrix [batch*np][3*p*p] (column = c*p*p + py*p + px). */
int netcuda_op_patchify(int device, const float *d_img, void *d_patches, int batch, int image_size, int patch_size, void *stream);

/* ---- frame ring: the image side channel (net_abstract::filter_image / get_filtered_image) ------------- *
 * Replaces the 24-slot event-chained ring of src/netFPGA.cpp:292-365 (BATCH_SIZE, :12): push = filter_image (copy the
 * frame, H2D, `image_process`, non-blocking D2H; NETCUDA_ERR_RING_FULL when every slot is in flight -- the reference's
 * "PILA LLENA", :333), pop = get_filtered_image (wait for the oldest slot, :349; NETCUDA_ERR_RING_EMPTY = "PILA VACIA",
 * :359).  Frames are single-channel u8, h * w <= max_pixels.  `image_process` is absent from the reference; the filter
 * here is a 3 x 3 binomial smoothing with replicated borders in integer arithmetic (csrc/frame_ring.cu). */
typedef struct netcuda_ring netcuda_ring_t;
#define NETCUDA_RING_DEPTH 24 /* BATCH_SIZE, src/netFPGA.cpp:12 */
int netcuda_ring_create(int device, int depth, size_t max_pixels, netcuda_ring_t **out);
int netcuda_ring_push(netcuda_ring_t *r, const uint8_t *pixels, size_t h, size_t w);
int netcuda_ring_pop(netcuda_ring_t *r, uint8_t *pixels_out, size_t capacity, size_t *h, size_t *w);
int netcuda_ring_peek(netcuda_ring_t *r, size_t *h, size_t *w); /* dimensions of the oldest frame in flight */
int netcuda_ring_in_flight(const netcuda_ring_t *r, int *count, uint64_t *dropped);
int netcuda_ring_destroy(netcuda_ring_t *r);
/* The filter alone on device buffers (h x w bytes each). */
int netcuda_op_filter3x3(int device, const uint8_t *d_in, uint8_t *d_out, int h, int w, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NETCUDA_H */

/* oracle.h -- CPU restatement of the net forward path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
 * load this library.  The product (libnetcuda.so, include/netCUDA.h) never links or calls it and
 * has no CPU fallback.
 *
 * PARITY STATUS: "parity unpinned" for the arithmetic.  The reference ships neither its device
 * kernel (`network_v1` lives in an absent vector_kernels.aocx, src/netFPGA.cpp:250,388-390) nor a
 * single test or golden vector.  What IS pinned:
 *   - the weight/bias memory layout and the I/O contract, against the reference's own host code
 *     (oracle/_ref: the unmodified src/netFPGA.cpp driven through an OpenCL shim, see
 *     oracle/Makefile), and
 *   - the random-init rule `float(rand()%200-100)/100` (src/netFPGA.cpp:82-88) as a known-answer
 *     test (tests/test_oracle.py), and
 *   - the ViT restatement, against torchvision's VisionTransformer (tests/golden/).
 * The activation ("RELU2", src/netFPGA.cpp:79, never sent to the device) and the INT8 format are
 * builder decisions documented in DESIGN.md, not reference behaviour.
 */
#ifndef NETCUDA_ORACLE_H
#define NETCUDA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_ACT_RELU_HIDDEN 0
#define ORACLE_ACT_RELU_ALL 1
#define ORACLE_ACT_NONE 2

/* ---- MLP, fp32 -------------------------------------------------------------------------- *
 * Layout follows src/netFPGA.cpp:91-106: `params` = per layer a row-major W[out][in] matrix,
 * layers back to back; `bias` = per-layer biases back to back; npl = fan-outs (int[], as the
 * kernel argument at src/netFPGA.cpp:409,435).  Arithmetic, fixed here:
 *     acc = bias[j];  for k ascending: acc = fmaf(W[j][k], h[k], acc);  h'[j] = act(acc)
 * One sample: the exact shape of a `network_v1` task (src/netFPGA.cpp:275). */
void oracle_mlp_forward_one(const float *inputs, const float *params, const float *bias, float *outs,
                            const int *npl, int n_layers, int n_ins, int activation);

/* Batched: in [batch][n_ins] -> out [batch][npl[last]], OpenMP over samples when threads > 1. */
void oracle_mlp_forward(const float *in, size_t batch, const float *params, const float *bias, float *out,
                        const int *npl, int n_layers, int n_ins, int activation, int threads);

/* The reference's random initialisation (src/netFPGA.cpp:82-88): srand(seed), then every param,
 * then every bias, each float(rand() % 200 - 100) / 100. */
void oracle_rand_init(unsigned seed, float *params, size_t n_params, float *bias, size_t n_neurons);

/* ---- MLP, INT8 Q1.7 (builder-defined) ------------------------------------------ *
 * weights/activations int8 (value = q/128), bias int32 Q2.14.
 *     acc = bias[j] + sum_k int32(h[k]) * int32(W[j][k])
 *     hidden (ReLU) layers: h'[j] = min(127, max(0, acc) >> 7);  last layer: out[j] = acc (int32)
 * With activation RELU_ALL the last layer's acc is clamped at 0 before being returned;
 * with NONE hidden layers requantise as clamp(acc >> 7, -128, 127) (arithmetic shift). */
void oracle_quantize_q17(const float *x, size_t n, int8_t *q);      /* clamp(rintf(x*128), -128, 127) */
void oracle_quantize_bias_q214(const float *b, size_t n, int32_t *q); /* (int32) rintf(b*16384)        */
void oracle_mlp_forward_i8(const int8_t *in, size_t batch, const int8_t *params, const int32_t *bias,
                           int32_t *out, const int *npl, int n_layers, int n_ins, int activation, int threads);

/* ---- ViT, fp32 ----------------------------------------------------------------------------- *
 * Dosovitskiy et al. as implemented by torchvision.models.vision_transformer (pre-norm encoder,
 * cls token, learned position embedding, LayerNorm eps 1e-6, exact-erf GELU, head on the cls
 * token).  `flat` uses the layout documented at netcuda_vit_param_count (include/netcuda.h). */
typedef struct oracle_vit_cfg
{
    int image_size, patch_size, dim, depth, heads, mlp_dim, n_classes;
} oracle_vit_cfg;

size_t oracle_vit_param_count(const oracle_vit_cfg *cfg);
/* images fp32 [batch][3][S][S] -> logits [batch][n_classes].  Returns 0, or -1 on bad config. */
int oracle_vit_forward(const oracle_vit_cfg *cfg, const float *flat, const float *images, size_t batch,
                       float *logits, int threads);

/* Building blocks exposed for kernel-level parity tests (all fp32, row-major). */
void oracle_linear(const float *a, size_t m, int k, const float *w, const float *bias, int n, float *out, int threads);
void oracle_layernorm(const float *x, size_t rows, int dim, const float *gamma, const float *beta, float eps, float *y);
void oracle_attention(const float *qkv, size_t batch, int tokens, int heads, int head_dim, float *out, int threads);
float oracle_gelu(float x);

/* ---- image side channel: the ring's device stage (oracle_image.c; builder-defined filter) ---- */
void oracle_filter3x3(const uint8_t *in, uint8_t *out, int h, int w);

int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

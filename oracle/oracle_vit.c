/* oracle_vit.c -- fp32 CPU restatement of the ViT forward + shared dense building blocks.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The reference contains no ViT (def/defines.h:14-23 can only describe an MLP), so there is no
 * reference file to follow for the transformer blocks.  This restates the published algorithm
 * (Dosovitskiy et al., "An Image is Worth 16x16 Words") exactly as torchvision 0.26's
 * torchvision/models/vision_transformer.py implements it; tests/test_oracle.py pins it against
 * torchvision itself and against the committed fixture tests/golden/vit_small_*.npy.
 * Dense layers reuse the reference's W[out][in] row-major convention (src/netFPGA.cpp:91-106).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ROW_BLOCK 4
#define COL_BLOCK 512

float oracle_gelu(float x)
{
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

static float *transpose_w(const float *w, int n, int k)
{
    /* w[n][k] -> wt[k][n] */
    float *wt = (float *)malloc(sizeof(float) * (size_t)n * (size_t)k);
    for (int j = 0; j < n; j++)
        for (int kk = 0; kk < k; kk++)
            wt[(size_t)kk * n + j] = w[(size_t)j * k + kk];
    return wt;
}

/* out[i][j] = bias[j] (+) sum_k a[i][k]*wt[k][j], accumulated as acc = fmaf(w, a, acc) in
 * ascending k for every (i, j): the same operation sequence as the scalar dot product in
 * oracle_mlp_forward_one, just iterated j-innermost so the compiler can vectorise across j. */
static void linear_rows(const float *a, size_t row0, size_t row1, int k, const float *wt, const float *bias, int n,
                        float *out)
{
    float acc[ROW_BLOCK][COL_BLOCK];
    for (size_t i0 = row0; i0 < row1; i0 += ROW_BLOCK)
    {
        const int rb = (int)((row1 - i0) < ROW_BLOCK ? (row1 - i0) : ROW_BLOCK);
        for (int j0 = 0; j0 < n; j0 += COL_BLOCK)
        {
            const int nb = (n - j0) < COL_BLOCK ? (n - j0) : COL_BLOCK;
            for (int r = 0; r < rb; r++)
                for (int j = 0; j < nb; j++)
                    acc[r][j] = bias ? bias[j0 + j] : 0.0f;
            for (int kk = 0; kk < k; kk++)
            {
                const float *restrict wrow = wt + (size_t)kk * n + j0;
                for (int r = 0; r < rb; r++)
                {
                    const float av = a[(i0 + r) * (size_t)k + kk];
                    float *restrict c = acc[r];
#pragma omp simd
                    for (int j = 0; j < nb; j++)
                        c[j] = __builtin_fmaf(wrow[j], av, c[j]);
                }
            }
            for (int r = 0; r < rb; r++)
                memcpy(out + (i0 + r) * (size_t)n + j0, acc[r], sizeof(float) * (size_t)nb);
        }
    }
}

static void linear_pret(const float *a, size_t m, int k, const float *wt, const float *bias, int n, float *out,
                        int threads)
{
    if (threads <= 1 || m < 2 * ROW_BLOCK)
    {
        linear_rows(a, 0, m, k, wt, bias, n, out);
        return;
    }
    const size_t blocks = (m + ROW_BLOCK - 1) / ROW_BLOCK;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long long b = 0; b < (long long)blocks; b++)
    {
        size_t r0 = (size_t)b * ROW_BLOCK;
        size_t r1 = r0 + ROW_BLOCK < m ? r0 + ROW_BLOCK : m;
        linear_rows(a, r0, r1, k, wt, bias, n, out);
    }
}

void oracle_linear(const float *a, size_t m, int k, const float *w, const float *bias, int n, float *out, int threads)
{
    float *wt = transpose_w(w, n, k);
    linear_pret(a, m, k, wt, bias, n, out, threads);
    free(wt);
}

void oracle_layernorm(const float *x, size_t rows, int dim, const float *gamma, const float *beta, float eps, float *y)
{
    for (size_t r = 0; r < rows; r++)
    {
        const float *xr = x + r * (size_t)dim;
        double s = 0.0;
        for (int i = 0; i < dim; i++)
            s += xr[i];
        const double mean = s / dim;
        double v = 0.0;
        for (int i = 0; i < dim; i++)
        {
            const double d = xr[i] - mean;
            v += d * d;
        }
        const float rstd = (float)(1.0 / sqrt(v / dim + (double)eps));
        const float meanf = (float)mean;
        float *yr = y + r * (size_t)dim;
        for (int i = 0; i < dim; i++)
            yr[i] = (xr[i] - meanf) * rstd * gamma[i] + beta[i];
    }
}

/* qkv rows are [q(heads*hd) | k(heads*hd) | v(heads*hd)] per token -- torch's in_proj order. */
static void attention_one(const float *qkv, int tokens, int heads, int hd, float *out, float *scores)
{
    const int d = heads * hd;
    const int ld = 3 * d;
    const float scale = 1.0f / sqrtf((float)hd);
    for (int h = 0; h < heads; h++)
    {
        for (int i = 0; i < tokens; i++)
        {
            const float *q = qkv + (size_t)i * ld + h * hd;
            float mx = -INFINITY;
            for (int j = 0; j < tokens; j++)
            {
                const float *kk = qkv + (size_t)j * ld + d + h * hd;
                float s = 0.0f;
                for (int c = 0; c < hd; c++)
                    s = fmaf(q[c], kk[c], s);
                s *= scale;
                scores[j] = s;
                if (s > mx)
                    mx = s;
            }
            float sum = 0.0f;
            for (int j = 0; j < tokens; j++)
            {
                scores[j] = expf(scores[j] - mx);
                sum += scores[j];
            }
            const float inv = 1.0f / sum;
            float *o = out + (size_t)i * d + h * hd;
            for (int c = 0; c < hd; c++)
                o[c] = 0.0f;
            for (int j = 0; j < tokens; j++)
            {
                const float p = scores[j] * inv;
                const float *v = qkv + (size_t)j * ld + 2 * d + h * hd;
                for (int c = 0; c < hd; c++)
                    o[c] = fmaf(p, v[c], o[c]);
            }
        }
    }
}

void oracle_attention(const float *qkv, size_t batch, int tokens, int heads, int head_dim, float *out, int threads)
{
    const size_t d = (size_t)heads * head_dim;
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
        float *scores = (float *)malloc(sizeof(float) * (size_t)tokens);
#pragma omp for schedule(static)
        for (long long b = 0; b < (long long)batch; b++)
            attention_one(qkv + (size_t)b * tokens * 3 * d, tokens, heads, head_dim, out + (size_t)b * tokens * d, scores);
        free(scores);
    }
}

/* ---- ViT ----------------------------------------------------------------------------------- */

typedef struct
{
    const float *ln1_g, *ln1_b, *qkv_w, *qkv_b, *proj_w, *proj_b, *ln2_g, *ln2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
    float *qkv_wt, *proj_wt, *fc1_wt, *fc2_wt;
} vit_layer;

size_t oracle_vit_param_count(const oracle_vit_cfg *c)
{
    const size_t D = c->dim, F = c->mlp_dim, C = c->n_classes;
    const size_t g = c->image_size / c->patch_size, np = g * g, N = np + 1;
    const size_t pk = 3u * c->patch_size * c->patch_size;
    size_t n = D * pk + D + D + N * D;
    n += (size_t)c->depth * (2 * D + 3 * D * D + 3 * D + D * D + D + 2 * D + F * D + F + D * F + D);
    n += 2 * D + C * D + C;
    return n;
}

int oracle_vit_forward(const oracle_vit_cfg *c, const float *flat, const float *images, size_t batch, float *logits,
                       int threads)
{
    if (!c || c->patch_size <= 0 || c->image_size % c->patch_size || c->heads <= 0 || c->dim % c->heads)
        return -1;
    const int D = c->dim, F = c->mlp_dim, C = c->n_classes, P = c->patch_size, S = c->image_size;
    const int g = S / P, np = g * g, N = np + 1, pk = 3 * P * P, hd = D / c->heads;

    /* carve the flat parameter vector */
    const float *p = flat;
    const float *patch_w = p;
    p += (size_t)D * pk;
    const float *patch_b = p;
    p += D;
    const float *cls = p;
    p += D;
    const float *pos = p;
    p += (size_t)N * D;
    vit_layer *L = (vit_layer *)malloc(sizeof(vit_layer) * (size_t)c->depth);
    for (int l = 0; l < c->depth; l++)
    {
        L[l].ln1_g = p, p += D;
        L[l].ln1_b = p, p += D;
        L[l].qkv_w = p, p += (size_t)3 * D * D;
        L[l].qkv_b = p, p += 3 * D;
        L[l].proj_w = p, p += (size_t)D * D;
        L[l].proj_b = p, p += D;
        L[l].ln2_g = p, p += D;
        L[l].ln2_b = p, p += D;
        L[l].fc1_w = p, p += (size_t)F * D;
        L[l].fc1_b = p, p += F;
        L[l].fc2_w = p, p += (size_t)D * F;
        L[l].fc2_b = p, p += D;
        L[l].qkv_wt = transpose_w(L[l].qkv_w, 3 * D, D);
        L[l].proj_wt = transpose_w(L[l].proj_w, D, D);
        L[l].fc1_wt = transpose_w(L[l].fc1_w, F, D);
        L[l].fc2_wt = transpose_w(L[l].fc2_w, D, F);
    }
    const float *lnf_g = p;
    p += D;
    const float *lnf_b = p;
    p += D;
    const float *head_w = p;
    p += (size_t)C * D;
    const float *head_b = p;
    float *patch_wt = transpose_w(patch_w, D, pk);
    float *head_wt = transpose_w(head_w, C, D);

#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
        float *patches = (float *)malloc(sizeof(float) * (size_t)np * pk);
        float *x = (float *)malloc(sizeof(float) * (size_t)N * D);
        float *y = (float *)malloc(sizeof(float) * (size_t)N * D);
        float *qkv = (float *)malloc(sizeof(float) * (size_t)N * 3 * D);
        float *att = (float *)malloc(sizeof(float) * (size_t)N * D);
        float *hid = (float *)malloc(sizeof(float) * (size_t)N * F);
        float *tmp = (float *)malloc(sizeof(float) * (size_t)N * D);
        float *scores = (float *)malloc(sizeof(float) * (size_t)N);
#pragma omp for schedule(dynamic, 1)
        for (long long b = 0; b < (long long)batch; b++)
        {
            const float *img = images + (size_t)b * 3 * S * S;
            /* conv_proj with stride == kernel == P is a GEMM over flattened patches:
             * patch (gy,gx), column c*P*P + py*P + px  (torch conv weight [D][3][P][P] flattened) */
            for (int gy = 0; gy < g; gy++)
                for (int gx = 0; gx < g; gx++)
                    for (int ch = 0; ch < 3; ch++)
                        for (int py = 0; py < P; py++)
                            memcpy(patches + ((size_t)(gy * g + gx) * pk + ch * P * P + py * P),
                                   img + ((size_t)ch * S + gy * P + py) * S + gx * P, sizeof(float) * (size_t)P);
            linear_rows(patches, 0, np, pk, patch_wt, patch_b, D, x + D); /* tokens 1..np */
            for (int i = 0; i < D; i++)
                x[i] = cls[i];
            for (size_t i = 0; i < (size_t)N * D; i++)
                x[i] += pos[i];

            for (int l = 0; l < c->depth; l++)
            {
                oracle_layernorm(x, N, D, L[l].ln1_g, L[l].ln1_b, 1e-6f, y);
                linear_rows(y, 0, N, D, L[l].qkv_wt, L[l].qkv_b, 3 * D, qkv);
                attention_one(qkv, N, c->heads, hd, att, scores);
                linear_rows(att, 0, N, D, L[l].proj_wt, L[l].proj_b, D, tmp);
                for (size_t i = 0; i < (size_t)N * D; i++)
                    x[i] += tmp[i];
                oracle_layernorm(x, N, D, L[l].ln2_g, L[l].ln2_b, 1e-6f, y);
                linear_rows(y, 0, N, D, L[l].fc1_wt, L[l].fc1_b, F, hid);
                for (size_t i = 0; i < (size_t)N * F; i++)
                    hid[i] = oracle_gelu(hid[i]);
                linear_rows(hid, 0, N, F, L[l].fc2_wt, L[l].fc2_b, D, tmp);
                for (size_t i = 0; i < (size_t)N * D; i++)
                    x[i] += tmp[i];
            }
            oracle_layernorm(x, 1, D, lnf_g, lnf_b, 1e-6f, y); /* only the cls token feeds the head */
            linear_rows(y, 0, 1, D, head_wt, head_b, C, logits + (size_t)b * C);
        }
        free(patches), free(x), free(y), free(qkv), free(att), free(hid), free(tmp), free(scores);
    }
    for (int l = 0; l < c->depth; l++)
        free(L[l].qkv_wt), free(L[l].proj_wt), free(L[l].fc1_wt), free(L[l].fc2_wt);
    free(L), free(patch_wt), free(head_wt);
    return 0;
}

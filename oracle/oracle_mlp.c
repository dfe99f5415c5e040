/* oracle_mlp.c -- CPU restatement of the MLP forward (`network_v1`).  TEST INFRASTRUCTURE ONLY;
 * see oracle.h for the parity status ("parity unpinned" for the arithmetic, layout pinned).
 *
 * What is taken from the reference (file:line in /root/reference):
 *   - flat weight layout: layer-major, then output neuron, then input index  src/netFPGA.cpp:91-106
 *   - sizes: fan-in of layer 0 is n_ins, of layer i is npl[i-1]              src/netFPGA.cpp:64-76
 *   - kernel argument list (inputs, params, bias, outs, npl, n_layers, n_ins) :427-436, :499-502
 *   - one sample in, npl[last] floats out                                    src/netFPGA.cpp:266-289
 *   - random init rule                                                       src/netFPGA.cpp:82-88
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int layer_is_relu(int layer, int n_layers, int activation)
{
    if (activation == ORACLE_ACT_NONE)
        return 0;
    if (activation == ORACLE_ACT_RELU_ALL)
        return 1;
    return layer != n_layers - 1; /* RELU_HIDDEN */
}

/* The scalar statement of the arithmetic: this function *is* the specification. */
void oracle_mlp_forward_one(const float *inputs, const float *params, const float *bias, float *outs,
                            const int *npl, int n_layers, int n_ins, int activation)
{
    int widest = n_ins;
    for (int l = 0; l < n_layers; l++)
        if (npl[l] > widest)
            widest = npl[l];
    float *cur = (float *)malloc(sizeof(float) * (size_t)widest);
    float *nxt = (float *)malloc(sizeof(float) * (size_t)widest);
    memcpy(cur, inputs, sizeof(float) * (size_t)n_ins);

    const float *w = params;
    const float *b = bias;
    int fan_in = n_ins;
    for (int l = 0; l < n_layers; l++)
    {
        const int fan_out = npl[l];
        const int relu = layer_is_relu(l, n_layers, activation);
        for (int j = 0; j < fan_out; j++)
        {
            float acc = b[j];
            const float *row = w + (size_t)j * fan_in;
            for (int k = 0; k < fan_in; k++)
                acc = fmaf(row[k], cur[k], acc);
            nxt[j] = (relu && acc < 0.0f) ? 0.0f : acc;
        }
        w += (size_t)fan_out * fan_in;
        b += fan_out;
        fan_in = fan_out;
        float *t = cur;
        cur = nxt;
        nxt = t;
    }
    memcpy(outs, cur, sizeof(float) * (size_t)fan_in);
    free(cur);
    free(nxt);
}

/* Batched form.  Uses oracle_linear (blocked, vectorisable) whose per-element operation sequence
 * is identical to the scalar loop above: acc = bias, then fmaf in ascending k.  tests/ assert the
 * two are bit-identical. */
void oracle_mlp_forward(const float *in, size_t batch, const float *params, const float *bias, float *out,
                        const int *npl, int n_layers, int n_ins, int activation, int threads)
{
    int widest = n_ins;
    for (int l = 0; l < n_layers; l++)
        if (npl[l] > widest)
            widest = npl[l];
    float *cur = (float *)malloc(sizeof(float) * batch * (size_t)widest);
    float *nxt = (float *)malloc(sizeof(float) * batch * (size_t)widest);
    memcpy(cur, in, sizeof(float) * batch * (size_t)n_ins);

    const float *w = params;
    const float *b = bias;
    int fan_in = n_ins;
    for (int l = 0; l < n_layers; l++)
    {
        const int fan_out = npl[l];
        oracle_linear(cur, batch, fan_in, w, b, fan_out, nxt, threads);
        if (layer_is_relu(l, n_layers, activation))
        {
            const size_t n = batch * (size_t)fan_out;
            for (size_t i = 0; i < n; i++)
                if (nxt[i] < 0.0f)
                    nxt[i] = 0.0f;
        }
        w += (size_t)fan_out * fan_in;
        b += fan_out;
        fan_in = fan_out;
        float *t = cur;
        cur = nxt;
        nxt = t;
    }
    memcpy(out, cur, sizeof(float) * batch * (size_t)fan_in);
    free(cur);
    free(nxt);
}

void oracle_rand_init(unsigned seed, float *params, size_t n_params, float *bias, size_t n_neurons)
{
    srand(seed);
    for (size_t i = 0; i < n_params; i++)
        params[i] = (float)(rand() % 200 - 100) / 100;
    for (size_t i = 0; i < n_neurons; i++)
        bias[i] = (float)(rand() % 200 - 100) / 100;
}

/* ---- INT8 Q1.7 --------------------------------------------------------------------------- */

void oracle_quantize_q17(const float *x, size_t n, int8_t *q)
{
    for (size_t i = 0; i < n; i++)
    {
        float v = rintf(x[i] * 128.0f); /* ties to even, default rounding mode */
        if (v > 127.0f)
            v = 127.0f;
        if (v < -128.0f)
            v = -128.0f;
        q[i] = (int8_t)v;
    }
}

void oracle_quantize_bias_q214(const float *b, size_t n, int32_t *q)
{
    for (size_t i = 0; i < n; i++)
        q[i] = (int32_t)rintf(b[i] * 16384.0f);
}

void oracle_mlp_forward_i8(const int8_t *in, size_t batch, const int8_t *params, const int32_t *bias,
                           int32_t *out, const int *npl, int n_layers, int n_ins, int activation, int threads)
{
    int widest = n_ins;
    for (int l = 0; l < n_layers; l++)
        if (npl[l] > widest)
            widest = npl[l];
    const int n_out = npl[n_layers - 1];
    (void)threads;
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
        int8_t *cur = (int8_t *)malloc((size_t)widest);
        int8_t *nxt = (int8_t *)malloc((size_t)widest);
#pragma omp for schedule(static)
        for (long long s = 0; s < (long long)batch; s++)
        {
            memcpy(cur, in + (size_t)s * n_ins, (size_t)n_ins);
            const int8_t *w = params;
            const int32_t *b = bias;
            int fan_in = n_ins;
            for (int l = 0; l < n_layers; l++)
            {
                const int fan_out = npl[l];
                const int relu = layer_is_relu(l, n_layers, activation);
                const int last = (l == n_layers - 1);
                for (int j = 0; j < fan_out; j++)
                {
                    int32_t acc = b[j];
                    const int8_t *row = w + (size_t)j * fan_in;
                    for (int k = 0; k < fan_in; k++)
                        acc += (int32_t)row[k] * (int32_t)cur[k];
                    if (relu && acc < 0)
                        acc = 0;
                    if (last)
                        out[(size_t)s * n_out + j] = acc;
                    else
                    {
                        int32_t q = acc >> 7; /* arithmetic shift (gcc/nvcc: sign-propagating) */
                        if (q > 127)
                            q = 127;
                        if (q < -128)
                            q = -128;
                        nxt[j] = (int8_t)q;
                    }
                }
                w += (size_t)fan_out * fan_in;
                b += fan_out;
                fan_in = fan_out;
                int8_t *t = cur;
                cur = nxt;
                nxt = t;
            }
        }
        free(cur);
        free(nxt);
    }
}

// runtime.cu -- host runtime behind the C ABI of include/netcuda.h.
//
// Replaces the OpenCL plumbing of the reference's src/netFPGA.cpp: device discovery and context
// (_init_program :367-400), buffer creation (_init_kernel :402-441), weight upload (_load_params
// :484-515), the per-sample write/task/read triple (launch_forward :266-277) and teardown
// (cleanup :639-651).  Differences by design: state is per handle (the reference uses namespace
// globals, :21-56), inputs are batched, errors are returned instead of exit()ing, and the device
// side is a sequence of sm_100a kernels (csrc/*.cu) instead of one FPGA task.
//
// HBM layout per handle
//   weights  one arena; every tensor 256-byte aligned.  MLP: per layer W[out][ld(in)] in the operand
//            type (fp32 / bf16 / int8; ld = fan-in rounded up to 16 bytes) + bias (fp32 or int32).
//            ViT: bf16 W[out][in] for the six GEMMs per block + patch/head, fp32 for bias/LN/cls/pos.
//   work     activations of one pass of `max_batch` samples (MLP: two ping-pong matrices;
//            ViT: patches, fp32 residual stream x, bf16 LN-out / qkv / attention-out / MLP-hidden).
//   staging  (host API only, lazily) 2 pinned host + 2 device input slots and a device/pinned output.
#include "../../include/netcuda.h"
#include "gemm_tcgen05.cuh"
#include "kernels.h"

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace nc;

// ---- error reporting -----------------------------------------------------------------------------

static thread_local char g_last_error[512] = "";

int nc::set_last_error_v(int code, const char *fmt, va_list ap)
{
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    return code;
}

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    nc::set_last_error_v(code, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(expr)                                                                                              \
    do                                                                                                        \
    {                                                                                                         \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return fail(NETCUDA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

static inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

// INT8 layers at small batch are weight streaming: up to this many samples a layer is split along K over the whole GPU
constexpr int SPLITK_MAX_M = 128;
constexpr size_t SMALL_CALL_BYTES = 256 << 10; // host calls of an MLP whose inputs and outputs are at most this large take the single-stream path
constexpr int GRAPH_MAX_SAMPLES = 4096;   // MLP passes up to this many samples are launch-latency bound (graph replay, PDL)
constexpr int VIT_GRAPH_MAX_ROWS = 8192; // ViT passes up to this many token rows are launch-latency bound (graph replay, PDL)

// ---- handle --------------------------------------------------------------------------------------

struct MlpLayer
{
    int fan_in, fan_out;
    long long ldw;   // elements
    void *w;         // operand type
    void *bias;      // float or int32
    int8_t *w_tiled = nullptr; // INT8 nets the cluster streaming kernel can serve: the same weights in 16 KB streaming blocks
};

struct VitBlock
{
    float *ln1_g, *ln1_b, *qkv_b, *proj_b, *ln2_g, *ln2_b, *fc1_b, *fc2_b;
    void *qkv_w, *proj_w, *fc1_w, *fc2_w; // operand type: bf16, or fp32 (tf32 nets)
};

struct netcuda_net
{
    netcuda_desc desc;
    std::vector<int> npl;
    int device = 0, num_sms = 148;
    int max_batch = 0;
    int gemm_variant = 0;
    bool use_graphs = true; // NETCUDA_GRAPHS=0 disables the CUDA-graph replay of small MLP passes
    bool weights_loaded = false;
    size_t n_in = 0, n_out = 0;
    double flops_per_sample = 0;
    uint64_t launches = 0;
    int64_t last_us = 0;

    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t h2d_done[2] = {nullptr, nullptr}, compute_done[2] = {nullptr, nullptr};

    int *h_err = nullptr; // mapped pinned: readable even after a device-side trap
    int *d_err = nullptr;

    // weights
    char *arena = nullptr;
    size_t arena_bytes = 0, arena_used = 0;
    std::vector<MlpLayer> layers;
    int elem = 4;        // operand element size (MLP)
    long long max_ld = 0; // widest padded activation row (MLP)
    // ViT
    int vit_kind = GK_BF16; // operand kind of the linear layers: GK_BF16 or GK_TF32 (attention core: bf16 operands either way)
    int T = 0, NP = 0, PK = 0;
    void *patch_w = nullptr, *head_w = nullptr;
    float *patch_b = nullptr, *cls = nullptr, *pos = nullptr, *lnf_g = nullptr, *lnf_b = nullptr, *head_b = nullptr;
    std::vector<VitBlock> blocks;

    // work buffers
    void *act[2] = {nullptr, nullptr};     // MLP ping-pong
    int32_t *acc_out = nullptr;            // INT8 float API: last layer accumulators
    int32_t *splitk_ws = nullptr;          // INT8 small batch: [128][widest layer] int32 partial sums, all zero between layers
    int stream_pf_tiles = 3;               // ... and how many weight tiles it prefetches into L2 ahead of its ring (NETCUDA_MLP_STREAM_PF)
    void *stream_ll[2] = {nullptr, nullptr}; // ... and its tagged-word activation buffers (<= 4 samples: no grid barrier)
    void *w_tiled_base = nullptr;          // INT8: the tiled weight copy of every layer (retile_int8_weights)
    unsigned *stream_bar = nullptr;        // INT8, <= 32 samples: the two counters of the weight-streaming kernel's grid barrier
    bool use_stream = true;                // NETCUDA_MLP_STREAM=0 keeps such batches on the split-K GEMM path
    int stream_max_batch = 16;                   // up to here the mma.sync streaming kernel (it can serve 32: NETCUDA_MLP_STREAM_SPLIT), above it (<= 128) the tcgen05 one
                                                 // (NETCUDA_MLP_STREAM_SPLIT: A/B of the hand-over point)
    int umma_min_batch = 17;                     // ... from here on (NETCUDA_MLP_UMMA_MIN moves the hand-over, for A/B runs against the split-K path)
    int umma_pair = 1;                           // the tcgen05 streaming kernel as split-K CTA clusters where the net allows it: 1 = clusters of four up to
                                                 // 88 samples, pairs above; 2 = pairs, four issuers; 3 = pairs, two issuers; 4 = clusters of four at every
                                                 // batch; 0 = single CTAs (NETCUDA_MLP_UMMA_PAIR, for A/B runs)
    void *patches = nullptr, *ybuf = nullptr, *qkv = nullptr, *att = nullptr, *hid = nullptr, *cls_ln = nullptr;
    float *x = nullptr;

    // CUDA graph of the last small MLP pass (launch-bound regime): replayed while the buffers, the batch and the variant repeat
    struct PassGraph
    {
        cudaGraphExec_t exec = nullptr;
        const void *in = nullptr;
        void *out = nullptr;
        int n = 0, variant = 0;
        bool in_is_i8 = false, out_is_i32 = false;
        uint64_t launches = 0; // kernels inside the graph (for the launch counter)
    };
    static constexpr int PASS_GRAPHS = 40; // the host API cycles through 2 input slots x up to 4 output buffers; a device-resident call of
                                           // several small passes (NETCUDA_VIT_GRAPH_MULTIPASS) keeps one graph per pass
    PassGraph pass_graphs[PASS_GRAPHS];
    PassGraph graph_candidates[PASS_GRAPHS]; // ViT: a (buffers, batch) combination is captured the second time it shows up
    int graph_candidate_next = 0;
    int pass_graph_next = 0; // round-robin replacement

    // per-kernel profiling (netcuda_profile_enable)
    bool profiling = false;
    std::vector<std::string> prof_labels;
    struct ProfRec
    {
        int label;
        cudaEvent_t start, stop;
        double flops, bytes;
    };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_free;

    // host-API staging (lazy)
    void *pin_in[2] = {nullptr, nullptr}, *dev_in[2] = {nullptr, nullptr};
    size_t stage_in_bytes = 0;
    // u8 frame input (netcuda_forward_u8): v = (u8 / 255 - mean) * inv_std; the default maps 0..255 onto [-1, 1], the input range
    // the reference names (MIN_RANGE / MAX_RANGE, def/defines.h:11-12)
    float u8_mean[3] = {0.5f, 0.5f, 0.5f}, u8_inv_std[3] = {2.0f, 2.0f, 2.0f};
    uint64_t chunk_seq = 0; // passes staged so far: slot = chunk_seq & 1, across calls
    // Host-API calls in flight (netcuda_submit / netcuda_wait); netcuda_forward is submit + wait.
    struct Pending
    {
        uint64_t ticket = 0;
        bool active = false;
        int status = NETCUDA_OK; // of a retired ticket
        cudaEvent_t done = nullptr;
        void *dev_out = nullptr, *pin_out = nullptr;
        void *small_in = nullptr, *small_dev_in = nullptr; // page-locked / device input of a small call (SMALL_CALL_BYTES), see submit_host_impl
        size_t dev_cap = 0, pin_cap = 0;
        void *user_out = nullptr;
        size_t out_bytes = 0;
        bool pinned_out = false;
        std::chrono::steady_clock::time_point t0;
    };
    static constexpr int MAX_IN_FLIGHT = 4;
    Pending pending[MAX_IN_FLIGHT];
    uint64_t next_ticket = 1;
};

static int check_handle(const netcuda_net *h)
{
    if (!h) return fail(NETCUDA_ERR_INVALID, "null handle");
    return NETCUDA_OK;
}

// Every kernel launch of a forward pass goes through one KernelScope: it counts the launch and, when
// profiling is on, brackets it with a CUDA-event pair on the launching stream.
struct KernelScope
{
    cudaStream_t s;
    cudaEvent_t stop = nullptr;
    KernelScope(netcuda_net *h, cudaStream_t stream, const char *label, double flops, double bytes) : s(stream)
    {
        h->launches++;
        if (!h->profiling) return;
        int li = -1;
        for (size_t i = 0; i < h->prof_labels.size(); i++)
            if (h->prof_labels[i] == label) li = (int)i;
        if (li < 0)
        {
            h->prof_labels.push_back(label);
            li = (int)h->prof_labels.size() - 1;
        }
        cudaEvent_t ev[2] = {nullptr, nullptr};
        for (int i = 0; i < 2; i++)
        {
            if (!h->prof_free.empty())
            {
                ev[i] = h->prof_free.back();
                h->prof_free.pop_back();
            }
            else if (cudaEventCreate(&ev[i]) != cudaSuccess)
            {
                (void)cudaGetLastError();
                if (ev[0]) h->prof_free.push_back(ev[0]);
                return;
            }
        }
        cudaEventRecord(ev[0], s);
        stop = ev[1];
        h->prof_recs.push_back({li, ev[0], ev[1], flops, bytes});
    }
    ~KernelScope()
    {
        if (stop) cudaEventRecord(stop, s);
    }
    KernelScope(const KernelScope &) = delete;
    KernelScope &operator=(const KernelScope &) = delete;
};

static void *arena_take(netcuda_net *h, size_t bytes)
{
    const size_t off = (size_t)round_up((long long)h->arena_used, 256);
    if (off + bytes > h->arena_bytes) return nullptr;
    h->arena_used = off + bytes;
    return h->arena + off;
}

static int operand_kind(int precision)
{
    switch (precision)
    {
    case NETCUDA_PREC_FP32: return GK_FP32_SIMT;
    case NETCUDA_PREC_TF32: return GK_TF32;
    case NETCUDA_PREC_BF16: return GK_BF16;
    default: return GK_I8;
    }
}

static size_t vit_param_count(const netcuda_desc *d)
{
    const size_t D = d->dim, F = d->mlp_dim, C = d->n_classes;
    const size_t g = d->image_size / d->patch_size, N = g * g + 1, pk = 3u * d->patch_size * d->patch_size;
    size_t n = D * pk + D + D + N * D;
    n += (size_t)d->depth * (2 * D + 3 * D * D + 3 * D + D * D + D + 2 * D + F * D + F + D * F + D);
    n += 2 * D + C * D + C;
    return n;
}

static int validate_desc(const netcuda_desc *d)
{
    if (!d) return fail(NETCUDA_ERR_INVALID, "null descriptor");
    if (d->precision < NETCUDA_PREC_FP32 || d->precision > NETCUDA_PREC_INT8)
        return fail(NETCUDA_ERR_INVALID, "unknown precision %d", d->precision);
    if (d->kind == NETCUDA_KIND_MLP)
    {
        if (d->n_ins <= 0 || d->n_layers <= 0 || !d->n_p_l) return fail(NETCUDA_ERR_INVALID, "MLP needs n_ins, n_layers, n_p_l");
        for (int i = 0; i < d->n_layers; i++)
            if (d->n_p_l[i] <= 0) return fail(NETCUDA_ERR_INVALID, "n_p_l[%d] must be positive", i);
        if (d->activation < NETCUDA_ACT_RELU_HIDDEN || d->activation > NETCUDA_ACT_NONE)
            return fail(NETCUDA_ERR_INVALID, "unknown activation %d", d->activation);
        return NETCUDA_OK;
    }
    if (d->kind == NETCUDA_KIND_VIT)
    {
        if (d->precision != NETCUDA_PREC_BF16 && d->precision != NETCUDA_PREC_TF32)
            return fail(NETCUDA_ERR_UNSUPPORTED, "ViT nets run in NETCUDA_PREC_BF16 or NETCUDA_PREC_TF32 (got precision %d)", d->precision);
        if (d->image_size <= 0 || d->patch_size <= 0 || d->image_size % d->patch_size || d->patch_size % 8)
            return fail(NETCUDA_ERR_INVALID, "image_size must be a multiple of patch_size, patch_size a multiple of 8");
        if (d->dim <= 0 || d->heads <= 0 || d->dim != d->heads * 64)
            return fail(NETCUDA_ERR_UNSUPPORTED, "dim must equal heads * 64 (head_dim 64 only)");
        if (d->depth <= 0 || d->mlp_dim <= 0 || d->n_classes <= 0 || d->dim % 8 || d->mlp_dim % 8)
            return fail(NETCUDA_ERR_INVALID, "bad ViT dimensions");
        return NETCUDA_OK;
    }
    return fail(NETCUDA_ERR_INVALID, "unknown net kind %d", d->kind);
}

// ---- C ABI: life cycle -----------------------------------------------------------------------------

extern "C" int netcuda_abi_version(void) { return NETCUDA_ABI_VERSION; }
extern "C" const char *netcuda_last_error(void) { return g_last_error; }

extern "C" int netcuda_device_count(int *count)
{
    if (!count) return fail(NETCUDA_ERR_INVALID, "null count");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess)
    {
        *count = 0;
        (void)cudaGetLastError();
        return fail(NETCUDA_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return NETCUDA_OK;
}

extern "C" int netcuda_destroy(netcuda_net *h)
{
    if (!h) return NETCUDA_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    void *dev_ptrs[] = {h->arena, h->act[0], h->act[1], h->acc_out, h->splitk_ws, h->w_tiled_base, h->stream_bar, h->stream_ll[0], h->stream_ll[1], h->patches, h->ybuf, h->qkv, h->att, h->hid, h->cls_ln, h->x,
                        h->dev_in[0], h->dev_in[1]};
    for (void *p : dev_ptrs)
        if (p) cudaFree(p);
    void *host_ptrs[] = {h->pin_in[0], h->pin_in[1], h->h_err};
    for (void *p : host_ptrs)
        if (p) cudaFreeHost(p);
    for (int i = 0; i < 2; i++)
    {
        if (h->h2d_done[i]) cudaEventDestroy(h->h2d_done[i]);
        if (h->compute_done[i]) cudaEventDestroy(h->compute_done[i]);
    }
    for (auto &g : h->pass_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto &pd : h->pending)
    {
        if (pd.done) cudaEventDestroy(pd.done);
        if (pd.dev_out) cudaFree(pd.dev_out);
        if (pd.pin_out) cudaFreeHost(pd.pin_out);
        if (pd.small_in) cudaFreeHost(pd.small_in);
        if (pd.small_dev_in) cudaFree(pd.small_dev_in);
    }
    for (auto &r : h->prof_recs)
    {
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    for (cudaEvent_t e : h->prof_free) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    (void)cudaGetLastError();
    delete h;
    return NETCUDA_OK;
}

static int create_impl(const netcuda_desc *desc, netcuda_net *h)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        (void)cudaGetLastError();
        return fail(NETCUDA_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU fallback)");
    }
    if (desc->device < 0 || desc->device >= ndev) return fail(NETCUDA_ERR_INVALID, "device %d out of range [0,%d)", desc->device, ndev);
    h->device = desc->device;
    CK(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    if (prop.major != 10)
        return fail(NETCUDA_ERR_NO_DEVICE, "device %d is sm_%d%d; this library contains sm_100a code only", h->device, prop.major, prop.minor);
    h->num_sms = prop.multiProcessorCount;
    CK(gemm_global_init());
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++)
    {
        CK(cudaEventCreateWithFlags(&h->h2d_done[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->compute_done[i], cudaEventDisableTiming));
    }
    CK(cudaHostAlloc((void **)&h->h_err, sizeof(int), cudaHostAllocMapped));
    *h->h_err = 0;
    CK(cudaHostGetDevicePointer((void **)&h->d_err, h->h_err, 0));

    h->desc = *desc;
    h->desc.n_p_l = nullptr;
    if (const char *env = getenv("NETCUDA_GRAPHS")) h->use_graphs = atoi(env) != 0;

    if (desc->kind == NETCUDA_KIND_MLP)
    {
        h->npl.assign(desc->n_p_l, desc->n_p_l + desc->n_layers);
        h->n_in = (size_t)desc->n_ins;
        h->n_out = (size_t)h->npl.back();
        h->max_batch = desc->max_batch > 0 ? desc->max_batch : 16384;
        const int kind = operand_kind(desc->precision);
        h->elem = kind == GK_BF16 ? 2 : kind == GK_I8 ? 1 : 4;
        const int pad = kind == GK_FP32_SIMT ? 1 : 16 / h->elem; // TMA: row pitch multiple of 16 bytes
        size_t bytes = 0;
        long long fan_in = desc->n_ins;
        h->max_ld = round_up(fan_in, pad);
        double macs = 0;
        for (int l = 0; l < desc->n_layers; l++)
        {
            MlpLayer L;
            L.fan_in = (int)fan_in, L.fan_out = h->npl[l];
            L.ldw = round_up(fan_in, pad);
            L.w = L.bias = nullptr;
            bytes += (size_t)round_up((long long)L.fan_out * L.ldw * h->elem, 256) + (size_t)round_up((long long)L.fan_out * 4, 256) + 512;
            macs += (double)fan_in * L.fan_out;
            h->layers.push_back(L);
            fan_in = L.fan_out;
            if (round_up(fan_in, pad) > h->max_ld) h->max_ld = round_up(fan_in, pad);
        }
        h->flops_per_sample = 2.0 * macs;
        h->arena_bytes = bytes;
        CK(cudaMalloc((void **)&h->arena, h->arena_bytes));
        CK(cudaMemset(h->arena, 0, h->arena_bytes));
        for (auto &L : h->layers)
        {
            L.w = arena_take(h, (size_t)L.fan_out * L.ldw * h->elem);
            L.bias = arena_take(h, (size_t)L.fan_out * 4);
            if (!L.w || !L.bias) return fail(NETCUDA_ERR_CUDA, "weight arena overflow");
        }
        const size_t act_bytes = (size_t)h->max_batch * h->max_ld * h->elem;
        CK(cudaMalloc(&h->act[0], act_bytes));
        CK(cudaMalloc(&h->act[1], act_bytes));
        CK(cudaMemset(h->act[0], 0, act_bytes));
        CK(cudaMemset(h->act[1], 0, act_bytes));
        if (desc->precision == NETCUDA_PREC_INT8)
        {
            CK(cudaMalloc((void **)&h->acc_out, (size_t)h->max_batch * h->n_out * 4));
            int widest = 0;
            for (auto &L : h->layers) widest = std::max(widest, L.fan_out);
            CK(cudaMalloc((void **)&h->splitk_ws, (size_t)SPLITK_MAX_M * widest * 4));
            CK(cudaMemset(h->splitk_ws, 0, (size_t)SPLITK_MAX_M * widest * 4));
            CK(cudaMalloc((void **)&h->stream_bar, 4 * sizeof(unsigned)));
            CK(cudaMemset(h->stream_bar, 0, 4 * sizeof(unsigned)));
            // tagged-word activation buffers of the weight-streaming kernel (<= 4 samples): 2 bytes per activation, two layers
            if (const char *e = getenv("NETCUDA_MLP_STREAM_PF")) h->stream_pf_tiles = std::min(std::max(atoi(e), 0), 64);
            const char *ll_env = getenv("NETCUDA_MLP_STREAM_LL"); // (=0: grid-barrier exchange at every batch size, for A/B runs)
            if (widest <= 4096 && !(ll_env && atoi(ll_env) == 0))
            {
                const size_t ll_bytes = (size_t)MLP_STREAM_LL_MAX_BATCH * widest * 2;
                for (int i = 0; i < 2; i++)
                {
                    CK(cudaMalloc(&h->stream_ll[i], ll_bytes));
                    CK(cudaMemset(h->stream_ll[i], 0, ll_bytes));
                }
            }
            if (const char *e = getenv("NETCUDA_MLP_STREAM")) h->use_stream = atoi(e) != 0;
            if (const char *e = getenv("NETCUDA_MLP_STREAM_SPLIT")) h->stream_max_batch = std::min(std::max(atoi(e), 0), MLP_STREAM_MAX_BATCH);
            if (const char *e = getenv("NETCUDA_MLP_UMMA_MIN")) h->umma_min_batch = std::max(atoi(e), 1);
            if (const char *e = getenv("NETCUDA_MLP_UMMA_PAIR")) h->umma_pair = std::min(std::max(atoi(e), 0), 4);
        }
    }
    else
    {
        const int D = desc->dim, F = desc->mlp_dim, C = desc->n_classes, P = desc->patch_size;
        const int g = desc->image_size / P;
        h->NP = g * g, h->T = h->NP + 1, h->PK = 3 * P * P;
        h->n_in = (size_t)3 * desc->image_size * desc->image_size;
        h->n_out = (size_t)C;
        h->max_batch = desc->max_batch > 0 ? desc->max_batch : 256;
        const double N = h->T;
        // dense MACs per block: qkv 3D^2 + proj D^2 + fc1 D*F + fc2 F*D per token, attention 2*N*D per token
        h->flops_per_sample = 2.0 * ((double)h->NP * h->PK * D +
                                     desc->depth * (N * (4.0 * D * D + 2.0 * D * F) + 2.0 * N * N * D) + (double)D * C);
        const size_t nparams = vit_param_count(desc);
        h->arena_bytes = nparams * 4 + (size_t)(desc->depth * 12 + 16) * 256; // generous: every tensor as fp32 + alignment
        CK(cudaMalloc((void **)&h->arena, h->arena_bytes));
        auto take = [&](size_t bytes) { return arena_take(h, bytes); };
        // TF32 nets (the reference's DATA_TYPE is float, def/defines.h:10): fp32 weights and activations feed kind::tf32 MMAs in every
        // linear layer; only the attention core keeps bf16 operands (q, k, v and P are rounded to bf16, fp32 accumulate).
        h->vit_kind = desc->precision == NETCUDA_PREC_TF32 ? GK_TF32 : GK_BF16;
        const size_t es = h->vit_kind == GK_TF32 ? 4 : 2;
        h->patch_w = take((size_t)D * h->PK * es);
        h->patch_b = (float *)take((size_t)D * 4);
        h->cls = (float *)take((size_t)D * 4);
        h->pos = (float *)take((size_t)h->T * D * 4);
        h->blocks.resize(desc->depth);
        for (auto &b : h->blocks)
        {
            b.ln1_g = (float *)take((size_t)D * 4), b.ln1_b = (float *)take((size_t)D * 4);
            b.qkv_w = take((size_t)3 * D * D * es), b.qkv_b = (float *)take((size_t)3 * D * 4);
            b.proj_w = take((size_t)D * D * es), b.proj_b = (float *)take((size_t)D * 4);
            b.ln2_g = (float *)take((size_t)D * 4), b.ln2_b = (float *)take((size_t)D * 4);
            b.fc1_w = take((size_t)F * D * es), b.fc1_b = (float *)take((size_t)F * 4);
            b.fc2_w = take((size_t)D * F * es), b.fc2_b = (float *)take((size_t)D * 4);
        }
        h->lnf_g = (float *)take((size_t)D * 4), h->lnf_b = (float *)take((size_t)D * 4);
        h->head_w = take((size_t)C * D * es), h->head_b = (float *)take((size_t)C * 4);
        if (!h->head_b) return fail(NETCUDA_ERR_CUDA, "weight arena overflow");

        const size_t mb = (size_t)h->max_batch, rows = mb * h->T;
        CK(cudaMalloc(&h->patches, mb * h->NP * h->PK * es));
        CK(cudaMalloc((void **)&h->x, rows * D * 4));
        CK(cudaMalloc(&h->ybuf, rows * D * es));
        CK(cudaMalloc(&h->qkv, rows * 3 * D * 2)); // always bf16: the attention kernels' TMA source
        CK(cudaMalloc(&h->att, rows * D * es));
        CK(cudaMalloc(&h->hid, rows * F * es));
        CK(cudaMalloc(&h->cls_ln, mb * D * es));
    }
    return NETCUDA_OK;
}

extern "C" int netcuda_create(const netcuda_desc *desc, netcuda_net **out)
{
    if (!out) return fail(NETCUDA_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int rc = validate_desc(desc);
    if (rc != NETCUDA_OK) return rc;
    netcuda_net *h = new netcuda_net();
    rc = create_impl(desc, h);
    if (rc != NETCUDA_OK)
    {
        char keep[sizeof(g_last_error)];
        memcpy(keep, g_last_error, sizeof(keep));
        netcuda_destroy(h);
        memcpy(g_last_error, keep, sizeof(keep));
        return rc;
    }
    *out = h;
    return NETCUDA_OK;
}

// ---- weights ----------------------------------------------------------------------------------------

// fp32 host tensor [rows][cols] -> device tensor [rows][ld] in the operand type, via a device scratch copy.
static int upload_matrix(netcuda_net *h, const float *src, long long rows, int cols, long long ld, int kind, void *dst, float *scratch)
{
    CK(cudaMemcpyAsync(scratch, src, (size_t)rows * cols * 4, cudaMemcpyHostToDevice, h->stream));
    cudaError_t e;
    if (kind == GK_BF16)
        e = launch_convert_rows_bf16(scratch, dst, rows, cols, (int)ld, h->stream);
    else if (kind == GK_I8)
        e = launch_quantize_rows_q17(scratch, (int8_t *)dst, rows, cols, (int)ld, h->stream);
    else
        e = launch_convert_rows_f32(scratch, (float *)dst, rows, cols, (int)ld, h->stream);
    CK(e);
    CK(cudaStreamSynchronize(h->stream)); // scratch and the pageable source are reused by the caller
    return NETCUDA_OK;
}

// INT8 nets whose every layer the cluster streaming kernel can serve (fan-ins: multiples of 16 with at least two k-blocks) keep a
// second copy of their weights in that kernel's streaming layout: +1 byte per weight of HBM (config C5: 128 MiB), for 17..128-sample
// forwards that read contiguous 8..16 KB runs instead of 128-byte pieces 4 KB apart.  Rebuilt by every upload.
static int retile_int8_weights(netcuda_net *h)
{
    if (h->desc.precision != NETCUDA_PREC_INT8 || !h->use_stream || !h->umma_pair) return NETCUDA_OK;
    size_t total = 0;
    for (auto &L : h->layers)
    {
        if (L.ldw != L.fan_in || (L.fan_in & 15) || L.fan_in <= 128) return NETCUDA_OK; // the single-CTA kernels serve this net
        total += mlp_tiled_weight_bytes(L.fan_in, L.fan_out);
    }
    if (!h->w_tiled_base) CK(cudaMalloc(&h->w_tiled_base, total));
    size_t off = 0;
    for (auto &L : h->layers)
    {
        L.w_tiled = (int8_t *)h->w_tiled_base + off;
        CK(launch_retile_i8_weights((const int8_t *)L.w, L.ldw, L.fan_out, L.fan_in, L.w_tiled, h->stream));
        off += mlp_tiled_weight_bytes(L.fan_in, L.fan_out);
    }
    CK(cudaStreamSynchronize(h->stream));
    return NETCUDA_OK;
}

extern "C" int netcuda_upload_mlp(netcuda_net *h, const float *w_flat, const float *b_flat)
{
    if (int rc = check_handle(h)) return rc;
    if (h->desc.kind != NETCUDA_KIND_MLP) return fail(NETCUDA_ERR_INVALID, "not an MLP handle");
    if (!w_flat || !b_flat) return fail(NETCUDA_ERR_INVALID, "null weights");
    CK(cudaSetDevice(h->device));
    const int kind = operand_kind(h->desc.precision);
    size_t biggest = 0;
    for (auto &L : h->layers) biggest = std::max(biggest, (size_t)L.fan_in * L.fan_out);
    float *scratch = nullptr;
    CK(cudaMalloc((void **)&scratch, biggest * 4));
    int rc = NETCUDA_OK;
    for (auto &L : h->layers)
    {
        rc = upload_matrix(h, w_flat, L.fan_out, L.fan_in, L.ldw, kind, L.w, scratch);
        if (rc != NETCUDA_OK) break;
        if (kind == GK_I8)
        {
            // bias: Q2.14 int32, (int32) rintf(b * 16384) -- same expression as oracle_quantize_bias_q214
            std::vector<int32_t> q(L.fan_out);
            for (int j = 0; j < L.fan_out; j++) q[j] = (int32_t)rintf(b_flat[j] * 16384.0f);
            cudaError_t e = cudaMemcpy(L.bias, q.data(), (size_t)L.fan_out * 4, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { rc = fail(NETCUDA_ERR_CUDA, "bias upload: %s", cudaGetErrorString(e)); break; }
        }
        else
        {
            cudaError_t e = cudaMemcpy(L.bias, b_flat, (size_t)L.fan_out * 4, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { rc = fail(NETCUDA_ERR_CUDA, "bias upload: %s", cudaGetErrorString(e)); break; }
        }
        w_flat += (size_t)L.fan_in * L.fan_out;
        b_flat += L.fan_out;
    }
    cudaFree(scratch);
    if (rc == NETCUDA_OK) rc = retile_int8_weights(h);
    if (rc == NETCUDA_OK) h->weights_loaded = true;
    return rc;
}

extern "C" int netcuda_upload_mlp_i8(netcuda_net *h, const int8_t *w_flat, const int32_t *b_flat)
{
    if (int rc = check_handle(h)) return rc;
    if (h->desc.kind != NETCUDA_KIND_MLP || h->desc.precision != NETCUDA_PREC_INT8)
        return fail(NETCUDA_ERR_INVALID, "netcuda_upload_mlp_i8 needs an INT8 MLP handle");
    if (!w_flat || !b_flat) return fail(NETCUDA_ERR_INVALID, "null weights");
    CK(cudaSetDevice(h->device));
    for (auto &L : h->layers)
    {
        CK(cudaMemcpy2D(L.w, (size_t)L.ldw, w_flat, (size_t)L.fan_in, (size_t)L.fan_in, (size_t)L.fan_out, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(L.bias, b_flat, (size_t)L.fan_out * 4, cudaMemcpyHostToDevice));
        w_flat += (size_t)L.fan_in * L.fan_out;
        b_flat += L.fan_out;
    }
    if (int rc = retile_int8_weights(h)) return rc;
    h->weights_loaded = true;
    return NETCUDA_OK;
}

extern "C" int netcuda_vit_param_count(const netcuda_desc *desc, size_t *count)
{
    if (!desc || !count) return fail(NETCUDA_ERR_INVALID, "null argument");
    if (desc->kind != NETCUDA_KIND_VIT || desc->patch_size <= 0) return fail(NETCUDA_ERR_INVALID, "not a ViT descriptor");
    *count = vit_param_count(desc);
    return NETCUDA_OK;
}

extern "C" int netcuda_upload_vit(netcuda_net *h, const float *flat, size_t count)
{
    if (int rc = check_handle(h)) return rc;
    if (h->desc.kind != NETCUDA_KIND_VIT) return fail(NETCUDA_ERR_INVALID, "not a ViT handle");
    if (!flat || count != vit_param_count(&h->desc))
        return fail(NETCUDA_ERR_INVALID, "ViT parameter count mismatch: got %zu, expected %zu", count, vit_param_count(&h->desc));
    CK(cudaSetDevice(h->device));
    const int D = h->desc.dim, F = h->desc.mlp_dim, C = h->desc.n_classes;
    size_t biggest = std::max((size_t)D * h->PK, std::max((size_t)3 * D * D, std::max((size_t)F * D, (size_t)C * D)));
    float *scratch = nullptr;
    CK(cudaMalloc((void **)&scratch, biggest * 4));
    const float *p = flat;
    int rc = NETCUDA_OK;
    auto mat = [&](void *dst, long long rows, int cols) {
        if (rc == NETCUDA_OK) rc = upload_matrix(h, p, rows, cols, cols, h->vit_kind, dst, scratch);
        p += (size_t)rows * cols;
    };
    auto vec = [&](float *dst, size_t n) {
        if (rc == NETCUDA_OK && cudaMemcpy(dst, p, n * 4, cudaMemcpyHostToDevice) != cudaSuccess)
            rc = fail(NETCUDA_ERR_CUDA, "ViT vector upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        p += n;
    };
    mat(h->patch_w, D, h->PK);
    vec(h->patch_b, D);
    vec(h->cls, D);
    vec(h->pos, (size_t)h->T * D);
    for (auto &b : h->blocks)
    {
        vec(b.ln1_g, D), vec(b.ln1_b, D);
        mat(b.qkv_w, 3LL * D, D), vec(b.qkv_b, 3 * (size_t)D);
        mat(b.proj_w, D, D), vec(b.proj_b, D);
        vec(b.ln2_g, D), vec(b.ln2_b, D);
        mat(b.fc1_w, F, D), vec(b.fc1_b, F);
        mat(b.fc2_w, D, F), vec(b.fc2_b, D);
    }
    vec(h->lnf_g, D), vec(h->lnf_b, D);
    mat(h->head_w, C, D), vec(h->head_b, C);
    cudaFree(scratch);
    if (rc == NETCUDA_OK) h->weights_loaded = true;
    return rc;
}

// ---- forward: one pass over <= max_batch samples, everything on `s` -----------------------------------

static int out_elem_size(int out_type) { return out_type == OUT_BF16 ? 2 : out_type == OUT_S8 ? 1 : 4; }

static cudaError_t run_gemm(netcuda_net *h, const char *label, int kind, const void *a, long long lda, int a_rows, const void *w,
                            long long ldw, const void *bias, void *out, long long ldc, int out_type, int epi, int m, int n, int k,
                            cudaStream_t s, int remap_in = 0, int remap_out = 0, const float *pos = nullptr, int k_splits = 1)
{
    GemmCall c;
    c.k_splits = k_splits;
    c.kind = kind, c.variant = h->gemm_variant;
    c.a = a, c.lda = lda, c.a_rows = a_rows, c.w = w, c.ldw = ldw, c.bias = bias;
    c.out = out, c.ldc = ldc, c.out_type = out_type, c.epi = epi;
    c.m = m, c.n = n, c.k = k;
    c.remap_in = remap_in, c.remap_out = remap_out, c.pos = pos;
    c.error_flag = h->d_err, c.num_sms = h->num_sms;
    const double es = kind == GK_BF16 ? 2 : kind == GK_I8 ? 1 : 4;
    double bytes = ((double)m * k + (double)n * k) * es + (double)m * n * out_elem_size(out_type) + (double)n * 4;
    if (epi == EPI_RESIDUAL) bytes += (double)m * n * 4;
    if (epi == EPI_PATCH) bytes += (double)remap_in * n * 4;
    KernelScope scope(h, s, label, 2.0 * m * n * k, bytes);
    return launch_gemm(c, s);
}

// INT8 nets, up to 32 samples: the whole forward is one persistent weight-streaming kernel (mlp_stream.cu) when every layer's
// fan-in is a multiple of 16 bytes (no row padding anywhere) and the per-CTA output slices fit its shared memory.
// (n <= 16: the register-resident mma.sync kernel; 17..128: the tcgen05 kernel of mlp_umma_stream.cu -- same parameter block.  Measured
//  on config C5, us per forward: mma.sync stream 36 / 37 / 50 / 73 / 83 at 1 / 8 / 16 / 17 / 32 samples; tcgen05 stream with two issuing
//  threads 66 / 67 / 68 / 70 at 1 / 17 / 64 / 128 (one issuing thread: 105 / 107 / 110 at 33 / 64 / 128); split-K GEMM graph 98 / 108 / 125
//  at 33 / 64 / 128; the tcgen05 stream as split-K clusters, mlp_i8_umma_cluster_kernel: 49 / 52 / 59 at 17 / 64 / 128.)
static bool mlp_stream_params(netcuda_net *h, int n, const int8_t *in, int32_t *out, MlpStreamParams &p)
{
    if (h->desc.kind != NETCUDA_KIND_MLP || h->desc.precision != NETCUDA_PREC_INT8 || !h->use_stream || h->gemm_variant != 0 || !h->stream_bar)
        return false;
    const int L = (int)h->layers.size();
    if (n < 1 || n > MLP_UMMA_STREAM_MAX_BATCH || L > MLP_STREAM_MAX_LAYERS) return false;
    p.n_layers = L, p.batch = n, p.relu_mask = 0, p.max_fan_in = 0;
    for (int l = 0; l < L; l++)
    {
        const MlpLayer &ly = h->layers[l];
        if (ly.ldw != ly.fan_in) return false;
        p.layers[l].w = (const int8_t *)ly.w, p.layers[l].bias = (const int32_t *)ly.bias;
        p.layers[l].w_tiled = ly.w_tiled;
        p.layers[l].fan_in = ly.fan_in, p.layers[l].fan_out = ly.fan_out;
        p.max_fan_in = std::max(p.max_fan_in, ly.fan_in);
        const bool last = l == L - 1;
        if (h->desc.activation == NETCUDA_ACT_RELU_ALL || (h->desc.activation == NETCUDA_ACT_RELU_HIDDEN && !last)) p.relu_mask |= 1u << l;
    }
    p.in = in, p.act[0] = (int8_t *)h->act[0], p.act[1] = (int8_t *)h->act[1], p.out = out;
    p.barrier = h->stream_bar, p.error_flag = h->d_err;
    p.ll[0] = h->stream_ll[0], p.ll[1] = h->stream_ll[1];
    p.l2_prefetch_tiles = h->stream_pf_tiles;
    p.debug = nullptr, p.debug_cta = 0;
#ifdef NETCUDA_DEBUG_TIMELINE // clock-stamp hooks of tools/*_timeline.py: compiled out of release builds (NETCUDA_DEBUG_TIMELINE=1 python build.py)
    if (const char *dbg = getenv("NETCUDA_STREAM_DEBUG_PTR")) p.debug = reinterpret_cast<long long *>(strtoull(dbg, nullptr, 0));
    if (const char *dc = getenv("NETCUDA_STREAM_DEBUG_CTA")) p.debug_cta = atoi(dc);
#endif
    if (n <= h->stream_max_batch) return mlp_stream_supported(p, h->num_sms);
    return n >= h->umma_min_batch && mlp_umma_stream_supported(p);
}

namespace nc
{
thread_local bool g_pdl_small_pass = false;
}
struct SmallPassPdl // programmatic dependent launch for the kernels of a small (launch-latency bound) pass
{
    bool prev;
    explicit SmallPassPdl(bool on) : prev(nc::g_pdl_small_pass) { nc::g_pdl_small_pass = on; }
    ~SmallPassPdl() { nc::g_pdl_small_pass = prev; }
};

// MLP pass.  `in_f32` (fp32 [n][n_in]) or `in_i8` (int8 [n][n_in]); writes fp32 `out_f32` or int32 `out_i32`.
static int mlp_pass(netcuda_net *h, const float *in_f32, const int8_t *in_i8, int n, float *out_f32, int32_t *out_i32, cudaStream_t s)
{
    const int prec = h->desc.precision, kind = operand_kind(prec);
    const int L = (int)h->layers.size();
    const long long ld0 = h->layers[0].ldw;
    SmallPassPdl pdl(n <= GRAPH_MAX_SAMPLES);
    const void *cur;
    long long cur_ld;
    MlpStreamParams sp;
    // (an int8 input of a streamed pass is read in place: its rows are unpadded)
    const bool stream_in_place = in_i8 && n <= MLP_UMMA_STREAM_MAX_BATCH && mlp_stream_params(h, n, in_i8, out_i32 ? out_i32 : h->acc_out, sp);
    if (prec == NETCUDA_PREC_FP32)
    {
        cur = in_f32, cur_ld = (long long)h->n_in; // CUDA-core path reads the caller's matrix in place
    }
    else if (stream_in_place)
        cur = in_i8, cur_ld = ld0;
    else
    {
        KernelScope scope(h, s, "convert_in", 0.0, (double)n * ((double)h->n_in * (in_i8 ? 1 : 4) + (double)ld0 * h->elem));
        cudaError_t e;
        if (prec == NETCUDA_PREC_BF16)
            e = launch_convert_rows_bf16(in_f32, h->act[0], n, (int)h->n_in, (int)ld0, s);
        else if (prec == NETCUDA_PREC_TF32)
            e = launch_convert_rows_f32(in_f32, (float *)h->act[0], n, (int)h->n_in, (int)ld0, s);
        else if (in_i8)
            e = launch_pad_rows_i8(in_i8, (int8_t *)h->act[0], n, (int)h->n_in, (int)ld0, s);
        else
            e = launch_quantize_rows_q17(in_f32, (int8_t *)h->act[0], n, (int)h->n_in, (int)ld0, s);
        CK(e);
        cur = h->act[0], cur_ld = ld0;
    }
    int slot = 1;
    bool streamed = false;
    if (prec == NETCUDA_PREC_INT8 && n <= MLP_UMMA_STREAM_MAX_BATCH)
    {
        if (mlp_stream_params(h, n, (const int8_t *)cur, out_i32 ? out_i32 : h->acc_out, sp))
        {
            double bytes = 0.0, ops = 0.0;
            for (auto &ly : h->layers) bytes += (double)ly.fan_in * ly.fan_out + 4.0 * ly.fan_out, ops += 2.0 * n * (double)ly.fan_in * ly.fan_out;
            if (n <= h->stream_max_batch)
            {
                KernelScope scope(h, s, "mlp_stream", ops, bytes);
                CK(launch_mlp_i8_stream(sp, h->num_sms, s));
            }
            else
            {
                bool launched = false;
                if (h->umma_pair && mlp_umma_cluster_size(sp, h->num_sms) > 0)
                {
                    KernelScope scope(h, s, "mlp_umma_stream_cluster", ops, bytes);
                    const cudaError_t e = launch_mlp_i8_umma_cluster(sp, h->num_sms, h->umma_pair, s);
                    launched = e == cudaSuccess;
                    if (!launched)
                    {
                        // a driver that refuses cooperative cluster launches: the single-CTA kernel serves this handle from now on
                        (void)cudaGetLastError();
                        h->umma_pair = 0;
                        fprintf(stderr, "[netcuda] split-K cluster streaming kernel not launchable (%s); using the single-CTA kernel\n", cudaGetErrorString(e));
                    }
                }
                if (!launched)
                {
                    KernelScope scope(h, s, "mlp_umma_stream", ops, bytes);
                    CK(launch_mlp_i8_umma_stream(sp, h->num_sms, s));
                }
            }
            streamed = true;
        }
    }
    for (int l = 0; l < L && !streamed; l++)
    {
        const MlpLayer &ly = h->layers[l];
        const bool last = (l == L - 1);
        const bool relu = h->desc.activation == NETCUDA_ACT_RELU_ALL || (h->desc.activation == NETCUDA_ACT_RELU_HIDDEN && !last);
        void *dst;
        long long ldc;
        int out_type, epi;
        if (prec == NETCUDA_PREC_INT8)
        {
            if (last)
                dst = out_i32 ? (void *)out_i32 : (void *)h->acc_out, ldc = ly.fan_out, out_type = OUT_S32, epi = relu ? EPI_RELU : EPI_NONE;
            else
                dst = h->act[slot], ldc = h->layers[l + 1].ldw, out_type = OUT_S8, epi = relu ? EPI_REQUANT_RELU : EPI_REQUANT;
        }
        else
        {
            epi = relu ? EPI_RELU : EPI_NONE;
            if (last)
                dst = out_f32, ldc = ly.fan_out, out_type = OUT_F32;
            else
            {
                dst = h->act[slot], ldc = h->layers[l + 1].ldw;
                out_type = prec == NETCUDA_PREC_BF16 ? OUT_BF16 : OUT_F32;
            }
        }
        // the tensor maps are built with inner extent K = fan_in, so pad columns [fan_in, ld) are never read
        const int a_rows = (cur == (const void *)in_f32) ? n : h->max_batch;
        // Small batch, long K: a handful of tiles cannot pull the weights out of HBM fast enough.  Split K over the whole GPU;
        // the int32 partial sums meet in a zeroed workspace (TMA reduce-add: exact, order-independent), a second kernel adds
        // the bias, requantises and re-zeroes the workspace.  Same integers as the single-pass path, bit for bit.
        int splits = 1;
        if (prec == NETCUDA_PREC_INT8 && n <= SPLITK_MAX_M && h->gemm_variant == 0 && (ly.fan_out & 3) == 0)
        {
            const int tiles = (ly.fan_out + 255) / 256, num_kb = (ly.fan_in + 127) / 128;
            splits = std::min(h->num_sms / std::max(tiles, 1), num_kb / 4); // at least 4 k-blocks (512 B of K) per CTA
        }
        if (splits >= 2)
        {
            CK(run_gemm(h, "mlp_layer_splitk", kind, cur, cur_ld, a_rows, ly.w, ly.ldw, nullptr, h->splitk_ws, ly.fan_out, OUT_S32, EPI_SPLITK, n,
                        ly.fan_out, ly.fan_in, s, 0, 0, nullptr, splits));
            KernelScope scope(h, s, "splitk_finalize", 0.0, (double)n * ly.fan_out * (out_type == OUT_S8 ? 9.0 : 12.0));
            CK(launch_splitk_finalize(h->splitk_ws, (const int32_t *)ly.bias, dst, ldc, out_type == OUT_S8, relu, n, ly.fan_out, s));
        }
        else
            CK(run_gemm(h, "mlp_layer", kind, cur, cur_ld, a_rows, ly.w, ly.ldw, ly.bias, dst, ldc, out_type, epi, n, ly.fan_out, ly.fan_in, s));
        cur = dst, cur_ld = ldc;
        slot ^= 1;
    }
    if (prec == NETCUDA_PREC_INT8 && !out_i32)
    {
        KernelScope scope(h, s, "dequant_out", 0.0, (double)n * (double)h->n_out * 8.0);
        CK(launch_dequant_q214(h->acc_out, out_f32, (long long)n * (long long)h->n_out, s));
    }
    return NETCUDA_OK;
}

static int run_layernorm(netcuda_net *h, const char *label, const float *x, long long ldx, const float *g, const float *b, void *y,
                         long long ldy, int rows, int dim, cudaStream_t s)
{
    const bool f32 = h->vit_kind == GK_TF32;
    KernelScope scope(h, s, label, 0.0, (double)rows * dim * (f32 ? 8.0 : 6.0) + (double)dim * 8.0);
    CK(launch_layernorm(x, ldx, g, b, y, ldy, rows, dim, 1e-6f, s, f32));
    return NETCUDA_OK;
}

// ViT passes up to this many token rows run with programmatic dependent launch / are replayed from CUDA graphs (NETCUDA_VIT_PDL_ELEMS,
// NETCUDA_VIT_GRAPH_ROWS override the thresholds for A/B measurements; read once)
static long long vit_pdl_max_elems()
{
    static const long long v = getenv("NETCUDA_VIT_PDL_ELEMS") ? atoll(getenv("NETCUDA_VIT_PDL_ELEMS")) : (16LL << 20);
    return v;
}
static long long vit_graph_max_rows()
{
    static const long long v = getenv("NETCUDA_VIT_GRAPH_ROWS") ? atoll(getenv("NETCUDA_VIT_GRAPH_ROWS")) : (long long)VIT_GRAPH_MAX_ROWS;
    return v;
}

static int vit_pass(netcuda_net *h, const float *img, const uint8_t *img_u8, int n, float *logits, cudaStream_t s)
{
    const int D = h->desc.dim, F = h->desc.mlp_dim, C = h->desc.n_classes, T = h->T, NP = h->NP, PK = h->PK;
    const int rows = n * T, cap = h->max_batch * T;
    const int kind = h->vit_kind;
    const bool f32 = kind == GK_TF32;
    const int act_out = f32 ? OUT_F32 : OUT_BF16; // activations that feed the next GEMM's A operand
    // programmatic dependent launch pays while the kernels are short: ViT-Tiny, 256 images (50 k rows x 192): 112.8 -> 122.6 k images/s;
    // ViT-B, 512 images (101 k rows x 768): 24.8 -> 24.2 k images/s -- so it follows the size of the residual stream
    SmallPassPdl pdl((long long)rows * D <= vit_pdl_max_elems());
    {
        KernelScope scope(h, s, img_u8 ? "patchify_u8" : "patchify", 0.0, (double)n * (double)h->n_in * ((img_u8 ? 1.0 : 4.0) + (f32 ? 4.0 : 2.0)));
        if (img_u8)
            CK(launch_patchify_u8(img_u8, h->patches, n, h->desc.image_size, h->desc.patch_size, h->u8_mean, h->u8_inv_std, s, f32));
        else
            CK(launch_patchify(img, h->patches, n, h->desc.image_size, h->desc.patch_size, s, f32));
    }
    // patch embedding: x[b*T + 1 + t] = patches . patch_w^T + patch_b + pos[1 + t]
    CK(run_gemm(h, "patch_embed", kind, h->patches, PK, h->max_batch * NP, h->patch_w, PK, h->patch_b, h->x, D, OUT_F32, EPI_PATCH,
                n * NP, D, PK, s, NP, T, h->pos));
    {
        KernelScope scope(h, s, "cls_rows", 0.0, (double)n * D * 4.0 + (double)D * 8.0);
        CK(launch_cls_rows(h->x, h->cls, h->pos, n, T, D, s));
    }
    for (auto &b : h->blocks)
    {
        if (int rc = run_layernorm(h, "layernorm", h->x, D, b.ln1_g, b.ln1_b, h->ybuf, D, rows, D, s)) return rc;
        CK(run_gemm(h, "qkv", kind, h->ybuf, D, cap, b.qkv_w, D, b.qkv_b, h->qkv, 3LL * D, OUT_BF16, EPI_NONE, rows, 3 * D, D, s));
        {
            KernelScope scope(h, s, "attention", 4.0 * n * (double)T * T * D, (double)rows * D * (f32 ? 10.0 : 8.0));
            CK(launch_attention(h->qkv, h->att, n, T, h->desc.heads, s, h->d_err, h->num_sms, h->gemm_variant == 1 ? 1 : 0, f32));
        }
        CK(run_gemm(h, "proj", kind, h->att, D, cap, b.proj_w, D, b.proj_b, h->x, D, OUT_F32, EPI_RESIDUAL, rows, D, D, s));
        if (int rc = run_layernorm(h, "layernorm", h->x, D, b.ln2_g, b.ln2_b, h->ybuf, D, rows, D, s)) return rc;
        CK(run_gemm(h, "fc1", kind, h->ybuf, D, cap, b.fc1_w, D, b.fc1_b, h->hid, F, act_out, EPI_GELU, rows, F, D, s));
        CK(run_gemm(h, "fc2", kind, h->hid, F, cap, b.fc2_w, F, b.fc2_b, h->x, D, OUT_F32, EPI_RESIDUAL, rows, D, F, s));
    }
    // final LayerNorm on the class-token rows only (row pitch T*D), then the head
    if (int rc = run_layernorm(h, "layernorm_cls", h->x, (long long)T * D, h->lnf_g, h->lnf_b, h->cls_ln, D, n, D, s)) return rc;
    CK(run_gemm(h, "head", kind, h->cls_ln, D, h->max_batch, h->head_w, D, h->head_b, logits, C, OUT_F32, EPI_NONE, n, C, D, s));
    return NETCUDA_OK;
}

// MLP passes of a few thousand samples are a handful of microsecond kernels: the host cannot enqueue them (three tensor-map
// encodes + a launch each) as fast as the GPU runs them.  Such a pass is captured into a CUDA graph once and replayed while the
// caller keeps presenting the same buffers (what a serving loop and the host API's staging slots do).

// `body` enqueues the pass on `s`.  `second_sight`: capture only when the same (buffers, batch) combination has been seen before --
// capturing and instantiating ~90 kernel nodes costs more than a plain pass, so callers with ever-changing batch sizes never pay it.
template <class Body>
static int pass_graphed(netcuda_net *h, const void *in, void *out, int n, bool in_is_i8, bool out_is_i32, cudaStream_t s, bool second_sight,
                        Body body)
{
    netcuda_net::PassGraph *found = nullptr;
    for (auto &c : h->pass_graphs)
        if (c.exec && c.in == in && c.out == out && c.n == n && c.variant == h->gemm_variant && c.in_is_i8 == in_is_i8 && c.out_is_i32 == out_is_i32)
            found = &c;
    const bool hit = found != nullptr;
    if (!hit && second_sight)
    {
        bool again = false;
        for (auto &cand : h->graph_candidates)
            again = again || (cand.n == n && cand.in == in && cand.out == out && cand.variant == h->gemm_variant && cand.in_is_i8 == in_is_i8);
        if (!again)
        {
            netcuda_net::PassGraph &cand = h->graph_candidates[h->graph_candidate_next];
            h->graph_candidate_next = (h->graph_candidate_next + 1) % netcuda_net::PASS_GRAPHS;
            cand.in = in, cand.out = out, cand.n = n, cand.variant = h->gemm_variant, cand.in_is_i8 = in_is_i8;
            return body();
        }
    }
    if (!hit)
    {
        found = &h->pass_graphs[h->pass_graph_next];
        h->pass_graph_next = (h->pass_graph_next + 1) % netcuda_net::PASS_GRAPHS;
    }
    netcuda_net::PassGraph &g = *found;
    if (!hit)
    {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g.exec = nullptr;
        cudaGraph_t graph = nullptr;
        const uint64_t l0 = h->launches;
        if (s == nullptr || s == cudaStreamLegacy || cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        {
            // the default stream cannot be captured, and neither can a stream the caller is capturing already: plain launches
            (void)cudaGetLastError();
            return body();
        }
        const int rc = body();
        const cudaError_t e = cudaStreamEndCapture(s, &graph);
        if (rc != NETCUDA_OK)
        {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (e != cudaSuccess) return fail(NETCUDA_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
        const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) return fail(NETCUDA_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ei));
        g.in = in, g.out = out, g.n = n, g.variant = h->gemm_variant, g.in_is_i8 = in_is_i8, g.out_is_i32 = out_is_i32;
        g.launches = h->launches - l0;
        h->launches = l0;
    }
    CK(cudaGraphLaunch(g.exec, s));
    h->launches += g.launches;
    return NETCUDA_OK;
}

static int mlp_pass_graphed(netcuda_net *h, const float *in_f32, const int8_t *in_i8, int n, float *out_f32, int32_t *out_i32, cudaStream_t s)
{
    const void *in = in_i8 ? (const void *)in_i8 : (const void *)in_f32;
    void *out = out_i32 ? (void *)out_i32 : (void *)out_f32;
    return pass_graphed(h, in, out, n, in_i8 != nullptr, out_i32 != nullptr, s, false,
                        [&]() { return mlp_pass(h, in_f32, in_i8, n, out_f32, out_i32, s); });
}

// A ViT pass of a few samples is ~90 kernels of a few microseconds each: what a caller sees is the host's launch rate (three tensor-map
// encodes and a launch per GEMM: ViT-B, one sample: 1.18 ms per call).  The reference's own contract is one sample per launch_forward
// (src/netFPGA.cpp:266-289), so this is the drop-in's latency: such passes are replayed from a CUDA graph.

static int forward_device_impl(netcuda_net *h, const void *d_in, bool in_is_i8, size_t batch, void *d_out, bool out_is_i32, cudaStream_t s,
                               bool allow_graph = true)
{
    if (!h->weights_loaded) return fail(NETCUDA_ERR_INVALID, "forward before weights were uploaded");
    if (batch == 0) return NETCUDA_OK;
    if (!d_in || !d_out) return fail(NETCUDA_ERR_INVALID, "null device buffer");
    for (size_t done = 0; done < batch; done += (size_t)h->max_batch)
    {
        const int n = (int)std::min((size_t)h->max_batch, batch - done);
        int rc;
        if (h->desc.kind == NETCUDA_KIND_MLP)
        {
            const float *f = in_is_i8 ? nullptr : (const float *)d_in + done * h->n_in;
            const int8_t *q = in_is_i8 ? (const int8_t *)d_in + done * h->n_in : nullptr;
            float *of = out_is_i32 ? nullptr : (float *)d_out + done * h->n_out;
            int32_t *oi = out_is_i32 ? (int32_t *)d_out + done * h->n_out : nullptr;
            // single small pass, not being profiled (the profile brackets individual launches): graph replay
            // (NETCUDA_MLP_GRAPH_MIN_LAYERS: nets with fewer layers take plain launches; A/B)
            static const size_t graph_min_layers = getenv("NETCUDA_MLP_GRAPH_MIN_LAYERS") ? (size_t)atoi(getenv("NETCUDA_MLP_GRAPH_MIN_LAYERS")) : 2;
            MlpStreamParams sp;
            const bool streamed = mlp_stream_params(h, n, q ? q : (const int8_t *)h->act[0], oi ? oi : h->acc_out, sp); // (three launches at most)
            if (allow_graph && !streamed && h->layers.size() >= graph_min_layers && batch <= (size_t)GRAPH_MAX_SAMPLES && batch <= (size_t)h->max_batch &&
                !h->profiling && h->use_graphs)
                rc = mlp_pass_graphed(h, f, q, n, of, oi, s);
            else
                rc = mlp_pass(h, f, q, n, of, oi, s);
        }
        else
        {
            const float *img = in_is_i8 ? nullptr : (const float *)d_in + done * h->n_in;
            const uint8_t *img_u8 = in_is_i8 ? (const uint8_t *)d_in + done * h->n_in : nullptr;
            float *logits = (float *)d_out + done * h->n_out;
            static const bool multipass = getenv("NETCUDA_VIT_GRAPH_MULTIPASS") != nullptr; // (A/B: graphs for every pass of a multi-pass call)
            if (allow_graph && (long long)n * h->T <= vit_graph_max_rows() && (batch <= (size_t)h->max_batch || multipass) && !h->profiling && h->use_graphs)
                rc = pass_graphed(h, in_is_i8 ? (const void *)img_u8 : (const void *)img, logits, n, in_is_i8, false, s, true,
                                  [&]() { return vit_pass(h, img, img_u8, n, logits, s); });
            else
                rc = vit_pass(h, img, img_u8, n, logits, s);
        }
        if (rc != NETCUDA_OK) return rc;
    }
    return NETCUDA_OK;
}

extern "C" int netcuda_forward_device(netcuda_net *h, const void *d_in, size_t batch, void *d_out, void *stream)
{
    if (int rc = check_handle(h)) return rc;
    CK(cudaSetDevice(h->device));
    return forward_device_impl(h, d_in, false, batch, d_out, false, stream ? (cudaStream_t)stream : h->stream);
}

extern "C" int netcuda_forward_device_i8(netcuda_net *h, const int8_t *d_in, size_t batch, int32_t *d_out, void *stream)
{
    if (int rc = check_handle(h)) return rc;
    if (h->desc.kind != NETCUDA_KIND_MLP || h->desc.precision != NETCUDA_PREC_INT8)
        return fail(NETCUDA_ERR_INVALID, "netcuda_forward_device_i8 needs an INT8 MLP handle");
    CK(cudaSetDevice(h->device));
    return forward_device_impl(h, d_in, true, batch, d_out, true, stream ? (cudaStream_t)stream : h->stream);
}

// ---- forward: host buffers, pipelined staging ------------------------------------------------------------

// A device-side wait that ran out of its budget (~4 s: a pipeline bug, never a slow GPU) records its site code in the mapped flag and
// traps; the trap poisons the CUDA context, so the handle -- like every other handle of the process -- is unusable afterwards.  The
// code is decoded per kernel and the flag cleared once reported, so that a later call does not report a stale site.
static int kernel_error(netcuda_net *h)
{
    const int code = h->h_err ? *h->h_err : 0;
    if (code == 0) return NETCUDA_OK;
    *h->h_err = 0;
    const char *where = code == KERR_SMEM_ALIGN ? "misaligned dynamic shared memory window"
                        : code < 10               ? "tcgen05 GEMM pipeline"
                        : code < 20               ? "attention kernel pipeline"
                                                  : "INT8 weight-streaming kernel (grid barrier / weight ring)";
    return fail(NETCUDA_ERR_KERNEL, "device-side wait timed out in the %s (wait site %d); the CUDA context is poisoned by the trap", where, code);
}

// Pageable host inputs are staged through page-locked slots by the process-wide copy pool (staging.cpp: all cores the process
// may use, non-temporal stores).  The staging copy, not the GPU, sets the pace of net_cuda::launch_forward(std::vector) otherwise:
// a 1024-image ViT-B batch is 616 MB per call.
static bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

static int ensure_staging(netcuda_net *h, size_t in_elem, bool need_pin_in)
{
    const size_t slot_bytes = (size_t)h->max_batch * h->n_in * in_elem;
    if (h->stage_in_bytes < slot_bytes)
    {
        CK(cudaStreamSynchronize(h->copy_stream)); // (growing the slots: nothing may still be copying into them)
        CK(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < 2; i++)
        {
            if (h->dev_in[i]) cudaFree(h->dev_in[i]);
            if (h->pin_in[i]) cudaFreeHost(h->pin_in[i]);
            h->dev_in[i] = h->pin_in[i] = nullptr;
        }
        h->stage_in_bytes = 0;
        for (int i = 0; i < 2; i++) CK(cudaMalloc(&h->dev_in[i], slot_bytes));
        h->stage_in_bytes = slot_bytes;
    }
    if (need_pin_in && !h->pin_in[0])
        // write-combined: the CPU only ever writes these slots (with non-temporal stores), the DMA engine only reads them
        for (int i = 0; i < 2; i++) CK(cudaHostAlloc(&h->pin_in[i], h->stage_in_bytes, cudaHostAllocWriteCombined));
    return NETCUDA_OK;
}

// Retire one in-flight call: wait for its D2H, surface kernel errors, hand pageable callers their bytes.
static int finish_pending(netcuda_net *h, netcuda_net::Pending &pd)
{
    if (!pd.active) return pd.status;
    pd.active = false;
    const cudaError_t e = cudaEventSynchronize(pd.done);
    int rc = kernel_error(h);
    if (rc == NETCUDA_OK && e != cudaSuccess) rc = fail(NETCUDA_ERR_CUDA, "forward failed: %s", cudaGetErrorString(e));
    if (rc == NETCUDA_OK && !pd.pinned_out) memcpy(pd.user_out, pd.pin_out, pd.out_bytes);
    h->last_us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - pd.t0).count();
    if (getenv("NETCUDA_HOST_TRACE")) fprintf(stderr, "[netcuda] call retired after %.2f ms\n", h->last_us / 1e3);
    pd.status = rc;
    return rc;
}

// Enqueue one host-buffer forward: H2D of pass i+1 (copy stream) overlaps the kernels of pass i (compute stream), across
// calls as well as inside one; the D2H of the outputs follows the last pass on the compute stream.  Returns without waiting.
static int submit_host_impl(netcuda_net *h, const void *in, bool in_is_i8, size_t batch, void *out, bool out_is_i32, uint64_t *ticket)
{
    if (!h->weights_loaded) return fail(NETCUDA_ERR_INVALID, "forward before weights were uploaded");
    if (!in || !out) return fail(NETCUDA_ERR_INVALID, "null host buffer");
    if (batch == 0) return fail(NETCUDA_ERR_INVALID, "empty batch");
    const uint64_t tk = h->next_ticket;
    netcuda_net::Pending &pd = h->pending[tk % netcuda_net::MAX_IN_FLIGHT];
    if (pd.active) (void)finish_pending(h, pd); // ring full: the oldest call is retired first (its status stays readable)
    const size_t in_elem = in_is_i8 ? 1 : 4, out_elem = 4;
    // Small call (the reference's own contract is ONE sample per launch_forward, src/netFPGA.cpp:266-289): what the caller waits for is
    // driver calls, not bytes.  Such a call stays on the compute stream -- input copied into a page-locked slot of this ticket, one H2D,
    // the kernels, and (fp32 nets, whose last layer is a CUDA-core kernel with plain stores) the outputs written by the kernel straight
    // into page-locked host memory -- instead of the two-stream pipeline below: no pointer-attribute queries, no copy-stream events,
    // no D2H copy.  NETCUDA_SMALL_CALL=0 switches it off (A/B).
    static const bool small_call_on = !getenv("NETCUDA_SMALL_CALL") || atoi(getenv("NETCUDA_SMALL_CALL")) != 0;
    const bool small_call = small_call_on && h->desc.kind == NETCUDA_KIND_MLP && batch <= (size_t)h->max_batch &&
                            batch * h->n_in * in_elem <= SMALL_CALL_BYTES && batch * h->n_out * out_elem <= SMALL_CALL_BYTES;
    const bool pinned_in = small_call ? true : is_pinned(in);
    if (!small_call)
        if (int rc = ensure_staging(h, in_elem, !pinned_in)) return rc;
    pd.t0 = std::chrono::steady_clock::now();
    pd.pinned_out = small_call ? false : is_pinned(out);
    pd.user_out = out, pd.out_bytes = batch * h->n_out * out_elem;
    if (!pd.done) CK(cudaEventCreateWithFlags(&pd.done, cudaEventDisableTiming));
    // Output buffers of the in-flight ring.  When this call needs larger ones, every idle slot of the ring grows with it: the
    // allocations (milliseconds each, and cudaMalloc / cudaFree synchronise the device) are paid by one call instead of by each of the
    // next MAX_IN_FLIGHT calls as they come round to their slot.
    const bool need_pin_out = !pd.pinned_out;
    if (pd.dev_cap < pd.out_bytes || (need_pin_out && pd.pin_cap < pd.out_bytes))
    {
        // The graph of a small call holds BOTH output pointers of its ring slot (kernels -> dev_out, copy -> pin_out) but is found by
        // one of them: once a slot's buffers are replaced such a graph must go, or a later buffer that happens to get the old address
        // would revive it with the other pointer dangling.
        for (auto &g : h->pass_graphs)
            for (auto &q : h->pending)
                if (g.exec && g.in != nullptr && g.in == q.small_in)
                {
                    CK(cudaStreamSynchronize(h->stream)); // (a replay may still be running)
                    cudaGraphExecDestroy(g.exec);
                    g = netcuda_net::PassGraph();
                }
        for (auto &c : h->graph_candidates)
            for (auto &q : h->pending)
                if (c.in != nullptr && c.in == q.small_in) c = netcuda_net::PassGraph();
        for (auto &q : h->pending)
        {
            if (q.active) continue;
            if (q.dev_cap < pd.out_bytes)
            {
                if (q.dev_out) CK(cudaFree(q.dev_out)); // (cudaFree waits for the device: nothing can still be writing it)
                q.dev_out = nullptr, q.dev_cap = 0;
                CK(cudaMalloc(&q.dev_out, pd.out_bytes));
                q.dev_cap = pd.out_bytes;
            }
            if (need_pin_out && q.pin_cap < pd.out_bytes)
            {
                if (q.pin_out) CK(cudaFreeHost(q.pin_out));
                q.pin_out = nullptr, q.pin_cap = 0;
                CK(cudaHostAlloc(&q.pin_out, pd.out_bytes, cudaHostAllocDefault));
                q.pin_cap = pd.out_bytes;
            }
        }
    }

    if (small_call)
    {
        // Everything the call touches belongs to its ticket's ring slot (retired before it is reused), so nothing here depends on
        // the pipeline's input slots or their events.
        if (!pd.small_in) CK(cudaHostAlloc(&pd.small_in, SMALL_CALL_BYTES, cudaHostAllocDefault));
        if (!pd.small_dev_in) CK(cudaMalloc(&pd.small_dev_in, SMALL_CALL_BYTES));
        const size_t bytes = batch * h->n_in * in_elem;
        memcpy(pd.small_in, in, bytes);
        // page-locked host memory is device-addressable under unified addressing:
        //  * the ordered fp32 kernels store the last layer's outputs into it directly (no D2H copy);
        //  * a few samples of a tf32 / bf16 net are read from it in place (no H2D copy): the first kernel of such a pass converts the
        //    inputs, reading them once with plain loads (NETCUDA_SMALL_CALL_ZC_BYTES, default 16 KB; 0 = always copy).  Not for fp32
        //    nets -- every CTA of the first layer streams all the input rows through its cp.async ring, and over PCIe each ring
        //    refill is a bus round trip (config C1, one sample per call: 40.7 us in place, 27.3 us with the copy).
        // Copies and kernels of the call are replayed as ONE graph from the second sight of a (ring slot, batch) on
        // (NETCUDA_SMALL_CALL_GRAPH=0: plain launches; C1, one sample per call: bf16 37.9 -> 29.7 us, fp32 27.3 -> 26.3 us, tf32 unchanged).
        static const size_t zc_bytes = getenv("NETCUDA_SMALL_CALL_ZC_BYTES") ? (size_t)atoll(getenv("NETCUDA_SMALL_CALL_ZC_BYTES")) : (size_t)16 << 10;
        static const bool small_graph = !getenv("NETCUDA_SMALL_CALL_GRAPH") || atoi(getenv("NETCUDA_SMALL_CALL_GRAPH")) != 0;
        const bool direct_out = h->desc.precision == NETCUDA_PREC_FP32;
        const bool zc_in = bytes <= zc_bytes && !in_is_i8 && (h->desc.precision == NETCUDA_PREC_TF32 || h->desc.precision == NETCUDA_PREC_BF16);
        const void *src = zc_in ? pd.small_in : pd.small_dev_in;
        void *dst = direct_out ? pd.pin_out : pd.dev_out;
        const bool graphed = small_graph && h->use_graphs && !h->profiling && h->desc.precision != NETCUDA_PREC_INT8;
        auto body = [&]() -> int
        {
            if (!zc_in) CK(cudaMemcpyAsync(pd.small_dev_in, pd.small_in, bytes, cudaMemcpyHostToDevice, h->stream));
            if (int rc = forward_device_impl(h, src, in_is_i8, batch, dst, out_is_i32, h->stream, !graphed)) return rc;
            if (!direct_out) CK(cudaMemcpyAsync(pd.pin_out, pd.dev_out, pd.out_bytes, cudaMemcpyDeviceToHost, h->stream));
            return NETCUDA_OK;
        };
        if (int rc = graphed ? pass_graphed(h, pd.small_in, pd.pin_out, (int)batch, in_is_i8, out_is_i32, h->stream, true, body) : body()) return rc;
        CK(cudaEventRecord(pd.done, h->stream));
        pd.ticket = tk, pd.active = true, pd.status = NETCUDA_OK;
        h->next_ticket++;
        if (ticket) *ticket = tk;
        return NETCUDA_OK;
    }

    // The first chunk of a call is the only one whose H2D copy nothing hides (the compute stream may be idle): for large ViT batches
    // it is a quarter of a pass, so that the kernels start after a quarter of the copy time (blocking netcuda_forward, ViT-B,
    // 1024 images: 308 MB / 5.6 ms of exposed copy become 77 MB / 1.4 ms).
    // (When the previous call is still running, its kernels hide the copy and the pass keeps its full size.)
    bool gpu_busy = false;
    if (tk > 1)
    {
        netcuda_net::Pending &prev = h->pending[(tk - 1) % netcuda_net::MAX_IN_FLIGHT];
        gpu_busy = prev.active && prev.done && cudaEventQuery(prev.done) == cudaErrorNotReady;
        (void)cudaGetLastError();
    }
    // Pageable input: the staging copy (~17 us per ViT-B image on 16 cores) must hide behind the PREVIOUS chunk's kernels (~40 us per
    // image), so the chunks of a call into an idle GPU grow geometrically -- an eighth of a pass, then twice the previous one: only the
    // first, small staging copy is exposed, and no chunk waits for a staging copy longer than the kernels before it.
    size_t n = 0, ramp = 0;
    for (size_t done = 0; done < batch; done += n, h->chunk_seq++)
    {
        const int slot = (int)(h->chunk_seq & 1);
        n = std::min((size_t)h->max_batch, batch - done);
        if (!gpu_busy && h->desc.kind == NETCUDA_KIND_VIT && h->max_batch >= 256)
        {
            if (pinned_in)
            {
                if (done == 0 && n >= 256) n /= 4;
            }
            else
            {
                ramp = done == 0 ? (size_t)h->max_batch / 8 : ramp * 2;
                n = std::min(n, std::max(ramp, (size_t)32));
                // ... and the ramp ends ON a pass-size boundary (64, 128, 320, then full passes of 512 -- not 64, 128, 256, 512, 64: small
                // passes quantise badly into 256-row tiles, and a ragged one at the end of the call hides nothing)
                const size_t to_boundary = (size_t)h->max_batch - done % (size_t)h->max_batch;
                if (n < to_boundary && to_boundary <= 3 * n) n = std::min(to_boundary, batch - done);
            }
        }
        const size_t bytes = n * h->n_in * in_elem;
        const char *src = (const char *)in + done * h->n_in * in_elem;
        static const bool trace = getenv("NETCUDA_HOST_TRACE") != nullptr; // per-chunk host timeline on stderr (diagnostics)
        auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - pd.t0).count(); };
        if (trace) fprintf(stderr, "[netcuda] chunk %zu..%zu (slot %d): start %.2f ms", done, done + n, slot, since());
        if (h->chunk_seq >= 2) CK(cudaStreamWaitEvent(h->copy_stream, h->compute_done[slot], 0)); // device slot free again
        if (!pinned_in)
        {
            if (h->chunk_seq >= 2) CK(cudaEventSynchronize(h->h2d_done[slot])); // pinned slot drained
            if (trace) fprintf(stderr, ", slot drained %.2f", since());
            // staged and sent in up to four sub-copies: the DMA of one runs while the next is being staged, so a chunk costs
            // max(staging, H2D) instead of their sum (it matters most for the first chunk of a call, which nothing hides)
            const size_t sub = std::max<size_t>((bytes / 4 + ((size_t)1 << 21) - 1) & ~(((size_t)1 << 21) - 1), (size_t)8 << 20);
            for (size_t off = 0; off < bytes; off += sub)
            {
                const size_t len = std::min(sub, bytes - off);
                staging_copy((char *)h->pin_in[slot] + off, src + off, len);
                CK(cudaMemcpyAsync((char *)h->dev_in[slot] + off, (const char *)h->pin_in[slot] + off, len, cudaMemcpyHostToDevice, h->copy_stream));
            }
        }
        else
            CK(cudaMemcpyAsync(h->dev_in[slot], src, bytes, cudaMemcpyHostToDevice, h->copy_stream));
        CK(cudaEventRecord(h->h2d_done[slot], h->copy_stream));
        if (trace) fprintf(stderr, ", staged + H2D queued %.2f", since());
        CK(cudaStreamWaitEvent(h->stream, h->h2d_done[slot], 0));
        char *dout = (char *)pd.dev_out + done * h->n_out * out_elem;
        // (a multi-pass batch walks through fresh (input slot, output offset) pairs: nothing a cached graph could be reused for)
        if (int rc = forward_device_impl(h, h->dev_in[slot], in_is_i8, n, dout, out_is_i32, h->stream, batch <= (size_t)h->max_batch)) return rc;
        CK(cudaEventRecord(h->compute_done[slot], h->stream));
        if (trace) fprintf(stderr, ", kernels queued %.2f\n", since());
    }
    CK(cudaMemcpyAsync(pd.pinned_out ? out : pd.pin_out, pd.dev_out, pd.out_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaEventRecord(pd.done, h->stream));
    pd.ticket = tk, pd.active = true, pd.status = NETCUDA_OK;
    h->next_ticket++;
    if (ticket) *ticket = tk;
    return NETCUDA_OK;
}

static int wait_impl(netcuda_net *h, uint64_t ticket)
{
    if (ticket == 0 || ticket >= h->next_ticket) return fail(NETCUDA_ERR_INVALID, "unknown ticket %llu", (unsigned long long)ticket);
    netcuda_net::Pending &pd = h->pending[ticket % netcuda_net::MAX_IN_FLIGHT];
    if (pd.ticket != ticket) return NETCUDA_OK; // retired long ago (its slot has been reused since)
    return finish_pending(h, pd);
}

static int forward_host_impl(netcuda_net *h, const void *in, bool in_is_i8, size_t batch, void *out, bool out_is_i32)
{
    if (batch == 0) return h->weights_loaded ? NETCUDA_OK : fail(NETCUDA_ERR_INVALID, "forward before weights were uploaded");
    uint64_t tk = 0;
    if (int rc = submit_host_impl(h, in, in_is_i8, batch, out, out_is_i32, &tk)) return rc;
    return wait_impl(h, tk);
}

extern "C" int netcuda_host_register(const void *p, size_t bytes)
{
    if (!p || bytes == 0) return fail(NETCUDA_ERR_INVALID, "netcuda_host_register: empty range");
    const cudaError_t e = cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered)
    {
        (void)cudaGetLastError();
        return NETCUDA_OK;
    }
    if (e != cudaSuccess)
    {
        (void)cudaGetLastError();
        return fail(NETCUDA_ERR_CUDA, "cudaHostRegister(%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    return NETCUDA_OK;
}

extern "C" int netcuda_host_unregister(const void *p)
{
    if (!p) return NETCUDA_OK;
    const cudaError_t e = cudaHostUnregister(const_cast<void *>(p));
    if (e != cudaSuccess)
    {
        (void)cudaGetLastError();
        return fail(NETCUDA_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e));
    }
    return NETCUDA_OK;
}

extern "C" int netcuda_submit(netcuda_net *h, const float *in, size_t batch, float *out, uint64_t *ticket)
{
    if (int rc = check_handle(h)) return rc;
    if (!ticket) return fail(NETCUDA_ERR_INVALID, "null ticket");
    CK(cudaSetDevice(h->device));
    return submit_host_impl(h, in, false, batch, out, false, ticket);
}

extern "C" int netcuda_wait(netcuda_net *h, uint64_t ticket)
{
    if (int rc = check_handle(h)) return rc;
    CK(cudaSetDevice(h->device));
    return wait_impl(h, ticket);
}

extern "C" int netcuda_query(netcuda_net *h, uint64_t ticket, int *done)
{
    if (int rc = check_handle(h)) return rc;
    if (!done) return fail(NETCUDA_ERR_INVALID, "null done");
    if (ticket == 0 || ticket >= h->next_ticket) return fail(NETCUDA_ERR_INVALID, "unknown ticket %llu", (unsigned long long)ticket);
    netcuda_net::Pending &pd = h->pending[ticket % netcuda_net::MAX_IN_FLIGHT];
    *done = 1;
    if (pd.ticket == ticket && pd.active)
    {
        CK(cudaSetDevice(h->device));
        const cudaError_t e = cudaEventQuery(pd.done);
        if (e == cudaErrorNotReady)
            *done = 0;
        else if (e != cudaSuccess)
            return fail(NETCUDA_ERR_CUDA, "cudaEventQuery: %s", cudaGetErrorString(e));
    }
    return NETCUDA_OK;
}

extern "C" int netcuda_forward(netcuda_net *h, const float *in, size_t batch, float *out)
{
    if (int rc = check_handle(h)) return rc;
    CK(cudaSetDevice(h->device));
    return forward_host_impl(h, in, false, batch, out, false);
}

extern "C" int netcuda_forward_i8(netcuda_net *h, const int8_t *in, size_t batch, int32_t *out)
{
    if (int rc = check_handle(h)) return rc;
    if (h->desc.kind != NETCUDA_KIND_MLP || h->desc.precision != NETCUDA_PREC_INT8)
        return fail(NETCUDA_ERR_INVALID, "netcuda_forward_i8 needs an INT8 MLP handle");
    CK(cudaSetDevice(h->device));
    return forward_host_impl(h, in, true, batch, out, true);
}

// ---- u8 frames (ViT) ---------------------------------------------------------------------------------------

// Forget every captured pass (their kernel arguments are frozen at capture time).
static void invalidate_pass_graphs(netcuda_net *h)
{
    if (h->stream) cudaStreamSynchronize(h->stream); // a replay may still be running on the handle's stream
    for (auto &g : h->pass_graphs)
    {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = netcuda_net::PassGraph();
    }
    for (auto &c : h->graph_candidates) c = netcuda_net::PassGraph();
    (void)cudaGetLastError();
}

static int check_u8(netcuda_net *h)
{
    if (int rc = check_handle(h)) return rc;
    if (h->desc.kind != NETCUDA_KIND_VIT) return fail(NETCUDA_ERR_INVALID, "u8 frame input is a ViT feature (an MLP takes DATA_TYPE vectors)");
    return NETCUDA_OK;
}

extern "C" int netcuda_set_u8_normalization(netcuda_net *h, const float *mean, const float *stddev)
{
    if (int rc = check_u8(h)) return rc;
    if (!mean || !stddev) return fail(NETCUDA_ERR_INVALID, "null argument");
    for (int c = 0; c < 3; c++)
        if (!(stddev[c] > 0.0f)) return fail(NETCUDA_ERR_INVALID, "stddev[%d] must be positive", c);
    for (int c = 0; c < 3; c++)
    {
        h->u8_mean[c] = mean[c];
        h->u8_inv_std[c] = 1.0f / stddev[c];
    }
    // mean / inv_std are kernel arguments passed by value, so every captured pass has the old ones baked in: drop the graphs
    // (and the candidates, so that the next capture starts from a pass that ran with the new values)
    CK(cudaSetDevice(h->device));
    invalidate_pass_graphs(h);
    return NETCUDA_OK;
}

extern "C" int netcuda_forward_u8(netcuda_net *h, const uint8_t *frames, size_t batch, float *out)
{
    if (int rc = check_u8(h)) return rc;
    CK(cudaSetDevice(h->device));
    return forward_host_impl(h, frames, true, batch, out, false);
}

extern "C" int netcuda_submit_u8(netcuda_net *h, const uint8_t *frames, size_t batch, float *out, uint64_t *ticket)
{
    if (int rc = check_u8(h)) return rc;
    if (!ticket) return fail(NETCUDA_ERR_INVALID, "null ticket");
    CK(cudaSetDevice(h->device));
    return submit_host_impl(h, frames, true, batch, out, false, ticket);
}

extern "C" int netcuda_forward_device_u8(netcuda_net *h, const uint8_t *d_frames, size_t batch, float *d_out, void *stream)
{
    if (int rc = check_u8(h)) return rc;
    CK(cudaSetDevice(h->device));
    return forward_device_impl(h, d_frames, true, batch, d_out, false, stream ? (cudaStream_t)stream : h->stream);
}

// ---- introspection -----------------------------------------------------------------------------------------

extern "C" int netcuda_n_in(const netcuda_net *h, size_t *n)
{
    if (!h || !n) return fail(NETCUDA_ERR_INVALID, "null argument");
    *n = h->n_in;
    return NETCUDA_OK;
}
extern "C" int netcuda_n_out(const netcuda_net *h, size_t *n)
{
    if (!h || !n) return fail(NETCUDA_ERR_INVALID, "null argument");
    *n = h->n_out;
    return NETCUDA_OK;
}
extern "C" int netcuda_last_forward_us(const netcuda_net *h, int64_t *us)
{
    if (!h || !us) return fail(NETCUDA_ERR_INVALID, "null argument");
    *us = h->last_us;
    return NETCUDA_OK;
}
extern "C" int netcuda_launch_count(const netcuda_net *h, uint64_t *count)
{
    if (!h || !count) return fail(NETCUDA_ERR_INVALID, "null argument");
    *count = h->launches;
    return NETCUDA_OK;
}
extern "C" int netcuda_flops_per_sample(const netcuda_net *h, double *flops)
{
    if (!h || !flops) return fail(NETCUDA_ERR_INVALID, "null argument");
    *flops = h->flops_per_sample;
    return NETCUDA_OK;
}
extern "C" int netcuda_profile_enable(netcuda_net *h, int on)
{
    if (int rc = check_handle(h)) return rc;
    h->profiling = on != 0;
    return NETCUDA_OK;
}

extern "C" int netcuda_profile_read(netcuda_net *h, netcuda_kernel_stat *stats, int cap, int *count)
{
    if (int rc = check_handle(h)) return rc;
    if (!count || (cap > 0 && !stats)) return fail(NETCUDA_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    std::vector<netcuda_kernel_stat> acc(h->prof_labels.size());
    for (size_t i = 0; i < acc.size(); i++)
    {
        memset(&acc[i], 0, sizeof(acc[i]));
        snprintf(acc[i].label, sizeof(acc[i].label), "%s", h->prof_labels[i].c_str());
    }
    int rc = NETCUDA_OK;
    for (auto &r : h->prof_recs)
    {
        float ms = 0.0f;
        cudaError_t e = cudaEventSynchronize(r.stop);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.start, r.stop);
        if (e != cudaSuccess && rc == NETCUDA_OK) rc = fail(NETCUDA_ERR_CUDA, "profile read: %s", cudaGetErrorString(e));
        acc[r.label].launches++;
        acc[r.label].ms += ms;
        acc[r.label].flops += r.flops;
        acc[r.label].bytes += r.bytes;
        h->prof_free.push_back(r.start);
        h->prof_free.push_back(r.stop);
    }
    h->prof_recs.clear();
    *count = (int)acc.size();
    for (int i = 0; i < cap && i < (int)acc.size(); i++) stats[i] = acc[i];
    h->prof_labels.clear();
    return rc;
}

extern "C" int netcuda_set_gemm_variant(netcuda_net *h, int variant)
{
    if (int rc = check_handle(h)) return rc;
    if (variant < 0 || variant > 5) return fail(NETCUDA_ERR_INVALID, "variant must be 0..5");
    h->gemm_variant = variant;
    return NETCUDA_OK;
}

// ---- single-kernel entry points ---------------------------------------------------------------------------

static int op_prologue(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    {
        (void)cudaGetLastError();
        return fail(NETCUDA_ERR_NO_DEVICE, "device %d not available", device);
    }
    CK(cudaSetDevice(device));
    CK(gemm_global_init());
    return NETCUDA_OK;
}

extern "C" int netcuda_op_gemm(int device, int precision, int variant, const void *d_a, int lda, const void *d_w, int ldw,
                               const void *d_bias, void *d_out, int ldc, int out_type, int epilogue, int m, int n, int k, void *stream)
{
    if (int rc = op_prologue(device)) return rc;
    static const int epi_map[] = {EPI_NONE, EPI_RELU, EPI_GELU, EPI_RESIDUAL, EPI_REQUANT};
    if (epilogue < 0 || epilogue > NETCUDA_EPI_REQUANT) return fail(NETCUDA_ERR_INVALID, "unknown epilogue %d", epilogue);
    GemmCall c;
    c.kind = operand_kind(precision), c.variant = variant;
    c.a = d_a, c.lda = lda, c.a_rows = m, c.w = d_w, c.ldw = ldw, c.bias = d_bias;
    c.out = d_out, c.ldc = ldc, c.out_type = out_type, c.epi = epi_map[epilogue];
    // int8 output always requantises; NETCUDA_EPI_RELU selects the clamp-at-zero form
    if (out_type == NETCUDA_OUT_S8) c.epi = epilogue == NETCUDA_EPI_RELU ? EPI_REQUANT_RELU : EPI_REQUANT;
    c.m = m, c.n = n, c.k = k;
    c.k_splits = 1;
    c.remap_in = c.remap_out = 0, c.pos = nullptr;
    c.error_flag = nullptr;
    int sms = 0; // (cudaGetDeviceProperties costs milliseconds per call; the attribute query does not)
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    c.num_sms = sms;
    cudaError_t e = launch_gemm(c, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? NETCUDA_ERR_INVALID : NETCUDA_ERR_CUDA, "netcuda_op_gemm: %s", cudaGetErrorString(e));
    return NETCUDA_OK;
}

extern "C" int netcuda_op_layernorm(int device, const float *d_x, int ldx, const float *d_gamma, const float *d_beta, void *d_y, int ldy,
                                    int rows, int dim, float eps, void *stream)
{
    if (int rc = op_prologue(device)) return rc;
    CK(launch_layernorm(d_x, ldx, d_gamma, d_beta, d_y, ldy, rows, dim, eps, (cudaStream_t)stream));
    return NETCUDA_OK;
}

extern "C" int netcuda_op_attention(int device, const void *d_qkv, void *d_out, int batch, int tokens, int heads, void *stream)
{
    if (int rc = op_prologue(device)) return rc;
    // NETCUDA_ATTENTION_VARIANT=1 selects the mma.sync kernel for any sequence length (cross-check / profiling)
    const char *env = getenv("NETCUDA_ATTENTION_VARIANT");
    CK(launch_attention(d_qkv, d_out, batch, tokens, heads, (cudaStream_t)stream, nullptr, 0, env ? atoi(env) : 0));
    return NETCUDA_OK;
}

extern "C" int netcuda_op_attention_ex(int device, const void *d_qkv, void *d_out, int batch, int tokens, int heads, int kernel, int out_f32,
                                       void *stream)
{
    if (int rc = op_prologue(device)) return rc;
    const cudaError_t e = launch_attention(d_qkv, d_out, batch, tokens, heads, (cudaStream_t)stream, nullptr, 0, kernel == NETCUDA_ATT_KERNEL_MMA_SYNC ? 1 : 0,
                                           out_f32 != 0, kernel >= 0 ? kernel : -1);
    if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? NETCUDA_ERR_INVALID : NETCUDA_ERR_CUDA, "netcuda_op_attention_ex: %s", cudaGetErrorString(e));
    return NETCUDA_OK;
}

extern "C" int netcuda_op_patchify(int device, const float *d_img, void *d_patches, int batch, int image_size, int patch_size, void *stream)
{
    if (int rc = op_prologue(device)) return rc;
    CK(launch_patchify(d_img, d_patches, batch, image_size, patch_size, (cudaStream_t)stream));
    return NETCUDA_OK;
}

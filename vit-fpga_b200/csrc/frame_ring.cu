// frame_ring.cu -- the image side channel of net::net_abstract: filter_image / get_filtered_image.
//
// Reference (src/netFPGA.cpp:292-365, 443-482): a ring of BATCH_SIZE = 24 frames (:12).  filter_image copies the frame's
// original_h * original_w bytes into host slot `wr` (:314-315), writes them to the device (:321), enqueues the `image_process`
// task chained on the previous task's finish event (:326) and a NON-blocking read of the result into host slot `wr` (:328); when
// all 24 slots are in flight the frame is dropped ("PILA LLENA", :333).  get_filtered_image waits for the read event of the
// oldest slot (:349) and returns its bytes; an empty ring returns a header without pixels ("PILA VACIA", :359).
//
// Here: per ring one CUDA stream, `depth` page-locked input / output slots and device buffers, one event per slot.  push =
// copy into the pinned slot + H2D + filter kernel + D2H, all asynchronous on the stream (in-order = the reference's event
// chain); pop = cudaEventSynchronize on the oldest slot.  State is per ring (the reference's is namespace-global, :21-56).
//
// The device kernel `image_process` is absent from the reference (no .cl, no .aocx) and nothing describes what it computes
// beyond "one byte per pixel in, one byte per pixel out" (:441-442).  Builder decision, documented in DESIGN.md: a 3 x 3
// binomial smoothing filter on single-channel u8 frames, borders replicated, integer arithmetic
//     out[y][x] = (sum_{dy,dx in -1..1} w[dy] w[dx] in[clamp(y+dy)][clamp(x+dx)] + 8) >> 4,   w = (1, 2, 1)
// -- order-independent integers, hence bit-exact against oracle_filter3x3 (oracle/oracle_image.c).
#include "../../include/netcuda.h"
#include "kernels.h"

#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>

namespace
{

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    const int rc = nc::set_last_error_v(code, fmt, ap);
    va_end(ap);
    return rc;
}

#define RCK(expr)                                                                                                  \
    do                                                                                                             \
    {                                                                                                              \
        cudaError_t _e = (expr);                                                                                   \
        if (_e != cudaSuccess)                                                                                     \
            return fail(NETCUDA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// HBM-bound byte work: 1 byte read (+ halo, served by L1/L2) and 1 byte written per pixel.  One thread produces 4 horizontally
// adjacent pixels: three 6-byte row windows, horizontal (1, 2, 1) sums in registers, one 32-bit store when the row is aligned.
__global__ void __launch_bounds__(256) filter3x3_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int h, int w)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x0 >= w) return;
    int acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
    {
        const int yy = min(max(y + dy, 0), h - 1);
        const uint8_t *row = in + (long long)yy * w;
        int v[6];
#pragma unroll
        for (int i = 0; i < 6; i++) v[i] = row[min(max(x0 - 1 + i, 0), w - 1)];
        const int wy = dy == 0 ? 2 : 1;
#pragma unroll
        for (int i = 0; i < 4; i++) acc[i] += wy * (v[i] + 2 * v[i + 1] + v[i + 2]);
    }
    uint8_t *dst = out + (long long)y * w + x0;
    if (x0 + 4 <= w && ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0))
    {
        uint32_t word = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) word |= (uint32_t)((acc[i] + 8) >> 4) << (8 * i);
        *reinterpret_cast<uint32_t *>(dst) = word;
    }
    else
    {
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (x0 + i < w) dst[i] = (uint8_t)((acc[i] + 8) >> 4);
    }
}

cudaError_t launch_filter3x3(const uint8_t *in, uint8_t *out, int h, int w, cudaStream_t stream)
{
    if (h <= 0 || w <= 0) return cudaErrorInvalidValue;
    if (h > 65535) return cudaErrorInvalidValue;
    dim3 grid((unsigned)((w + 4 * 256 - 1) / (4 * 256)), (unsigned)h);
    filter3x3_kernel<<<grid, 256, 0, stream>>>(in, out, h, w);
    return cudaGetLastError();
}

} // namespace

struct netcuda_ring
{
    int device = 0, depth = 0;
    size_t max_pixels = 0;
    cudaStream_t stream = nullptr;
    struct Slot
    {
        uint8_t *pin_in = nullptr, *pin_out = nullptr, *dev_in = nullptr, *dev_out = nullptr;
        cudaEvent_t done = nullptr;
        size_t h = 0, w = 0;
    };
    std::vector<Slot> slots;
    int wr = 0, rd = 0, in_flight = 0; // g_wr_batch_cnt / g_rd_batch_cnt / BATCH_SIZE - g_free_batch of the reference
    uint64_t dropped = 0;              // frames refused because the ring was full
};

extern "C" int netcuda_ring_destroy(netcuda_ring *r)
{
    if (!r) return NETCUDA_OK;
    cudaSetDevice(r->device);
    if (r->stream) cudaStreamSynchronize(r->stream);
    for (auto &s : r->slots)
    {
        if (s.pin_in) cudaFreeHost(s.pin_in);
        if (s.pin_out) cudaFreeHost(s.pin_out);
        if (s.dev_in) cudaFree(s.dev_in);
        if (s.dev_out) cudaFree(s.dev_out);
        if (s.done) cudaEventDestroy(s.done);
    }
    if (r->stream) cudaStreamDestroy(r->stream);
    (void)cudaGetLastError();
    delete r;
    return NETCUDA_OK;
}

static int ring_create_impl(netcuda_ring *r)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        (void)cudaGetLastError();
        return fail(NETCUDA_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU fallback)");
    }
    if (r->device < 0 || r->device >= ndev) return fail(NETCUDA_ERR_INVALID, "device %d out of range [0,%d)", r->device, ndev);
    RCK(cudaSetDevice(r->device));
    RCK(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    r->slots.resize((size_t)r->depth);
    for (auto &s : r->slots)
    {
        RCK(cudaHostAlloc((void **)&s.pin_in, r->max_pixels, cudaHostAllocDefault));
        RCK(cudaHostAlloc((void **)&s.pin_out, r->max_pixels, cudaHostAllocDefault));
        RCK(cudaMalloc((void **)&s.dev_in, r->max_pixels));
        RCK(cudaMalloc((void **)&s.dev_out, r->max_pixels));
        RCK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    return NETCUDA_OK;
}

extern "C" int netcuda_ring_create(int device, int depth, size_t max_pixels, netcuda_ring **out)
{
    if (!out) return fail(NETCUDA_ERR_INVALID, "null out pointer");
    *out = nullptr;
    if (depth <= 0 || depth > 4096 || max_pixels == 0) return fail(NETCUDA_ERR_INVALID, "ring needs 1..4096 slots and a non-zero frame size");
    netcuda_ring *r = new netcuda_ring();
    r->device = device, r->depth = depth, r->max_pixels = max_pixels;
    const int rc = ring_create_impl(r);
    if (rc != NETCUDA_OK)
    {
        const std::string keep = netcuda_last_error();
        netcuda_ring_destroy(r);
        fail(rc, "%s", keep.c_str());
        return rc;
    }
    *out = r;
    return NETCUDA_OK;
}

extern "C" int netcuda_ring_push(netcuda_ring *r, const uint8_t *pixels, size_t h, size_t w)
{
    if (!r || !pixels) return fail(NETCUDA_ERR_INVALID, "null argument");
    if (h == 0 || w == 0 || h > 65535 || w > (1u << 30) / h || h * w > r->max_pixels)
        return fail(NETCUDA_ERR_INVALID, "frame of %zu x %zu pixels does not fit the ring's %zu-byte slots", h, w, r->max_pixels);
    if (r->in_flight == r->depth)
    {
        r->dropped++;
        return fail(NETCUDA_ERR_RING_FULL, "frame ring full (%d frames in flight): pop a result first", r->depth); // "PILA LLENA", src/netFPGA.cpp:333
    }
    RCK(cudaSetDevice(r->device));
    netcuda_ring::Slot &s = r->slots[(size_t)r->wr];
    const size_t bytes = h * w;
    memcpy(s.pin_in, pixels, bytes); // the caller's frame is free again when push returns (the reference copies it too, :314-315)
    s.h = h, s.w = w;
    RCK(cudaMemcpyAsync(s.dev_in, s.pin_in, bytes, cudaMemcpyHostToDevice, r->stream));
    RCK(launch_filter3x3(s.dev_in, s.dev_out, (int)h, (int)w, r->stream));
    RCK(cudaMemcpyAsync(s.pin_out, s.dev_out, bytes, cudaMemcpyDeviceToHost, r->stream));
    RCK(cudaEventRecord(s.done, r->stream));
    r->wr = (r->wr + 1) % r->depth;
    r->in_flight++;
    return NETCUDA_OK;
}

extern "C" int netcuda_ring_pop(netcuda_ring *r, uint8_t *pixels_out, size_t capacity, size_t *h, size_t *w)
{
    if (!r || !h || !w) return fail(NETCUDA_ERR_INVALID, "null argument");
    if (r->in_flight == 0) return fail(NETCUDA_ERR_RING_EMPTY, "frame ring empty"); // "PILA VACIA", src/netFPGA.cpp:359
    netcuda_ring::Slot &s = r->slots[(size_t)r->rd];
    const size_t bytes = s.h * s.w;
    if (!pixels_out || capacity < bytes) return fail(NETCUDA_ERR_INVALID, "output buffer of %zu bytes, the oldest frame holds %zu", capacity, bytes);
    RCK(cudaSetDevice(r->device));
    RCK(cudaEventSynchronize(s.done)); // clWaitForEvents on the slot's read event, src/netFPGA.cpp:349
    memcpy(pixels_out, s.pin_out, bytes);
    *h = s.h, *w = s.w;
    r->rd = (r->rd + 1) % r->depth;
    r->in_flight--;
    return NETCUDA_OK;
}

extern "C" int netcuda_ring_peek(netcuda_ring *r, size_t *h, size_t *w)
{
    if (!r || !h || !w) return fail(NETCUDA_ERR_INVALID, "null argument");
    if (r->in_flight == 0) return fail(NETCUDA_ERR_RING_EMPTY, "frame ring empty");
    *h = r->slots[(size_t)r->rd].h, *w = r->slots[(size_t)r->rd].w;
    return NETCUDA_OK;
}

extern "C" int netcuda_ring_in_flight(const netcuda_ring *r, int *count, uint64_t *dropped)
{
    if (!r || !count) return fail(NETCUDA_ERR_INVALID, "null argument");
    *count = r->in_flight;
    if (dropped) *dropped = r->dropped;
    return NETCUDA_OK;
}

extern "C" int netcuda_op_filter3x3(int device, const uint8_t *d_in, uint8_t *d_out, int h, int w, void *stream)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    {
        (void)cudaGetLastError();
        return fail(NETCUDA_ERR_NO_DEVICE, "device %d not available", device);
    }
    RCK(cudaSetDevice(device));
    const cudaError_t e = launch_filter3x3(d_in, d_out, h, w, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? NETCUDA_ERR_INVALID : NETCUDA_ERR_CUDA, "netcuda_op_filter3x3: %s", cudaGetErrorString(e));
    return NETCUDA_OK;
}

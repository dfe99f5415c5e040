// weights_io.cu -- weight files: the flat layout of src/netFPGA.cpp:91-106 on disk (format: include/netcuda.h).
//
// The reference keeps a net only in memory: the constructor flattens net::net_data (src/netFPGA.cpp:89-107) and
// get_net_data (:206-237) is meant to undo that.  A serving backend has to reload the same net in the next process, so the
// flat arrays get a small header, a CRC and a loader that feeds netcuda_create / netcuda_upload_*.  Pure host code
// except netcuda_create_from_file.
#include "../../include/netcuda.h"
#include "kernels.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
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

constexpr char MAGIC[8] = {'N', 'E', 'T', 'C', 'U', 'D', 'A', 'W'};
constexpr uint32_t VERSION = 2; // 2: the checksum covers the header fields as well (with the crc field itself read as zero)
constexpr size_t HEADER_BYTES = 96;

struct Header // 96 bytes, little endian, no implicit padding
{
    char magic[8];
    uint32_t version, kind, dtype, activation, n_ins, n_layers;
    uint32_t vit[7];
    uint32_t reserved;
    uint64_t n_weights, n_biases;
    uint32_t crc;
    uint8_t pad[12];
};
static_assert(sizeof(Header) == HEADER_BYTES, "header layout");

uint32_t crc32_update(uint32_t crc, const void *data, size_t n)
{
    static uint32_t table[256];
    static bool ready = false;
    if (!ready)
    {
        for (uint32_t i = 0; i < 256; i++)
        {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    const uint8_t *p = static_cast<const uint8_t *>(data);
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
    return ~crc;
}

size_t pad16(size_t n) { return (n + 15) / 16 * 16; }

struct FileCloser
{
    void operator()(FILE *f) const
    {
        if (f) fclose(f);
    }
};
using File = std::unique_ptr<FILE, FileCloser>;

// one section: `bytes` of payload + zero padding to 16 bytes, CRC carried along
int write_section(FILE *f, const void *data, size_t bytes, uint32_t &crc, const char *path)
{
    static const uint8_t zeros[16] = {0};
    const size_t padn = pad16(bytes) - bytes;
    if (bytes && fwrite(data, 1, bytes, f) != bytes) return fail(NETCUDA_ERR_INVALID, "short write to %s", path);
    if (padn && fwrite(zeros, 1, padn, f) != padn) return fail(NETCUDA_ERR_INVALID, "short write to %s", path);
    crc = crc32_update(crc, data, bytes);
    crc = crc32_update(crc, zeros, padn);
    return NETCUDA_OK;
}

int write_file(const char *path, Header hd, const int32_t *npl, const void *w, size_t w_elem, const void *b, size_t b_elem)
{
    if (!path) return fail(NETCUDA_ERR_INVALID, "null path");
    File f(fopen(path, "wb"));
    if (!f) return fail(NETCUDA_ERR_INVALID, "cannot open %s for writing", path);
    memcpy(hd.magic, MAGIC, 8);
    hd.version = VERSION, hd.reserved = 0, hd.crc = 0;
    memset(hd.pad, 0, sizeof(hd.pad));
    if (fwrite(&hd, 1, sizeof(hd), f.get()) != sizeof(hd)) return fail(NETCUDA_ERR_INVALID, "short write to %s", path);
    uint32_t crc = crc32_update(0, &hd, sizeof(hd)); // header with crc = 0: a flipped dimension or activation is caught like a flipped weight
    if (int rc = write_section(f.get(), npl, (size_t)hd.n_layers * 4, crc, path)) return rc;
    if (int rc = write_section(f.get(), w, (size_t)hd.n_weights * w_elem, crc, path)) return rc;
    if (int rc = write_section(f.get(), b, (size_t)hd.n_biases * b_elem, crc, path)) return rc;
    hd.crc = crc;
    if (fseek(f.get(), 0, SEEK_SET) != 0 || fwrite(&hd, 1, sizeof(hd), f.get()) != sizeof(hd) || fflush(f.get()) != 0)
        return fail(NETCUDA_ERR_INVALID, "cannot finalise %s", path);
    return NETCUDA_OK;
}

int mlp_counts(const int32_t *npl, int n_layers, int n_ins, uint64_t *nw, uint64_t *nb)
{
    if (!npl || n_layers <= 0 || n_layers > NETCUDA_FILE_MAX_LAYERS || n_ins <= 0)
        return fail(NETCUDA_ERR_INVALID, "MLP file needs n_ins > 0 and 1..%d layers", NETCUDA_FILE_MAX_LAYERS);
    uint64_t w = 0, b = 0, fan_in = (uint64_t)n_ins;
    for (int l = 0; l < n_layers; l++)
    {
        if (npl[l] <= 0) return fail(NETCUDA_ERR_INVALID, "n_p_l[%d] must be positive", l);
        w += fan_in * (uint64_t)npl[l]; // W_l[out][in], src/netFPGA.cpp:91-106 (each product < 2^62, the sum is bounded right below)
        b += (uint64_t)npl[l];
        // a crafted layer table must not wrap the 64-bit sum: no net this library can hold comes near 2^40 weights
        if (w > (1ull << 40)) return fail(NETCUDA_ERR_INVALID, "layer table describes more than 2^40 weights");
        fan_in = (uint64_t)npl[l];
    }
    *nw = w, *nb = b;
    return NETCUDA_OK;
}

size_t w_elem_size(uint32_t dtype) { return dtype == NETCUDA_FILE_Q17 ? 1 : 4; }

// header + n_p_l, validated against the file length
int read_header(FILE *f, const char *path, Header *hd, int32_t *npl)
{
    if (fread(hd, 1, sizeof(*hd), f) != sizeof(*hd)) return fail(NETCUDA_ERR_INVALID, "%s: shorter than a weight-file header", path);
    if (memcmp(hd->magic, MAGIC, 8) != 0) return fail(NETCUDA_ERR_INVALID, "%s: not a netcuda weight file (bad magic)", path);
    if (hd->version != VERSION) return fail(NETCUDA_ERR_UNSUPPORTED, "%s: file version %u, this library reads version %u", path, hd->version, VERSION);
    if (hd->kind > NETCUDA_KIND_VIT || hd->dtype > NETCUDA_FILE_Q17) return fail(NETCUDA_ERR_INVALID, "%s: unknown kind / dtype", path);
    if (hd->n_layers > NETCUDA_FILE_MAX_LAYERS) return fail(NETCUDA_ERR_INVALID, "%s: %u layers (limit %d)", path, hd->n_layers, NETCUDA_FILE_MAX_LAYERS);
    const size_t npl_bytes = (size_t)hd->n_layers * 4;
    int32_t tmp[NETCUDA_FILE_MAX_LAYERS + 4] = {0};
    if (fread(tmp, 1, pad16(npl_bytes), f) != pad16(npl_bytes)) return fail(NETCUDA_ERR_INVALID, "%s: truncated layer table", path);
    memcpy(npl, tmp, npl_bytes);
    if (hd->kind == NETCUDA_KIND_MLP)
    {
        uint64_t nw = 0, nb = 0;
        if (int rc = mlp_counts(npl, (int)hd->n_layers, (int)hd->n_ins, &nw, &nb)) return rc;
        if (nw != hd->n_weights || nb != hd->n_biases) return fail(NETCUDA_ERR_INVALID, "%s: element counts do not match the layer table", path);
    }
    else if (hd->dtype != NETCUDA_FILE_F32 || hd->n_biases != 0 || hd->n_layers != 0)
        return fail(NETCUDA_ERR_INVALID, "%s: a ViT file holds one fp32 vector", path);
    const uint64_t expect = HEADER_BYTES + pad16(npl_bytes) + pad16(hd->n_weights * w_elem_size(hd->dtype)) + pad16(hd->n_biases * 4);
    if (fseek(f, 0, SEEK_END) != 0) return fail(NETCUDA_ERR_INVALID, "%s: cannot seek", path);
    const long len = ftell(f);
    if (len < 0 || (uint64_t)len != expect)
        return fail(NETCUDA_ERR_INVALID, "%s: %ld bytes on disk, header describes %llu", path, len, (unsigned long long)expect);
    return NETCUDA_OK;
}

void fill_info(const Header &hd, const int32_t *npl, netcuda_file_info *info)
{
    memset(info, 0, sizeof(*info));
    memcpy(info->n_p_l, npl, (size_t)hd.n_layers * 4);
    netcuda_desc &d = info->desc;
    d.kind = (int32_t)hd.kind, d.activation = (int32_t)hd.activation;
    d.n_ins = (int32_t)hd.n_ins, d.n_layers = (int32_t)hd.n_layers, d.n_p_l = info->n_p_l;
    d.image_size = (int32_t)hd.vit[0], d.patch_size = (int32_t)hd.vit[1], d.dim = (int32_t)hd.vit[2], d.depth = (int32_t)hd.vit[3];
    d.heads = (int32_t)hd.vit[4], d.mlp_dim = (int32_t)hd.vit[5], d.n_classes = (int32_t)hd.vit[6];
    d.precision = hd.kind == NETCUDA_KIND_VIT ? NETCUDA_PREC_BF16 : hd.dtype == NETCUDA_FILE_Q17 ? NETCUDA_PREC_INT8 : NETCUDA_PREC_FP32;
    info->dtype = (int32_t)hd.dtype, info->n_weights = hd.n_weights, info->n_biases = hd.n_biases;
}

} // namespace

extern "C" int netcuda_file_write_mlp(const char *path, const int32_t *n_p_l, int n_layers, int n_ins, int activation, const float *w_flat,
                                      const float *b_flat)
{
    Header hd = {};
    if (int rc = mlp_counts(n_p_l, n_layers, n_ins, &hd.n_weights, &hd.n_biases)) return rc;
    if (!w_flat || !b_flat) return fail(NETCUDA_ERR_INVALID, "null weights");
    if (activation < NETCUDA_ACT_RELU_HIDDEN || activation > NETCUDA_ACT_NONE) return fail(NETCUDA_ERR_INVALID, "unknown activation %d", activation);
    hd.kind = NETCUDA_KIND_MLP, hd.dtype = NETCUDA_FILE_F32, hd.activation = (uint32_t)activation;
    hd.n_ins = (uint32_t)n_ins, hd.n_layers = (uint32_t)n_layers;
    return write_file(path, hd, n_p_l, w_flat, 4, b_flat, 4);
}

extern "C" int netcuda_file_write_mlp_i8(const char *path, const int32_t *n_p_l, int n_layers, int n_ins, int activation, const int8_t *w_flat,
                                         const int32_t *b_flat)
{
    Header hd = {};
    if (int rc = mlp_counts(n_p_l, n_layers, n_ins, &hd.n_weights, &hd.n_biases)) return rc;
    if (!w_flat || !b_flat) return fail(NETCUDA_ERR_INVALID, "null weights");
    if (activation < NETCUDA_ACT_RELU_HIDDEN || activation > NETCUDA_ACT_NONE) return fail(NETCUDA_ERR_INVALID, "unknown activation %d", activation);
    hd.kind = NETCUDA_KIND_MLP, hd.dtype = NETCUDA_FILE_Q17, hd.activation = (uint32_t)activation;
    hd.n_ins = (uint32_t)n_ins, hd.n_layers = (uint32_t)n_layers;
    return write_file(path, hd, n_p_l, w_flat, 1, b_flat, 4);
}

extern "C" int netcuda_file_write_vit(const char *path, const netcuda_desc *desc, const float *flat, size_t count)
{
    if (!desc || !flat) return fail(NETCUDA_ERR_INVALID, "null argument");
    size_t expect = 0;
    netcuda_desc d = *desc;
    d.kind = NETCUDA_KIND_VIT, d.precision = NETCUDA_PREC_BF16;
    if (int rc = netcuda_vit_param_count(&d, &expect)) return rc;
    if (count != expect) return fail(NETCUDA_ERR_INVALID, "ViT file: got %zu floats, the descriptor needs %zu", count, expect);
    Header hd = {};
    hd.kind = NETCUDA_KIND_VIT, hd.dtype = NETCUDA_FILE_F32;
    const int32_t v[7] = {d.image_size, d.patch_size, d.dim, d.depth, d.heads, d.mlp_dim, d.n_classes};
    for (int i = 0; i < 7; i++) hd.vit[i] = (uint32_t)v[i];
    hd.n_weights = count, hd.n_biases = 0;
    return write_file(path, hd, nullptr, flat, 4, nullptr, 4);
}

extern "C" int netcuda_file_info_read(const char *path, netcuda_file_info *info)
{
    if (!path || !info) return fail(NETCUDA_ERR_INVALID, "null argument");
    File f(fopen(path, "rb"));
    if (!f) return fail(NETCUDA_ERR_INVALID, "cannot open %s", path);
    Header hd;
    int32_t npl[NETCUDA_FILE_MAX_LAYERS] = {0};
    if (int rc = read_header(f.get(), path, &hd, npl)) return rc;
    fill_info(hd, npl, info);
    return NETCUDA_OK;
}

extern "C" int netcuda_file_read(const char *path, void *weights, size_t weight_bytes, void *biases, size_t bias_bytes)
{
    if (!path) return fail(NETCUDA_ERR_INVALID, "null path");
    File f(fopen(path, "rb"));
    if (!f) return fail(NETCUDA_ERR_INVALID, "cannot open %s", path);
    Header hd;
    int32_t npl[NETCUDA_FILE_MAX_LAYERS] = {0};
    if (int rc = read_header(f.get(), path, &hd, npl)) return rc;
    const size_t wb = (size_t)hd.n_weights * w_elem_size(hd.dtype), bb = (size_t)hd.n_biases * 4;
    if (weight_bytes != wb || bias_bytes != bb) return fail(NETCUDA_ERR_INVALID, "%s holds %zu + %zu payload bytes, buffers are %zu + %zu", path, wb, bb, weight_bytes, bias_bytes);
    if ((wb && !weights) || (bb && !biases)) return fail(NETCUDA_ERR_INVALID, "null buffer");
    const size_t npl_bytes = pad16((size_t)hd.n_layers * 4);
    if (fseek(f.get(), (long)HEADER_BYTES, SEEK_SET) != 0) return fail(NETCUDA_ERR_INVALID, "%s: cannot seek", path);
    uint8_t scratch[NETCUDA_FILE_MAX_LAYERS * 4 + 16];
    Header hz = hd;
    hz.crc = 0;
    uint32_t crc = crc32_update(0, &hz, sizeof(hz));
    if (fread(scratch, 1, npl_bytes, f.get()) != npl_bytes) return fail(NETCUDA_ERR_INVALID, "%s: truncated", path);
    crc = crc32_update(crc, scratch, npl_bytes);
    auto section = [&](void *dst, size_t bytes) -> int {
        if (bytes && fread(dst, 1, bytes, f.get()) != bytes) return fail(NETCUDA_ERR_INVALID, "%s: truncated", path);
        crc = crc32_update(crc, dst, bytes);
        const size_t padn = pad16(bytes) - bytes;
        if (padn && fread(scratch, 1, padn, f.get()) != padn) return fail(NETCUDA_ERR_INVALID, "%s: truncated", path);
        crc = crc32_update(crc, scratch, padn);
        return NETCUDA_OK;
    };
    if (int rc = section(weights, wb)) return rc;
    if (int rc = section(biases, bb)) return rc;
    if (crc != hd.crc) return fail(NETCUDA_ERR_INVALID, "%s: checksum mismatch (file %08x, computed %08x)", path, hd.crc, crc);
    return NETCUDA_OK;
}

extern "C" int netcuda_create_from_file(const char *path, int precision, int device, int max_batch, netcuda_t **out)
{
    if (!out) return fail(NETCUDA_ERR_INVALID, "null out");
    *out = nullptr;
    netcuda_file_info info;
    if (int rc = netcuda_file_info_read(path, &info)) return rc;
    netcuda_desc d = info.desc; // n_p_l points into `info`, alive for the whole call
    if (precision >= 0) d.precision = precision;
    if (info.dtype == NETCUDA_FILE_Q17 && d.precision != NETCUDA_PREC_INT8)
        return fail(NETCUDA_ERR_INVALID, "%s holds Q1.7 integers: it can only be opened as NETCUDA_PREC_INT8", path);
    d.device = device, d.max_batch = max_batch;
    const size_t wb = (size_t)info.n_weights * w_elem_size((uint32_t)info.dtype), bb = (size_t)info.n_biases * 4;
    std::vector<uint8_t> w(wb), b(bb);
    if (int rc = netcuda_file_read(path, w.data(), wb, b.data(), bb)) return rc;
    netcuda_t *h = nullptr;
    if (int rc = netcuda_create(&d, &h)) return rc;
    int rc;
    if (d.kind == NETCUDA_KIND_VIT)
        rc = netcuda_upload_vit(h, reinterpret_cast<const float *>(w.data()), (size_t)info.n_weights);
    else if (info.dtype == NETCUDA_FILE_Q17)
        rc = netcuda_upload_mlp_i8(h, reinterpret_cast<const int8_t *>(w.data()), reinterpret_cast<const int32_t *>(b.data()));
    else
        rc = netcuda_upload_mlp(h, reinterpret_cast<const float *>(w.data()), reinterpret_cast<const float *>(b.data()));
    if (rc != NETCUDA_OK)
    {
        netcuda_destroy(h); // (keeps the upload's error message: destroy does not touch it)
        return rc;
    }
    *out = h;
    return NETCUDA_OK;
}

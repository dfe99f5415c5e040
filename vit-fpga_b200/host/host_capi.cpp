// host_capi.cpp -- C entry points that drive cuda::net_cuda strictly through net::net_abstract*,
// the way the reference's (absent) host application drives fpga::net_fpga.  Used by tests/ and
// bench.py via ctypes; the same calls a C++ application would make are spelled out in INTEGRATION.md.
#include <netCUDA.h>

#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <exception>
#include <vector>

namespace
{
thread_local char g_err[512] = "";
void set_err(const char *what)
{
    snprintf(g_err, sizeof(g_err), "%s", what);
}

net::net_data make_data(const int *npl, int n_layers, int n_ins, const float *w, const float *b)
{
    net::net_data d;
    d.n_ins = (size_t)n_ins;
    d.n_layers = (size_t)n_layers;
    size_t fan_in = (size_t)n_ins;
    for (int l = 0; l < n_layers; l++)
    {
        d.n_p_l.push_back((size_t)npl[l]);
        d.params.emplace_back();
        d.bias.emplace_back();
        for (int j = 0; j < npl[l]; j++)
        {
            if (w)
                d.params[l].emplace_back(w, w + fan_in), w += fan_in;
            else
                d.params[l].emplace_back(fan_in, 0.0f);
            d.bias[l].push_back(b ? *b++ : 0.0f);
        }
        fan_in = (size_t)npl[l];
    }
    return d;
}

cuda::net_cuda_options make_opt(int precision, int device, int activation, int max_batch)
{
    cuda::net_cuda_options o;
    o.precision = precision, o.device = device, o.activation = activation, o.max_batch = max_batch;
    return o;
}
}

extern "C" {

const char *nch_last_error(void) { return g_err; }

// precision < 0: use the 3-argument reference-shaped constructor (defaults from the environment).
void *nch_mlp_create(const int *npl, int n_layers, int n_ins, const float *w, const float *b, int random, unsigned seed,
                     int precision, int device, int activation, int max_batch)
{
    try
    {
        net::net_data d = make_data(npl, n_layers, n_ins, random ? nullptr : w, random ? nullptr : b);
        if (random) srand(seed);
        net::net_abstract *n;
        if (precision < 0)
            n = new cuda::net_cuda(d, false, random != 0);
        else
            n = new cuda::net_cuda(d, make_opt(precision, device, activation, max_batch), random != 0);
        return n;
    }
    catch (const std::exception &e)
    {
        set_err(e.what());
        return nullptr;
    }
}

void *nch_vit_create(int image_size, int patch_size, int dim, int depth, int heads, int mlp_dim, int n_classes, const float *flat,
                     size_t count, int device, int max_batch, int precision)
{
    try
    {
        cuda::vit_data v;
        v.image_size = image_size, v.patch_size = patch_size, v.dim = dim, v.depth = depth, v.heads = heads;
        v.mlp_dim = mlp_dim, v.n_classes = n_classes;
        v.params.assign(flat, flat + count);
        net::net_abstract *n = new cuda::net_cuda(v, make_opt(precision, device, 0, max_batch));
        return n;
    }
    catch (const std::exception &e)
    {
        set_err(e.what());
        return nullptr;
    }
}

// net_cuda::save / net_cuda::load (weight files).  nch_load returns a net behind net::net_abstract*, or NULL.
int nch_save(void *net, const char *path)
{
    try
    {
        cuda::net_cuda *n = dynamic_cast<cuda::net_cuda *>(static_cast<net::net_abstract *>(net));
        if (!n) return 10;
        n->save(path);
        return 0;
    }
    catch (const std::exception &e)
    {
        set_err(e.what());
        return 1;
    }
}

void *nch_load(const char *path, int precision, int device, int max_batch, size_t *n_in, size_t *n_out)
{
    try
    {
        cuda::net_cuda *n = new cuda::net_cuda(cuda::net_cuda::load(path, make_opt(precision, device, 0, max_batch)));
        *n_in = n->n_in(), *n_out = n->n_out();
        return static_cast<net::net_abstract *>(n);
    }
    catch (const std::exception &e)
    {
        set_err(e.what());
        return nullptr;
    }
}

void nch_destroy(void *net) { delete static_cast<net::net_abstract *>(net); }

// launch_forward through the vtable; returns the number of outputs, or -1 on error.
long long nch_launch_forward(void *net, const float *in, size_t n_in_floats, float *out, size_t out_capacity)
{
    try
    {
        net::net_abstract *n = static_cast<net::net_abstract *>(net);
        std::vector<float> x(in, in + n_in_floats);
        std::vector<float> y = n->launch_forward(x);
        // (x dies with this call: a net that page-locks its inputs -- net_cuda_options::pin_inputs -- must forget it first)
        if (cuda::net_cuda *c = dynamic_cast<cuda::net_cuda *>(n)) c->release_inputs();
        if (y.size() > out_capacity)
        {
            set_err("output buffer too small");
            return -1;
        }
        memcpy(out, y.data(), y.size() * sizeof(float));
        return (long long)y.size();
    }
    catch (const std::exception &e)
    {
        set_err(e.what());
        return -1;
    }
}

// Seconds per launch_forward call as a C++ application sees it: the input std::vector exists before the clock starts (the
// application owns it), every call returns its outputs by value through the vtable.  -1 on error.
double nch_time_launch_forward(void *net, const float *in, size_t n_in_floats, int reps, float *last_out, size_t out_capacity)
{
    try
    {
        net::net_abstract *n = static_cast<net::net_abstract *>(net);
        const std::vector<float> x(in, in + n_in_floats);
        std::vector<float> y = n->launch_forward(x); // warm-up (staging buffers and the library's output ring are allocated by the first call)
        y = n->launch_forward(x);
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; i++) y = n->launch_forward(x);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / (reps > 0 ? reps : 1);
        if (cuda::net_cuda *c = dynamic_cast<cuda::net_cuda *>(n)) c->release_inputs(); // (x is about to be freed)
        if (last_out && y.size() <= out_capacity) memcpy(last_out, y.data(), y.size() * sizeof(float));
        return s;
    }
    catch (const std::exception &e)
    {
        set_err(e.what());
        return -1.0;
    }
}

// launch_forward(const net::image_set&): one u8 frame through a ViT.  Returns the number of outputs or -1.
long long nch_launch_forward_frame(void *net, const unsigned char *frame, size_t n_bytes, size_t h, size_t w, float *out, size_t out_capacity)
{
    try
    {
        cuda::net_cuda *n = dynamic_cast<cuda::net_cuda *>(static_cast<net::net_abstract *>(net));
        if (!n) return -1;
        net::image_set set;
        set.resized_image_data.assign(frame, frame + n_bytes);
        set.original_x_pos = set.original_y_pos = 0;
        set.original_h = h, set.original_w = w;
        std::vector<float> y = n->launch_forward(set);
        if (y.size() > out_capacity)
        {
            set_err("output buffer too small");
            return -1;
        }
        memcpy(out, y.data(), y.size() * sizeof(float));
        return (long long)y.size();
    }
    catch (const std::exception &e)
    {
        set_err(e.what());
        return -1;
    }
}

// filter_image / get_filtered_image through the vtable: push one single-channel frame; pop the oldest (returns the number of
// pixels written, 0 for an empty ring -- whose header is still reported through h / w -- or -1 on error).
int nch_filter_image(void *net, const unsigned char *pixels, size_t h, size_t w)
{
    try
    {
        net::image_set set;
        set.resized_image_data.assign(pixels, pixels + h * w);
        set.original_x_pos = set.original_y_pos = 0;
        set.original_h = h, set.original_w = w;
        static_cast<net::net_abstract *>(net)->filter_image(set);
        return 0;
    }
    catch (const std::exception &e)
    {
        set_err(e.what());
        return -1;
    }
}

long long nch_get_filtered_image(void *net, unsigned char *out, size_t capacity, size_t *h, size_t *w)
{
    try
    {
        net::image_set img = static_cast<net::net_abstract *>(net)->get_filtered_image();
        *h = img.original_h, *w = img.original_w;
        if (img.resized_image_data.size() > capacity)
        {
            set_err("output buffer too small");
            return -1;
        }
        if (!img.resized_image_data.empty()) memcpy(out, img.resized_image_data.data(), img.resized_image_data.size());
        return (long long)img.resized_image_data.size();
    }
    catch (const std::exception &e)
    {
        set_err(e.what());
        return -1;
    }
}

long nch_forward_us(void *net) { return static_cast<net::net_abstract *>(net)->get_forward_performance(); }
long nch_gradient_us(void *net) { return static_cast<net::net_abstract *>(net)->get_gradient_performance(); }

// get_net_data -> flat layout; returns 0 on success.
int nch_get_net_data(void *net, float *w_out, size_t w_cap, float *b_out, size_t b_cap, size_t *n_ins, size_t *n_layers)
{
    try
    {
        net::net_data d = static_cast<net::net_abstract *>(net)->get_net_data();
        size_t pc = 0, nc = 0;
        for (size_t l = 0; l < d.n_p_l.size(); l++)
            for (size_t j = 0; j < d.params[l].size(); j++)
            {
                for (float v : d.params[l][j])
                {
                    if (pc >= w_cap) return 2;
                    w_out[pc++] = v;
                }
                if (nc >= b_cap) return 2;
                b_out[nc++] = d.bias[l][j];
            }
        *n_ins = d.n_ins;
        *n_layers = d.n_p_l.size();
        return 0;
    }
    catch (const std::exception &e)
    {
        set_err(e.what());
        return 1;
    }
}

// The stub virtuals must behave like the reference's: launch_gradient -> `iterations` zeros
// (src/netFPGA.cpp:578), init_gradient / print_inner_vals no-ops, image getters harmless.
int nch_check_stubs(void *net)
{
    try
    {
        net::net_abstract *n = static_cast<net::net_abstract *>(net);
        net::net_sets sets;
        n->init_gradient(sets);
        std::vector<float> g = n->launch_gradient(5, 0.1f, 0.5f);
        if (g.size() != 5) return 1;
        for (float v : g)
            if (v != 0.0f) return 2;
        n->print_inner_vals();
        if (n->get_gradient_performance() != 0) return 3;
        net::image_set img;
        img.original_h = img.original_w = img.original_x_pos = img.original_y_pos = 0;
        n->filter_image(img);
        net::image_set out = n->get_filtered_image();
        if (!out.resized_image_data.empty()) return 4;
        return 0;
    }
    catch (const std::exception &e)
    {
        set_err(e.what());
        return -1;
    }
}

// Rule-of-five behaviour: move keeps the net usable, copy-assign yields an independent equal net.
// Returns 0 when moved / copied nets reproduce `expect` (n_out floats for the single sample `in`).
int nch_check_move_copy(void *net, const float *in, size_t n_in, const float *expect, size_t n_out)
{
    try
    {
        cuda::net_cuda *src = dynamic_cast<cuda::net_cuda *>(static_cast<net::net_abstract *>(net));
        if (!src) return 10;
        std::vector<float> x(in, in + n_in);
        net::net_data d = src->get_net_data();
        cuda::net_cuda other(d, false, false);
        other = *src; // copy-assign (deep)
        std::vector<float> y1 = other.launch_forward(x);
        cuda::net_cuda moved(std::move(other)); // move-construct
        std::vector<float> y2 = moved.launch_forward(x);
        cuda::net_cuda third(d, false, false);
        third = std::move(moved); // move-assign
        std::vector<float> y3 = third.launch_forward(x);
        if (y1.size() != n_out || y2.size() != n_out || y3.size() != n_out) return 1;
        for (size_t i = 0; i < n_out; i++)
            if (y1[i] != expect[i] || y2[i] != expect[i] || y3[i] != expect[i]) return 2;
        return 0;
    }
    catch (const std::exception &e)
    {
        set_err(e.what());
        return -1;
    }
}

} // extern "C"

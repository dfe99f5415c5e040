// ref_driver.cpp -- C entry points over the reference's own fpga::net_fpga.  TEST INFRASTRUCTURE ONLY.
//
// Compiled together with the unmodified /root/reference/src/netFPGA.cpp (never copied into this
// repository) and the OpenCL shim into oracle/_ref/libnetfpga_ref.so by oracle/Makefile.  Python
// tests and bench.py's reference arm drive the reference class through these functions exactly
// the way its absent host application would: construct from net_data, call launch_forward once
// per sample (the reference is not batched, src/netFPGA.cpp:266-277).
#include <netFPGA.h>

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

// The reference keeps its device state in namespace-scope globals that cleanup() releases but
// never nulls, and caches the identity of the last-uploaded net (src/netFPGA.cpp:21-45, 254, 639-651;
// SURVEY.md App. A).  Destroying one net and creating another would double-release them or skip
// the weight upload, so the driver resets them after each teardown.
namespace fpga
{
extern cl_kernel g_kernel;
extern cl_program g_program;
extern cl_command_queue g_queue;
extern cl_context g_context;
extern cl_event g_init_event, g_finish_event;
extern int g_n_ins_buff, g_n_layers_buff;
extern int *g_n_p_l_buff;
extern DATA_TYPE *g_params_buff;
}

namespace
{
void reset_reference_globals()
{
    fpga::g_kernel = NULL;
    fpga::g_program = NULL;
    fpga::g_queue = NULL;
    fpga::g_context = NULL;
    fpga::g_init_event = NULL;
    fpga::g_finish_event = NULL;
    fpga::g_n_ins_buff = 0;
    fpga::g_n_layers_buff = 0;
    fpga::g_n_p_l_buff = NULL;
    fpga::g_params_buff = NULL;
}

net::net_data make_data(const int *npl, int n_layers, int n_ins, const float *w, const float *b)
{
    net::net_data d;
    d.n_ins = n_ins;
    d.n_layers = n_layers;
    int fan_in = n_ins;
    for (int l = 0; l < n_layers; l++)
    {
        d.n_p_l.push_back(npl[l]);
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
        fan_in = npl[l];
    }
    return d;
}
}

extern "C" {

// The reference never initialises net_fpga_counter / program_init / forward_kernel_init
// (include/netFPGA.h:39-41); constructing into zeroed storage gives them the
// values the author evidently assumed.
void *ref_net_create(const int *npl, int n_layers, int n_ins, const float *w_flat, const float *b_flat, int random,
                     unsigned seed)
{
    try
    {
        net::net_data d = make_data(npl, n_layers, n_ins, random ? nullptr : w_flat, random ? nullptr : b_flat);
        void *mem = calloc(1, sizeof(fpga::net_fpga));
        if (random) srand(seed);
        return new (mem) fpga::net_fpga(d, false, random != 0);
    }
    catch (...)
    {
        return nullptr;
    }
}

void ref_net_destroy(void *h)
{
    if (!h) return;
    fpga::net_fpga *n = static_cast<fpga::net_fpga *>(h);
    n->~net_fpga();
    free(h);
    reset_reference_globals();
}

int ref_net_sizes(void *h, int *n_params, int *n_neurons, int *n_out)
{
    fpga::net_fpga *n = static_cast<fpga::net_fpga *>(h);
    *n_params = n->n_params;
    *n_neurons = n->n_neurons;
    *n_out = n->n_p_l[n->n_layers - 1];
    return 0;
}

// The flat arrays the reference would upload (src/netFPGA.cpp:506-508).
int ref_net_flat(void *h, float *w_out, float *b_out)
{
    fpga::net_fpga *n = static_cast<fpga::net_fpga *>(h);
    memcpy(w_out, n->params, sizeof(float) * (size_t)n->n_params);
    memcpy(b_out, n->bias, sizeof(float) * (size_t)n->n_neurons);
    return 0;
}

// One launch_forward per sample, through the abstract interface.
int ref_net_forward(void *h, const float *in, size_t batch, float *out)
{
    try
    {
        net::net_abstract *n = static_cast<fpga::net_fpga *>(h);
        fpga::net_fpga *f = static_cast<fpga::net_fpga *>(h);
        const size_t n_in = (size_t)f->n_ins, n_out = (size_t)f->n_p_l[f->n_layers - 1];
        std::vector<float> x(n_in);
        for (size_t s = 0; s < batch; s++)
        {
            memcpy(x.data(), in + s * n_in, sizeof(float) * n_in);
            std::vector<float> y = n->launch_forward(x);
            if (y.size() != n_out) return 2;
            memcpy(out + s * n_out, y.data(), sizeof(float) * n_out);
        }
        return 0;
    }
    catch (...)
    {
        return 1;
    }
}

long ref_net_forward_us(void *h) { return static_cast<fpga::net_fpga *>(h)->get_forward_performance(); }

} // extern "C"

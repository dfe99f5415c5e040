// net_cuda.cpp -- cuda::net_cuda: the net::net_abstract implementation on top of the C ABI.
//
// Mirrors fpga::net_fpga member by member (reference src/netFPGA.cpp):
//   constructor / flatten      :58-109   -> net_cuda(data, derivate, random)
//   move / copy semantics      :111-204  -> rule-of-five below (without the reference's defects)
//   get_net_data               :206-237  -> get_net_data (correct inverse of the flatten)
//   launch_forward             :239-290  -> launch_forward (batched, validated)
//   gradient stubs             :518-601  -> same observable behaviour (no-op / zeros)
//   get_forward_performance    :603-611  -> microseconds of the last forward, transfers included
//   destructor / cleanup       :613-651  -> ~net_cuda (per-instance, no globals)
// This file contains no CUDA: it only calls the extern "C" functions of include/netcuda.h.
#include <netCUDA.h>
#include <netcuda.h>

#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace cuda
{
    struct net_cuda::impl
    {
        netcuda_t *h = nullptr;
        std::vector<netcuda_t *> replicas; // the same net on devices device + 1 ... (net_cuda_options::n_devices)
        long long last_sharded_us = -1;    // wall time of the last multi-GPU forward (-1: the last forward used one GPU)
        bool is_vit = false;
        net_cuda_options opt;
        // MLP host copy in the reference's flat layout (src/netFPGA.cpp:78-107)
        std::size_t n_ins = 0;
        std::vector<int32_t> n_p_l;
        std::vector<float> params, bias;
        // ViT host copy
        vit_data vit;
        signed long gradient_performance = 0;
        netcuda_ring_t *ring = nullptr; // image side channel (filter_image / get_filtered_image), created by the first frame
        // net_cuda_options::pin_inputs: host buffers page-locked on first sight (start address, bytes), oldest first
        std::vector<std::pair<const void *, std::size_t>> pinned_inputs;

        void release_inputs()
        {
            for (auto &r : pinned_inputs) netcuda_host_unregister(r.first);
            pinned_inputs.clear();
        }
        // Page-lock [p, p + bytes) unless it already is; a buffer that moved or changed size is registered anew.
        void pin_input(const void *p, std::size_t bytes)
        {
            for (std::size_t i = 0; i < pinned_inputs.size(); i++)
                if (pinned_inputs[i].first == p)
                {
                    if (pinned_inputs[i].second == bytes) return;
                    netcuda_host_unregister(p); // same start, other size: the vector was resized in place
                    pinned_inputs.erase(pinned_inputs.begin() + (std::ptrdiff_t)i);
                    break;
                }
            if (pinned_inputs.size() >= 8)
            {
                netcuda_host_unregister(pinned_inputs.front().first);
                pinned_inputs.erase(pinned_inputs.begin());
            }
            // (a range that cannot be page-locked -- overlapping another registration, locked-memory limit -- is simply staged)
            if (netcuda_host_register(p, bytes) == NETCUDA_OK) pinned_inputs.emplace_back(p, bytes);
        }

        ~impl()
        {
            release_inputs();
            if (ring) netcuda_ring_destroy(ring);
            for (netcuda_t *r : replicas)
                if (r) netcuda_destroy(r);
            if (h) netcuda_destroy(h);
        }
        // Batched forward over every GPU of this net: contiguous slices, one host thread per extra GPU (slice 0 runs on the caller's).
        template <class In, class Call>
        void forward_sharded(const In *inputs, std::size_t in_per_sample, std::size_t batch, DATA_TYPE *outputs, std::size_t out_per_sample, Call call);
        void instantiate(); // build the device net from the host copies above
    };

    namespace
    {
        [[noreturn]] void throw_last(const char *what, int rc)
        {
            std::string msg = std::string(what) + ": " + netcuda_last_error();
            if (rc == NETCUDA_ERR_INVALID) throw std::invalid_argument(msg);
            throw std::runtime_error(msg);
        }

        net_cuda_options resolve(net_cuda_options o, int default_precision)
        {
            if (o.precision < 0)
            {
                o.precision = default_precision;
                if (const char *e = std::getenv("NETCUDA_PRECISION"))
                {
                    const std::string s(e);
                    if (s == "fp32") o.precision = PREC_FP32;
                    else if (s == "tf32") o.precision = PREC_TF32;
                    else if (s == "bf16") o.precision = PREC_BF16;
                    else if (s == "int8") o.precision = PREC_INT8;
                    else throw std::invalid_argument("NETCUDA_PRECISION must be fp32, tf32, bf16 or int8");
                }
            }
            if (o.device < 0)
            {
                o.device = 0;
                if (const char *e = std::getenv("NETCUDA_DEVICE")) o.device = std::atoi(e);
            }
            if (o.n_devices < 0)
            {
                o.n_devices = 1;
                if (const char *e = std::getenv("NETCUDA_DEVICES")) o.n_devices = std::atoi(e);
            }
            if (o.n_devices < 1) o.n_devices = 1;
            if (o.pin_inputs < 0)
            {
                o.pin_inputs = 0;
                if (const char *e = std::getenv("NETCUDA_PIN_INPUTS")) o.pin_inputs = std::atoi(e) != 0;
            }
            int visible = 0;
            if (o.n_devices > 1 && netcuda_device_count(&visible) == NETCUDA_OK && o.device + o.n_devices > visible)
                throw std::invalid_argument("net_cuda: n_devices asks for more GPUs than are visible");
            return o;
        }
    }

    net_cuda::net_cuda(const net::net_data &data, bool derivate, bool random) : net_cuda(data, net_cuda_options(), random)
    {
        (void)derivate; // ignored by the reference as well (src/netFPGA.cpp:58)
    }

    net_cuda::net_cuda(const net::net_data &data, const net_cuda_options &options, bool random) : p_(new impl)
    {
        try
        {
            // fp32 by default: the reference's DATA_TYPE is float (def/defines.h:10), and NETCUDA_PREC_FP32 reproduces the CPU oracle's
            // k-ordered fmaf sequence bit for bit.  The tensor-core precisions are opt-in (options.precision / NETCUDA_PRECISION).
            p_->opt = resolve(options, PREC_FP32);
            // depth comes from n_p_l.size(); data.n_layers is informational (src/netFPGA.cpp:59)
            if (data.n_p_l.empty() || data.n_ins == 0) throw std::invalid_argument("net_cuda: empty net description");
            p_->n_ins = data.n_ins;
            std::size_t n_params = 0, n_neurons = 0, fan_in = data.n_ins;
            for (std::size_t l = 0; l < data.n_p_l.size(); l++)
            {
                if (data.n_p_l[l] == 0) throw std::invalid_argument("net_cuda: layer with zero neurons");
                p_->n_p_l.push_back((int32_t)data.n_p_l[l]);
                n_params += data.n_p_l[l] * fan_in;
                n_neurons += data.n_p_l[l];
                fan_in = data.n_p_l[l];
            }
            p_->params.resize(n_params);
            p_->bias.resize(n_neurons);
            if (random)
            {
                // the reference's rule and order: all params, then all biases (src/netFPGA.cpp:82-88)
                for (std::size_t i = 0; i < n_params; i++) p_->params[i] = float(rand() % 200 - 100) / 100;
                for (std::size_t i = 0; i < n_neurons; i++) p_->bias[i] = float(rand() % 200 - 100) / 100;
            }
            else
            {
                if (data.params.size() < data.n_p_l.size() || data.bias.size() < data.n_p_l.size())
                    throw std::invalid_argument("net_cuda: params/bias have fewer layers than n_p_l");
                std::size_t pc = 0, nc = 0;
                fan_in = data.n_ins;
                for (std::size_t l = 0; l < data.n_p_l.size(); l++)
                {
                    if (data.params[l].size() != data.n_p_l[l] || data.bias[l].size() != data.n_p_l[l])
                        throw std::invalid_argument("net_cuda: layer size does not match n_p_l");
                    for (std::size_t j = 0; j < data.n_p_l[l]; j++)
                    {
                        if (data.params[l][j].size() != fan_in) throw std::invalid_argument("net_cuda: neuron fan-in mismatch");
                        std::memcpy(&p_->params[pc], data.params[l][j].data(), fan_in * sizeof(float));
                        pc += fan_in;
                        p_->bias[nc++] = data.bias[l][j];
                    }
                    fan_in = data.n_p_l[l];
                }
            }
            p_->instantiate();
        }
        catch (...)
        {
            delete p_;
            throw;
        }
    }

    net_cuda::net_cuda(const vit_data &vit, const net_cuda_options &options) : p_(new impl)
    {
        try
        {
            p_->opt = resolve(options, PREC_BF16);
            p_->is_vit = true;
            p_->vit = vit;
            p_->instantiate();
        }
        catch (...)
        {
            delete p_;
            throw;
        }
    }

    void net_cuda::impl::instantiate()
    {
        impl *p = this;
        auto build = [p](int device) {
            netcuda_t *h = nullptr;
            netcuda_desc d;
            std::memset(&d, 0, sizeof(d));
            d.precision = p->opt.precision;
            d.device = device;
            d.activation = p->opt.activation;
            d.max_batch = p->opt.max_batch;
            int rc;
            if (p->is_vit)
            {
                d.kind = NETCUDA_KIND_VIT;
                d.image_size = (int32_t)p->vit.image_size, d.patch_size = (int32_t)p->vit.patch_size;
                d.dim = (int32_t)p->vit.dim, d.depth = (int32_t)p->vit.depth, d.heads = (int32_t)p->vit.heads;
                d.mlp_dim = (int32_t)p->vit.mlp_dim, d.n_classes = (int32_t)p->vit.n_classes;
                if ((rc = netcuda_create(&d, &h)) != NETCUDA_OK) throw_last("netcuda_create", rc);
                if ((rc = netcuda_upload_vit(h, p->vit.params.data(), p->vit.params.size())) != NETCUDA_OK)
                {
                    netcuda_destroy(h);
                    throw_last("netcuda_upload_vit", rc);
                }
            }
            else
            {
                d.kind = NETCUDA_KIND_MLP;
                d.n_ins = (int32_t)p->n_ins;
                d.n_layers = (int32_t)p->n_p_l.size();
                d.n_p_l = p->n_p_l.data();
                if ((rc = netcuda_create(&d, &h)) != NETCUDA_OK) throw_last("netcuda_create", rc);
                if ((rc = netcuda_upload_mlp(h, p->params.data(), p->bias.data())) != NETCUDA_OK)
                {
                    netcuda_destroy(h);
                    throw_last("netcuda_upload_mlp", rc);
                }
            }
            return h;
        };
        p->h = build(p->opt.device);
        for (int g = 1; g < p->opt.n_devices; g++) p->replicas.push_back(build(p->opt.device + g)); // weights replicated per GPU
    }

    template <class In, class Call>
    void net_cuda::impl::forward_sharded(const In *inputs, std::size_t in_per_sample, std::size_t batch, DATA_TYPE *outputs,
                                         std::size_t out_per_sample, Call call)
    {
        const auto t0 = std::chrono::steady_clock::now();
        const std::size_t gpus = std::min(replicas.size() + 1, batch); // no empty slices
        const std::size_t per = (batch + gpus - 1) / gpus;             // GPU g takes samples [g * per, min(batch, (g + 1) * per))
        std::vector<std::string> errors(gpus);
        std::vector<int> codes(gpus, NETCUDA_OK);
        auto run = [&](std::size_t g) {
            const std::size_t lo = g * per, hi = std::min(batch, lo + per);
            if (lo >= hi) return;
            netcuda_t *hg = g == 0 ? h : replicas[g - 1];
            const auto s0 = std::chrono::steady_clock::now();
            codes[g] = call(hg, inputs + lo * in_per_sample, hi - lo, outputs + lo * out_per_sample);
            if (std::getenv("NETCUDA_SHARD_DEBUG"))
                std::fprintf(stderr, "slice %zu: %zu samples, started %+lld us after the call, took %lld us\n", g, hi - lo,
                             (long long)std::chrono::duration_cast<std::chrono::microseconds>(s0 - t0).count(),
                             (long long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - s0).count());
            if (codes[g] != NETCUDA_OK) errors[g] = netcuda_last_error(); // (the message is thread-local: fetched where it was set)
        };
        std::vector<std::thread> workers;
        for (std::size_t g = 1; g < gpus; g++) workers.emplace_back(run, g);
        run(0);
        for (auto &w : workers) w.join();
        last_sharded_us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        for (std::size_t g = 0; g < gpus; g++)
            if (codes[g] != NETCUDA_OK)
            {
                const std::string msg = "net_cuda: forward on GPU slice " + std::to_string(g) + ": " + errors[g];
                if (codes[g] == NETCUDA_ERR_INVALID) throw std::invalid_argument(msg);
                throw std::runtime_error(msg);
            }
    }

    void net_cuda::save(const char *path) const
    {
        if (!p_) throw std::runtime_error("net_cuda: moved-from net");
        int rc;
        if (p_->is_vit)
        {
            netcuda_desc d;
            std::memset(&d, 0, sizeof(d));
            d.kind = NETCUDA_KIND_VIT, d.precision = NETCUDA_PREC_BF16;
            d.image_size = (int32_t)p_->vit.image_size, d.patch_size = (int32_t)p_->vit.patch_size;
            d.dim = (int32_t)p_->vit.dim, d.depth = (int32_t)p_->vit.depth, d.heads = (int32_t)p_->vit.heads;
            d.mlp_dim = (int32_t)p_->vit.mlp_dim, d.n_classes = (int32_t)p_->vit.n_classes;
            rc = netcuda_file_write_vit(path, &d, p_->vit.params.data(), p_->vit.params.size());
        }
        else
            rc = netcuda_file_write_mlp(path, p_->n_p_l.data(), (int)p_->n_p_l.size(), (int)p_->n_ins, p_->opt.activation,
                                        p_->params.data(), p_->bias.data());
        if (rc != NETCUDA_OK) throw_last("net_cuda::save", rc);
    }

    net_cuda net_cuda::load(const char *path, const net_cuda_options &options)
    {
        netcuda_file_info info;
        int rc = netcuda_file_info_read(path, &info);
        if (rc != NETCUDA_OK) throw_last("net_cuda::load", rc);
        if (info.dtype != NETCUDA_FILE_F32)
            throw std::invalid_argument("net_cuda::load: the file holds Q1.7 integers; net_cuda keeps DATA_TYPE (float) nets -- use netcuda_create_from_file");
        std::vector<float> w((std::size_t)info.n_weights), b((std::size_t)info.n_biases);
        rc = netcuda_file_read(path, w.data(), w.size() * sizeof(float), b.data(), b.size() * sizeof(float));
        if (rc != NETCUDA_OK) throw_last("net_cuda::load", rc);
        if (info.desc.kind == NETCUDA_KIND_VIT)
        {
            vit_data v;
            v.image_size = (std::size_t)info.desc.image_size, v.patch_size = (std::size_t)info.desc.patch_size;
            v.dim = (std::size_t)info.desc.dim, v.depth = (std::size_t)info.desc.depth, v.heads = (std::size_t)info.desc.heads;
            v.mlp_dim = (std::size_t)info.desc.mlp_dim, v.n_classes = (std::size_t)info.desc.n_classes;
            v.params.swap(w);
            return net_cuda(v, options);
        }
        // back to the nested form the reference's constructor takes (def/defines.h:14-23)
        net::net_data data;
        data.n_ins = (std::size_t)info.desc.n_ins;
        data.n_layers = (std::size_t)info.desc.n_layers;
        std::size_t pc = 0, nc = 0, fan_in = data.n_ins;
        for (int l = 0; l < info.desc.n_layers; l++)
        {
            const std::size_t fan_out = (std::size_t)info.n_p_l[l];
            data.n_p_l.push_back(fan_out);
            data.params.emplace_back();
            data.bias.emplace_back();
            for (std::size_t j = 0; j < fan_out; j++)
            {
                if (pc + fan_in > w.size() || nc >= b.size()) throw std::invalid_argument("net_cuda::load: layer table exceeds the file's payload");
                data.params[l].emplace_back(w.begin() + pc, w.begin() + pc + fan_in);
                pc += fan_in;
                data.bias[l].push_back(b[nc++]);
            }
            fan_in = fan_out;
        }
        net_cuda_options o = options;
        o.activation = info.desc.activation;
        return net_cuda(data, o, false);
    }

    net_cuda::~net_cuda() { delete p_; }

    net_cuda::net_cuda(net_cuda &&rh) noexcept : p_(rh.p_) { rh.p_ = nullptr; }

    net_cuda &net_cuda::operator=(net_cuda &&rh) noexcept
    {
        if (this != &rh)
        {
            delete p_;
            p_ = rh.p_;
            rh.p_ = nullptr;
        }
        return *this;
    }

    net_cuda &net_cuda::operator=(const net_cuda &rh)
    {
        if (this == &rh) return *this;
        if (!rh.p_) throw std::invalid_argument("net_cuda: copy from a moved-from net");
        impl *np = new impl;
        try
        {
            np->is_vit = rh.p_->is_vit;
            np->opt = rh.p_->opt;
            np->n_ins = rh.p_->n_ins;
            np->n_p_l = rh.p_->n_p_l;
            np->params = rh.p_->params;
            np->bias = rh.p_->bias;
            np->vit = rh.p_->vit;
            np->instantiate();
        }
        catch (...)
        {
            delete np;
            throw;
        }
        delete p_;
        p_ = np;
        return *this;
    }

    net::net_data net_cuda::get_net_data()
    {
        if (!p_) throw std::runtime_error("net_cuda: moved-from net");
        if (p_->is_vit) throw std::runtime_error("net_cuda::get_net_data: a ViT cannot be expressed as net::net_data");
        net::net_data data;
        data.n_ins = p_->n_ins;
        data.n_layers = p_->n_p_l.size();
        std::size_t pc = 0, nc = 0, fan_in = p_->n_ins;
        for (std::size_t l = 0; l < p_->n_p_l.size(); l++)
        {
            const std::size_t fan_out = (std::size_t)p_->n_p_l[l];
            data.n_p_l.push_back(fan_out);
            data.params.emplace_back();
            data.bias.emplace_back();
            data.params[l].reserve(fan_out);
            for (std::size_t j = 0; j < fan_out; j++)
            {
                data.params[l].emplace_back(p_->params.begin() + pc, p_->params.begin() + pc + fan_in);
                pc += fan_in;
                data.bias[l].push_back(p_->bias[nc++]);
            }
            fan_in = fan_out;
        }
        return data;
    }

    std::size_t net_cuda::n_in() const
    {
        std::size_t n = 0;
        if (p_ && p_->h) netcuda_n_in(p_->h, &n);
        return n;
    }
    std::size_t net_cuda::n_out() const
    {
        std::size_t n = 0;
        if (p_ && p_->h) netcuda_n_out(p_->h, &n);
        return n;
    }
    void *net_cuda::c_handle() const { return p_ ? p_->h : nullptr; }

    void net_cuda::forward(const DATA_TYPE *inputs, std::size_t batch, DATA_TYPE *outputs)
    {
        if (!p_) throw std::runtime_error("net_cuda: moved-from net");
        if (!p_->replicas.empty() && batch > 1)
        {
            p_->forward_sharded(inputs, n_in(), batch, outputs, n_out(),
                                [](netcuda_t *h, const DATA_TYPE *in, std::size_t n, DATA_TYPE *out) { return netcuda_forward(h, in, n, out); });
            return;
        }
        p_->last_sharded_us = -1;
        const int rc = netcuda_forward(p_->h, inputs, batch, outputs);
        if (rc != NETCUDA_OK) throw_last("netcuda_forward", rc);
    }

    void net_cuda::forward_u8(const unsigned char *frames, std::size_t batch, DATA_TYPE *outputs)
    {
        if (!p_) throw std::runtime_error("net_cuda: moved-from net");
        if (!p_->replicas.empty() && batch > 1)
        {
            p_->forward_sharded(frames, n_in(), batch, outputs, n_out(), [](netcuda_t *h, const unsigned char *in, std::size_t n, DATA_TYPE *out) {
                return netcuda_forward_u8(h, in, n, out);
            });
            return;
        }
        p_->last_sharded_us = -1;
        const int rc = netcuda_forward_u8(p_->h, frames, batch, outputs);
        if (rc != NETCUDA_OK) throw_last("netcuda_forward_u8", rc);
    }

    void net_cuda::set_u8_normalization(const float mean[3], const float stddev[3])
    {
        if (!p_) throw std::runtime_error("net_cuda: moved-from net");
        for (netcuda_t *r : p_->replicas)
            if (netcuda_set_u8_normalization(r, mean, stddev) != NETCUDA_OK) throw_last("netcuda_set_u8_normalization", NETCUDA_ERR_CUDA);
        const int rc = netcuda_set_u8_normalization(p_->h, mean, stddev);
        if (rc != NETCUDA_OK) throw_last("netcuda_set_u8_normalization", rc);
    }

    std::vector<DATA_TYPE> net_cuda::launch_forward(const net::image_set &frame)
    {
        if (!p_) throw std::runtime_error("net_cuda: moved-from net");
        if (!p_->is_vit) throw std::invalid_argument("net_cuda::launch_forward(image_set): frames feed vision transformers only");
        const std::size_t side = p_->vit.image_size;
        if (frame.resized_image_data.size() != side * side * 3)
            throw std::invalid_argument("net_cuda::launch_forward(image_set): resized_image_data must hold image_size * image_size * 3 bytes");
        if ((frame.original_h && frame.original_h != side) || (frame.original_w && frame.original_w != side))
            throw std::invalid_argument("net_cuda::launch_forward(image_set): original_h / original_w do not match the net's image size");
        std::vector<DATA_TYPE> out(n_out());
        forward_u8(frame.resized_image_data.data(), 1, out.data());
        return out;
    }

    std::uint64_t net_cuda::submit(const DATA_TYPE *inputs, std::size_t batch, DATA_TYPE *outputs)
    {
        if (!p_) throw std::runtime_error("net_cuda: moved-from net");
        uint64_t ticket = 0;
        const int rc = netcuda_submit(p_->h, inputs, batch, outputs, &ticket);
        if (rc != NETCUDA_OK) throw_last("netcuda_submit", rc);
        return ticket;
    }

    void net_cuda::wait(std::uint64_t ticket)
    {
        if (!p_) throw std::runtime_error("net_cuda: moved-from net");
        const int rc = netcuda_wait(p_->h, ticket);
        if (rc != NETCUDA_OK) throw_last("netcuda_wait", rc);
    }

    void net_cuda::forward_device(const void *d_inputs, std::size_t batch, void *d_outputs, void *stream)
    {
        if (!p_) throw std::runtime_error("net_cuda: moved-from net");
        const int rc = netcuda_forward_device(p_->h, d_inputs, batch, d_outputs, stream);
        if (rc != NETCUDA_OK) throw_last("netcuda_forward_device", rc);
    }

    std::vector<DATA_TYPE> net_cuda::launch_forward(const std::vector<DATA_TYPE> &inputs)
    {
        if (!p_) throw std::runtime_error("net_cuda: moved-from net");
        const std::size_t ni = n_in(), no = n_out();
        if (inputs.empty() || inputs.size() % ni != 0)
            throw std::invalid_argument("net_cuda::launch_forward: inputs.size() must be a non-zero multiple of n_ins");
        const std::size_t batch = inputs.size() / ni;
        std::vector<DATA_TYPE> out(batch * no);
        // (worth it from a few MB on: a small input is staged faster than the driver walks its pages)
        if (p_->opt.pin_inputs && inputs.size() * sizeof(DATA_TYPE) >= (std::size_t)1 << 20) p_->pin_input(inputs.data(), inputs.size() * sizeof(DATA_TYPE));
        forward(inputs.data(), batch, out.data());
        return out;
    }

    void net_cuda::release_inputs()
    {
        if (p_) p_->release_inputs();
    }

    // Training is not implemented by the reference either: init_gradient's body is commented out
    // (src/netFPGA.cpp:518-542) and launch_gradient returns `iterations` zeros (:578).
    void net_cuda::init_gradient(const net::net_sets &sets) { (void)sets; }

    std::vector<DATA_TYPE> net_cuda::launch_gradient(size_t iterations, DATA_TYPE error_threshold, DATA_TYPE multiplier)
    {
        (void)error_threshold;
        (void)multiplier;
        return std::vector<DATA_TYPE>(iterations, 0);
    }

    void net_cuda::print_inner_vals() {}

    signed long net_cuda::get_gradient_performance() { return p_ ? p_->gradient_performance : 0; }

    signed long net_cuda::get_forward_performance()
    {
        int64_t us = 0;
        if (p_ && p_->h) netcuda_last_forward_us(p_->h, &us);
        if (p_ && p_->last_sharded_us >= 0) us = p_->last_sharded_us; // the last forward ran on several GPUs: the whole call
        return (signed long)us;
    }

    // Image side channel (src/netFPGA.cpp:292-365): filter_image enqueues a single-channel frame of original_h * original_w bytes
    // into a ring of 24 slots (BATCH_SIZE, :12) -- copy, H2D, filter kernel, non-blocking D2H -- and returns at once; a full ring
    // drops the frame (the reference prints "PILA LLENA", :333).  get_filtered_image waits for the oldest frame (:349).  The ring
    // (netcuda_ring_*, csrc/frame_ring.cu) is created by the first frame and sized from it, like the reference's
    // _init_kernel("image_process", set) (:443-458), but at least 1920 x 1080 (IMAGE_WIDTH x IMAGE_HEIGHT, include/netFPGA.h:14-15).
    void net_cuda::filter_image(const net::image_set &set)
    {
        if (!p_) throw std::runtime_error("net_cuda: moved-from net");
        const std::size_t pixels = set.original_h * set.original_w;
        if (pixels == 0 || set.resized_image_data.size() < pixels) return; // nothing to filter (the reference would read out of bounds)
        if (p_->ring == nullptr)
        {
            const int rc = netcuda_ring_create(p_->opt.device, NETCUDA_RING_DEPTH, std::max<std::size_t>(pixels, 1920u * 1080u), &p_->ring);
            if (rc != NETCUDA_OK) throw_last("netcuda_ring_create", rc);
        }
        const int rc = netcuda_ring_push(p_->ring, set.resized_image_data.data(), set.original_h, set.original_w);
        if (rc != NETCUDA_OK && rc != NETCUDA_ERR_RING_FULL) throw_last("netcuda_ring_push", rc); // full: the frame is dropped, as in the reference
    }

    net::image_set net_cuda::get_filtered_image()
    {
        if (!p_) throw std::runtime_error("net_cuda: moved-from net");
        net::image_set out;
        out.original_x_pos = 0;
        out.original_y_pos = 0;
        out.original_h = 1080; // IMAGE_HEIGHT / IMAGE_WIDTH: what the reference answers for an empty ring (:336-365)
        out.original_w = 1920;
        std::size_t h = 0, w = 0;
        if (p_->ring == nullptr || netcuda_ring_peek(p_->ring, &h, &w) != NETCUDA_OK) return out; // "PILA VACIA": a header, no pixels
        out.resized_image_data.resize(h * w);
        const int rc = netcuda_ring_pop(p_->ring, out.resized_image_data.data(), out.resized_image_data.size(), &h, &w);
        if (rc != NETCUDA_OK) throw_last("netcuda_ring_pop", rc);
        out.original_h = h, out.original_w = w;
        return out;
    }
}

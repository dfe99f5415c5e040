// netCUDA.h -- cuda::net_cuda, the B200 backend behind net::net_abstract.
//
// Drop-in counterpart of the reference's fpga::net_fpga (include/netFPGA.h:17-71): same base
// class, same constructor shape `(const net::net_data&, bool derivate, bool random)`
// (include/netFPGA.h:54), deleted default constructor, move construction / move assignment /
// copy assignment (include/netFPGA.h:53-57).  A host application switches backend by replacing
//     #include <netFPGA.h>   fpga::net_fpga net(data, false, false);
// with
//     #include <netCUDA.h>   cuda::net_cuda net(data, false, false);
// and linking libnetcuda_host.so + libnetcuda.so instead of libnetFPGA.a.
//
// Deliberate differences from net_fpga (SURVEY.md App. A lists the reference defects):
//   * no public data members and no process-global device state: everything is per instance;
//   * launch_forward accepts B * n_ins inputs and returns B * n_out outputs (B = 1 is the
//     reference contract), validates sizes and throws instead of reading out of bounds / exit()ing;
//   * get_net_data() really is the inverse of the constructor's flatten;
//   * a second family of nets (vision transformers) can be built from cuda::vit_data;
//   * one object can drive several GPUs of the box (net_cuda_options::n_devices / NETCUDA_DEVICES): the weights are replicated,
//     every batched forward is cut into contiguous slices, one per GPU, each fed by its own host thread -- samples are independent
//     (src/netFPGA.cpp:266-277: one sample in, one out), so the outputs are the single-GPU bits.
#ifndef NETCUDA_CLASS_H
#define NETCUDA_CLASS_H

#include <netAbstract.h>

#include <cstddef>
#include <cstdint>
#include <vector>

namespace cuda
{
    // Values mirror NETCUDA_PREC_* / NETCUDA_ACT_* in netcuda.h.
    enum precision_t
    {
        PREC_FP32 = 0, // CUDA-core fp32, bit-equal to the CPU oracle: the default for nets built from net::net_data (DATA_TYPE is float)
        PREC_TF32 = 1, // tcgen05 kind::tf32 (opt-in for MLPs: ~1e-3 relative; vision transformers: every linear layer in tf32)
        PREC_BF16 = 2, // default precision of vision transformers
        PREC_INT8 = 3  // Q1.7 fixed point, bit-exact
    };
    enum activation_t
    {
        ACT_RELU_HIDDEN = 0,
        ACT_RELU_ALL = 1,
        ACT_NONE = 2
    };

    struct net_cuda_options
    {
        int precision;  // precision_t; -1 = take NETCUDA_PRECISION from the environment, else FP32 (MLP) / BF16 (ViT)
        int device;     // CUDA ordinal; -1 = NETCUDA_DEVICE from the environment, else 0
        int activation; // activation_t
        int max_batch;  // samples per internal pass, 0 = library default
        int n_devices;  // GPUs [device, device + n_devices) share every batched forward; -1 = NETCUDA_DEVICES from the environment, else 1
        // 1: page-lock the vector a launch_forward(std::vector) call reads, the first time its buffer is seen (up to 8 buffers are
        // remembered), so that every later call with the same buffer is DMA'd straight from it -- no staging copy.  The caller
        // promises that such a buffer is not freed or reallocated while the net lives (release_inputs() forgets them earlier).
        // 0: every call stages its input through the library's pinned slots.  -1 = NETCUDA_PIN_INPUTS from the environment, else 0.
        int pin_inputs;
        net_cuda_options() : precision(-1), device(-1), activation(ACT_RELU_HIDDEN), max_batch(0), n_devices(-1), pin_inputs(-1) {}
    };

    // Vision-transformer description (net::net_data can only express an MLP, def/defines.h:14-23).
    // `params` is the flat fp32 vector documented at netcuda_vit_param_count (netcuda.h).
    struct vit_data
    {
        std::size_t image_size, patch_size, dim, depth, heads, mlp_dim, n_classes;
        std::vector<DATA_TYPE> params;
    };

    class net_cuda : public net::net_abstract
    {
    private:
        net_cuda() = delete;
        struct impl;
        impl *p_;

    public:
        ~net_cuda();
        net_cuda(const net::net_data &data, bool derivate, bool random); // same shape as net_fpga's
        net_cuda(const net::net_data &data, const net_cuda_options &options, bool random = false);
        net_cuda(const vit_data &vit, const net_cuda_options &options = net_cuda_options());
        net_cuda(net_cuda &&rh) noexcept;
        net_cuda &operator=(net_cuda &&rh) noexcept;
        net_cuda &operator=(const net_cuda &rh);

        net::net_data get_net_data() override;
        std::vector<DATA_TYPE> launch_forward(const std::vector<DATA_TYPE> &inputs) override;
        void init_gradient(const net::net_sets &sets) override;
        std::vector<DATA_TYPE> launch_gradient(size_t iterations, DATA_TYPE error_threshold, DATA_TYPE multiplier) override;
        void print_inner_vals() override;
        signed long get_gradient_performance() override;
        signed long get_forward_performance() override;
        void filter_image(const net::image_set &set) override;
        net::image_set get_filtered_image() override;

        // ---- weight files (netcuda.h, "weight files"): the flat layout of src/netFPGA.cpp:91-106 on disk ----
        // Writes this net (the fp32 values it was constructed from) to `path`.
        void save(const char *path) const;
        // Builds the net a file describes.  options.precision < 0 = the file's natural precision
        // (FP32 for an fp32 MLP file unless NETCUDA_PRECISION says otherwise, BF16 for a ViT).
        static net_cuda load(const char *path, const net_cuda_options &options = net_cuda_options());

        // ---- extensions beyond the abstract interface ----
        // Batched forward on raw host buffers (pinned buffers are DMA'd in place).
        void forward(const DATA_TYPE *inputs, std::size_t batch, DATA_TYPE *outputs);
        // Forget (and un-page-lock) the input buffers remembered under net_cuda_options::pin_inputs: call it before freeing one.
        void release_inputs();
        // Vision transformers fed from camera frames, the reference's image carrier (net::image_set, def/defines.h:31-38):
        // resized_image_data holds H x W x 3 interleaved bytes of one frame (original_h / original_w are checked against the
        // net's image size when non-zero).  Normalisation v = (u8 / 255 - mean) / stddev on the GPU; default [-1, 1].
        std::vector<DATA_TYPE> launch_forward(const net::image_set &frame);
        void forward_u8(const unsigned char *frames, std::size_t batch, DATA_TYPE *outputs);
        void set_u8_normalization(const float mean[3], const float stddev[3]);
        // Non-blocking forward (netcuda_submit / netcuda_wait): returns a ticket; the buffers must stay valid (and should be
        // page-locked) until wait(ticket) returns.  Up to 4 calls in flight; the copy of call i+1 overlaps the kernels of call i.
        std::uint64_t submit(const DATA_TYPE *inputs, std::size_t batch, DATA_TYPE *outputs);
        void wait(std::uint64_t ticket);
        // Device-resident forward on `stream` (cudaStream_t as void*; nullptr = the net's stream).
        void forward_device(const void *d_inputs, std::size_t batch, void *d_outputs, void *stream = nullptr);
        std::size_t n_in() const;
        std::size_t n_out() const;
        void *c_handle() const; // the underlying netcuda_t*
    };
}

#endif

// netAbstract.h -- the plug-in interface every net backend implements.
//
// Restates VIT-FPGA's include/netAbstract.h:8-21.  A backend is a class deriving from
// net::net_abstract; the host application only ever talks to a `net::net_abstract*`.
// The reference's one backend is fpga::net_fpga (include/netFPGA.h:17); this repository adds
// cuda::net_cuda (include/netCUDA.h) behind the very same vtable.
//
// ABI contract (checked by tests/test_boundary.py against the reference header and against the
// relocation order in the reference's prebuilt netFPGA.o):
//   vtable slot 0/1  virtual destructor (complete / deleting)
//   slot 2  get_net_data              slot 7  get_gradient_performance
//   slot 3  launch_forward            slot 8  get_forward_performance
//   slot 4  init_gradient             slot 9  filter_image
//   slot 5  launch_gradient           slot 10 get_filtered_image
//   slot 6  print_inner_vals
// Do NOT reorder, add or remove virtuals here: doing so silently breaks every binary that was
// compiled against the reference header.
#ifndef NETABSTRACT_H
#define NETABSTRACT_H

#include <defines.h>

namespace net
{
    class net_abstract
    {
    public:
        virtual ~net_abstract() {}

        // Export the weights back in nested-vector form (inverse of the constructor's flatten).
        virtual net_data get_net_data() = 0;

        // Forward pass.  Reference contract: `inputs` holds n_ins values, the result holds
        // n_p_l.back() values (src/netFPGA.cpp:266-289).  net_cuda extends this compatibly:
        // inputs.size() == B * n_ins  =>  B * n_out results, sample-major.
        virtual std::vector<DATA_TYPE> launch_forward(const std::vector<DATA_TYPE> &inputs) = 0;

        // Training hooks -- stubs in the reference (src/netFPGA.cpp:518-580).
        virtual void init_gradient(const net_sets &sets) = 0;
        virtual std::vector<DATA_TYPE> launch_gradient(size_t iterations, DATA_TYPE error_threshold, DATA_TYPE multiplier) = 0;
        virtual void print_inner_vals() = 0;

        // Wall-clock microseconds of the last gradient / forward call (PERFORMANCE builds).
        virtual signed long get_gradient_performance() = 0;
        virtual signed long get_forward_performance() = 0;

        // Image-filter side channel (src/netFPGA.cpp:292-365); not part of the net hot path.
        virtual void filter_image(const image_set &set) = 0;
        virtual image_set get_filtered_image() = 0;
    };
}
#endif

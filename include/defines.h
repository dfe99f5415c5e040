// defines.h -- data model shared by every backend that plugs into net::net_abstract.
//
// This header is the *contract* half of the drop-in boundary: a host application that was
// written against VIT-FPGA's `def/defines.h` (reference: def/defines.h:6-39) must compile and
// link against this file unchanged.  It is therefore a restatement, not an extension:
//
//   * same namespace (`net`), same type names, same member names, same member ORDER and the
//     same member types, so `sizeof`/`offsetof` and the Itanium-mangled names of every function
//     taking these types (`N3net8net_dataE`, ...) are identical.  `tests/test_boundary.py`
//     checks that, against the reference header when `/root/reference` is present;
//   * same three macros (`ASSERT`, `PERFORMANCE`, `DATA_TYPE`) and the same two range constants.
//
// The only deliberate difference: the reference forgets to include <cstddef> and only builds
// because older standard libraries leaked `size_t` out of <vector>.
//
// What the fields mean (taken from how src/netFPGA.cpp:58-107 consumes them):
//   n_ins      number of input features of the first layer
//   n_layers   informational only -- the reference derives the depth from n_p_l.size() (:59)
//   n_p_l      "neurons per layer": fan-out of every layer, hidden layers first, output last
//   params     params[layer][neuron][k] = weight from input k of that layer to `neuron`
//              (so each layer is a row-major W[out][in] matrix)
//   bias       bias[layer][neuron]
//   activations declared by the reference but never read ("TODO: IMPLEMENTAR", defines.h:21-22)
#ifndef DEFINES_H
#define DEFINES_H

#include <cstddef>
#include <vector>

namespace net
{
#define ASSERT
#define PERFORMANCE
#define DATA_TYPE float

    constexpr DATA_TYPE MAX_RANGE = 1;
    constexpr DATA_TYPE MIN_RANGE = -1;

    // Unnamed-struct-plus-typedef on purpose: the typedef name is the name "for linkage
    // purposes", exactly as in the reference, so mangled symbols agree.
    typedef struct
    {
        std::size_t n_ins;
        std::size_t n_layers;
        std::vector<std::size_t> n_p_l;
        std::vector<std::vector<std::vector<DATA_TYPE>>> params;
        std::vector<std::vector<DATA_TYPE>> bias;
        std::vector<std::vector<DATA_TYPE>> activations;
    } net_data;

    // Training sets: only needed to spell the signature of init_gradient (a stub in the
    // reference, src/netFPGA.cpp:518-542, and a stub here).
    typedef struct
    {
        std::vector<std::vector<DATA_TYPE>> set_ins;
        std::vector<std::vector<DATA_TYPE>> set_outs;
    } net_sets;

    // One frame for the image-filter side channel (reference: src/netFPGA.cpp:292-365).
    // Out of scope for the CUDA backend; kept so the vtable signatures match.
    typedef struct
    {
        std::vector<unsigned char> resized_image_data;
        std::size_t original_x_pos;
        std::size_t original_y_pos;
        std::size_t original_h;
        std::size_t original_w;
    } image_set;
}
#endif

// drop_in_app.cpp -- a host application written against net::net_abstract only, the way VIT-FPGA's consumer is
// (it includes the backend header, constructs the backend class and from then on talks to net::net_abstract*;
// reference: install_VIT_FPGA.sh:8, Makefile:75-76,94-95, include/netAbstract.h:8-21).
//
//   g++ -std=gnu++14 -O2 -Iinclude examples/drop_in_app.cpp -Lvit-fpga_b200/lib -lnetcuda_host -lnetcuda
//       -Wl,-rpath,$PWD/vit-fpga_b200/lib -o drop_in_app
//
// Switching backend is the two lines under BACKEND below.  The net is config C1 of BASELINE.json (784-128-64-10) with the
// reference's own random initialisation (`random = true`: float(rand() % 200 - 100) / 100, src/netFPGA.cpp:82-88, after srand(1)).
// Prints, for a fixed input, the 10 outputs of sample 0 with full precision (tests/test_boundary.py compares them with the CPU
// oracle: bit-equal under NETCUDA_PRECISION=fp32), then the per-call time of one-sample calls and of one 64-sample call.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

// ---- BACKEND -------------------------------------------------------------------------------------------------------------
#ifdef USE_NETFPGA
#include <netFPGA.h> // the reference:  fpga::net_fpga
typedef fpga::net_fpga backend_t;
#else
#include <netCUDA.h> // this repository: cuda::net_cuda
typedef cuda::net_cuda backend_t;
#endif
// ---------------------------------------------------------------------------------------------------------------------------

int main()
{
    net::net_data data; // def/defines.h:14-23
    data.n_ins = 784;
    data.n_layers = 3;
    data.n_p_l = {128, 64, 10};
    std::srand(1);
    net::net_abstract *net = new backend_t(data, /*derivate=*/false, /*random=*/true); // include/netFPGA.h:54

    // inputs in [-1, 1) (MIN_RANGE / MAX_RANGE, def/defines.h:11-12): a multiplicative hash of the index, exact in float
    const size_t batch = 64;
    std::vector<DATA_TYPE> all(batch * data.n_ins);
    for (size_t i = 0; i < all.size(); i++) all[i] = (float)((((unsigned)i * 2654435761u) >> 8) & 0xFFFFu) / 32768.0f - 1.0f;

    // one sample per call: the reference's contract (src/netFPGA.cpp:266-289)
    std::vector<DATA_TYPE> sample(all.begin(), all.begin() + data.n_ins);
    std::vector<DATA_TYPE> out = net->launch_forward(sample);
    std::printf("outputs");
    for (DATA_TYPE v : out) std::printf(" %.9g", (double)v);
    std::printf("\n");

    const int reps = 200;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) out = net->launch_forward(sample);
    const double us_one = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
    std::printf("one sample per call: %.1f us per call (get_forward_performance: %ld us)\n", us_one, net->get_forward_performance());

#ifndef USE_NETFPGA
    // batched extension of net_cuda: B * n_ins in, B * n_out out
    t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) out = net->launch_forward(all);
    const double us_batch = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
    std::printf("64 samples per call: %.1f us per call, %zu outputs\n", us_batch, out.size());
#endif
    delete net;
    return 0;
}

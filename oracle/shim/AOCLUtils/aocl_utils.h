// Stand-in for Intel's AOCLUtils host helpers -- TEST INFRASTRUCTURE ONLY.
//
// The reference links Intel's prebuilt helper objects (opencl.o / options.o, source not in the
// repo).  src/netFPGA.cpp uses three of their functions plus the checkError macro; they are
// declared here and implemented in oracle/shim/cl_shim.cpp.  Unlike Intel's version the shim's
// _checkError throws instead of exit()ing so a test process survives a failure.
#ifndef NETCUDA_SHIM_AOCL_UTILS_H
#define NETCUDA_SHIM_AOCL_UTILS_H

#include "CL/cl.hpp"
#include <string>

// Intel's library requires the client to provide this (src/netFPGA.cpp:639).
void cleanup();

namespace aocl_utils
{
void _checkError(int line, const char *file, cl_int error, const char *msg, ...);
#define checkError(status, ...) aocl_utils::_checkError(__LINE__, __FILE__, status, __VA_ARGS__)

std::string getBoardBinaryFile(const char *prefix, cl_device_id device);
cl_program createProgramFromBinary(cl_context context, const char *binary_file_name, const cl_device_id *devices,
                                   unsigned num_devices);
}
#endif

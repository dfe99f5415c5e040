// cl_shim.cpp -- in-process fake OpenCL device + AOCLUtils helpers.  TEST INFRASTRUCTURE ONLY.
//
// Purpose: let the UNMODIFIED reference host runtime (/root/reference/src/netFPGA.cpp) execute on
// the CPU, so that (a) the oracle's layout / I/O contract is pinned against the reference's own
// code and (b) bench.py has a faithful "reference path on host cores" baseline.  The device
// kernel the reference would run (`network_v1`, an FPGA bitstream that is not in the repo,
// src/netFPGA.cpp:250,388-390) is supplied by oracle_mlp_forward_one.
//
// Semantics: one in-order queue executed inline, so every enqueue completes before it returns and
// events are inert tokens.  Kernel arguments follow src/netFPGA.cpp:427-436 and :499-502:
//   0 inputs  1 params  2 bias  3 outs  4 npl (cl_mem)   5 n_layers  6 n_ins (cl_int by value)
#include "AOCLUtils/aocl_utils.h"
#include "../oracle.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

struct _cl_platform_id { int tag; };
struct _cl_device_id { int tag; };
struct _cl_context { int refs; };
struct _cl_command_queue { int refs; };
struct _cl_program { int built; };
struct _cl_event { int status; };
struct _cl_mem { unsigned char *data; size_t size; };
struct _cl_kernel
{
    std::string name;
    unsigned char arg[8][8];
    size_t arg_size[8];
};

static _cl_platform_id g_the_platform = {1};
static _cl_device_id g_the_device = {1};
static int g_shim_activation = ORACLE_ACT_RELU_HIDDEN;
static unsigned long g_shim_tasks = 0;

extern "C" void shim_set_activation(int act) { g_shim_activation = act; }
extern "C" unsigned long shim_task_count(void) { return g_shim_tasks; }

static cl_event new_event()
{
    _cl_event *e = new _cl_event;
    e->status = CL_COMPLETE;
    return e;
}

extern "C" {

cl_int clGetPlatformIDs(cl_uint n, cl_platform_id *platforms, cl_uint *num)
{
    if (n && platforms) platforms[0] = &g_the_platform;
    if (num) *num = 1;
    return CL_SUCCESS;
}

cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint n, cl_device_id *devices, cl_uint *num)
{
    if (n && devices) devices[0] = &g_the_device;
    if (num) *num = 1;
    return CL_SUCCESS;
}

cl_context clCreateContext(const cl_context_properties *, cl_uint, const cl_device_id *,
                           void (*)(const char *, const void *, size_t, void *), void *, cl_int *err)
{
    if (err) *err = CL_SUCCESS;
    return new _cl_context{1};
}

cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int *err)
{
    if (err) *err = CL_SUCCESS;
    return new _cl_command_queue{1};
}

cl_int clBuildProgram(cl_program p, cl_uint, const cl_device_id *, const char *, void (*)(cl_program, void *), void *)
{
    if (p) p->built = 1;
    return CL_SUCCESS;
}

cl_event clCreateUserEvent(cl_context, cl_int *err)
{
    if (err) *err = CL_SUCCESS;
    return new_event();
}

cl_int clSetUserEventStatus(cl_event e, cl_int status)
{
    if (e) e->status = status;
    return CL_SUCCESS;
}

cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t size, void *, cl_int *err)
{
    _cl_mem *m = new _cl_mem;
    m->size = size;
    m->data = (unsigned char *)calloc(size ? size : 1, 1);
    if (err) *err = CL_SUCCESS;
    return m;
}

cl_kernel clCreateKernel(cl_program, const char *name, cl_int *err)
{
    _cl_kernel *k = new _cl_kernel;
    k->name = name ? name : "";
    memset(k->arg, 0, sizeof(k->arg));
    memset(k->arg_size, 0, sizeof(k->arg_size));
    if (err) *err = CL_SUCCESS;
    return k;
}

cl_int clSetKernelArg(cl_kernel k, cl_uint index, size_t size, const void *value)
{
    if (!k || index >= 8 || size > 8) return CL_INVALID_ARG_INDEX;
    memcpy(k->arg[index], value, size);
    k->arg_size[index] = size;
    return CL_SUCCESS;
}

cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem buf, cl_bool, size_t offset, size_t size, const void *ptr, cl_uint,
                            const cl_event *, cl_event *event)
{
    if (!buf || offset + size > buf->size) return CL_INVALID_VALUE;
    memcpy(buf->data + offset, ptr, size);
    if (event) *event = new_event();
    return CL_SUCCESS;
}

cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem buf, cl_bool, size_t offset, size_t size, void *ptr, cl_uint,
                           const cl_event *, cl_event *event)
{
    if (!buf || offset + size > buf->size) return CL_INVALID_VALUE;
    memcpy(ptr, buf->data + offset, size);
    if (event) *event = new_event();
    return CL_SUCCESS;
}

static cl_mem arg_mem(cl_kernel k, int i)
{
    cl_mem m;
    memcpy(&m, k->arg[i], sizeof(m));
    return m;
}
static cl_int arg_int(cl_kernel k, int i)
{
    cl_int v;
    memcpy(&v, k->arg[i], sizeof(v));
    return v;
}

cl_int clEnqueueTask(cl_command_queue, cl_kernel k, cl_uint, const cl_event *, cl_event *event)
{
    if (!k) return CL_INVALID_VALUE;
    if (k->name == "network_v1")
    {
        cl_mem inputs = arg_mem(k, 0), params = arg_mem(k, 1), bias = arg_mem(k, 2), outs = arg_mem(k, 3),
               npl = arg_mem(k, 4);
        const cl_int n_layers = arg_int(k, 5), n_ins = arg_int(k, 6);
        oracle_mlp_forward_one((const float *)inputs->data, (const float *)params->data, (const float *)bias->data,
                               (float *)outs->data, (const int *)npl->data, n_layers, n_ins, g_shim_activation);
    }
    else if (k->name == "image_process")
    {
        // Semantics of the image filter bitstream are unknown (out of scope):
        // pass the frame through so the ring-buffer plumbing can still be exercised.
        cl_mem in = arg_mem(k, 0), out = arg_mem(k, 1);
        memcpy(out->data, in->data, in->size < out->size ? in->size : out->size);
    }
    else
        return CL_INVALID_KERNEL_NAME;
    g_shim_tasks++;
    if (event) *event = new_event();
    return CL_SUCCESS;
}

cl_int clWaitForEvents(cl_uint, const cl_event *) { return CL_SUCCESS; }
cl_int clReleaseEvent(cl_event e)
{
    delete e;
    return CL_SUCCESS;
}
cl_int clReleaseKernel(cl_kernel k)
{
    delete k;
    return CL_SUCCESS;
}
cl_int clReleaseProgram(cl_program p)
{
    delete p;
    return CL_SUCCESS;
}
cl_int clReleaseCommandQueue(cl_command_queue q)
{
    delete q;
    return CL_SUCCESS;
}
cl_int clReleaseContext(cl_context c)
{
    delete c;
    return CL_SUCCESS;
}

} // extern "C"

namespace aocl_utils
{
void _checkError(int line, const char *file, cl_int error, const char *msg, ...)
{
    if (error == CL_SUCCESS) return;
    char text[512];
    va_list ap;
    va_start(ap, msg);
    vsnprintf(text, sizeof(text), msg, ap);
    va_end(ap);
    char full[768];
    snprintf(full, sizeof(full), "OpenCL shim error %d at %s:%d: %s", (int)error, file, line, text);
    throw std::runtime_error(full);
}

std::string getBoardBinaryFile(const char *prefix, cl_device_id) { return std::string(prefix) + ".aocx(shim)"; }

cl_program createProgramFromBinary(cl_context, const char *, const cl_device_id *, unsigned) { return new _cl_program{0}; }
}

// Minimal stand-in for the Khronos OpenCL header -- TEST INFRASTRUCTURE ONLY.
//
// The reference host code (/root/reference/src/netFPGA.cpp) includes "CL/cl.hpp" and calls 19
// OpenCL C entry points; no OpenCL header, ICD or PoCL exists in this image.
// This file declares exactly the subset that source uses, with the standard Khronos signatures,
// so that the UNMODIFIED reference file compiles.  oracle/shim/cl_shim.cpp implements them as an
// in-process fake device: buffers are malloc'd, the queue is in-order and runs inline, and
// clEnqueueTask("network_v1") executes the CPU oracle (oracle/oracle_mlp.c).
#ifndef NETCUDA_SHIM_CL_HPP
#define NETCUDA_SHIM_CL_HPP

#include <stddef.h>
#include <stdint.h>

extern "C" {

typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef cl_uint cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_device_type;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_command_queue_properties;
typedef intptr_t cl_context_properties;

typedef struct _cl_platform_id *cl_platform_id;
typedef struct _cl_device_id *cl_device_id;
typedef struct _cl_context *cl_context;
typedef struct _cl_command_queue *cl_command_queue;
typedef struct _cl_mem *cl_mem;
typedef struct _cl_program *cl_program;
typedef struct _cl_kernel *cl_kernel;
typedef struct _cl_event *cl_event;

#define CL_SUCCESS 0
#define CL_INVALID_VALUE -30
#define CL_INVALID_KERNEL_NAME -46
#define CL_INVALID_ARG_INDEX -49
#define CL_FALSE 0
#define CL_TRUE 1
#define CL_COMPLETE 0x0
#define CL_DEVICE_TYPE_ACCELERATOR (1 << 3)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY (1 << 2)

cl_int clGetPlatformIDs(cl_uint num_entries, cl_platform_id *platforms, cl_uint *num_platforms);
cl_int clGetDeviceIDs(cl_platform_id platform, cl_device_type type, cl_uint num_entries, cl_device_id *devices,
                      cl_uint *num_devices);
cl_context clCreateContext(const cl_context_properties *props, cl_uint num_devices, const cl_device_id *devices,
                           void (*notify)(const char *, const void *, size_t, void *), void *user_data,
                           cl_int *errcode_ret);
cl_command_queue clCreateCommandQueue(cl_context ctx, cl_device_id dev, cl_command_queue_properties props,
                                      cl_int *errcode_ret);
cl_int clBuildProgram(cl_program program, cl_uint num_devices, const cl_device_id *devices, const char *options,
                      void (*notify)(cl_program, void *), void *user_data);
cl_event clCreateUserEvent(cl_context ctx, cl_int *errcode_ret);
cl_int clSetUserEventStatus(cl_event ev, cl_int status);
cl_mem clCreateBuffer(cl_context ctx, cl_mem_flags flags, size_t size, void *host_ptr, cl_int *errcode_ret);
cl_kernel clCreateKernel(cl_program program, const char *name, cl_int *errcode_ret);
cl_int clSetKernelArg(cl_kernel kernel, cl_uint index, size_t size, const void *value);
cl_int clEnqueueWriteBuffer(cl_command_queue q, cl_mem buf, cl_bool blocking, size_t offset, size_t size,
                            const void *ptr, cl_uint n_wait, const cl_event *wait_list, cl_event *event);
cl_int clEnqueueReadBuffer(cl_command_queue q, cl_mem buf, cl_bool blocking, size_t offset, size_t size, void *ptr,
                           cl_uint n_wait, const cl_event *wait_list, cl_event *event);
cl_int clEnqueueTask(cl_command_queue q, cl_kernel kernel, cl_uint n_wait, const cl_event *wait_list,
                     cl_event *event);
cl_int clWaitForEvents(cl_uint n, const cl_event *list);
cl_int clReleaseEvent(cl_event ev);
cl_int clReleaseKernel(cl_kernel k);
cl_int clReleaseProgram(cl_program p);
cl_int clReleaseCommandQueue(cl_command_queue q);
cl_int clReleaseContext(cl_context c);

} // extern "C"
#endif

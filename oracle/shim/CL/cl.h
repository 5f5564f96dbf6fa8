/*
 * Minimal OpenCL 1.2 declarations -- exactly the subset the reference host
 * program (/root/reference/main.cpp, main_aux_functions.h) uses.  Written for
 * this repo because the image ships no OpenCL headers; constant values are the
 * ones fixed by the Khronos OpenCL 1.2 specification.  TEST INFRASTRUCTURE:
 * only used to build oracle/_ref (the unmodified reference) -- never by the
 * product.  The functions are defined in cl_shim.cpp, which forwards them to
 * the OpenCL runtime found with dlopen() at run time.
 */
#ifndef ORACLE_SHIM_CL_H
#define ORACLE_SHIM_CL_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int8_t cl_char;
typedef uint8_t cl_uchar;
typedef int16_t cl_short;
typedef uint16_t cl_ushort;
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef int64_t cl_long;
typedef uint64_t cl_ulong;
typedef float cl_float;
typedef double cl_double;

typedef struct _cl_platform_id *cl_platform_id;
typedef struct _cl_device_id *cl_device_id;
typedef struct _cl_context *cl_context;
typedef struct _cl_command_queue *cl_command_queue;
typedef struct _cl_mem *cl_mem;
typedef struct _cl_program *cl_program;
typedef struct _cl_kernel *cl_kernel;
typedef struct _cl_event *cl_event;

typedef cl_uint cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_device_type;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_command_queue_properties;
typedef cl_uint cl_platform_info;
typedef cl_uint cl_device_info;
typedef cl_uint cl_mem_info;
typedef cl_uint cl_program_build_info;
typedef cl_uint cl_profiling_info;
typedef intptr_t cl_context_properties;

#define CL_SUCCESS 0
#define CL_BUILD_PROGRAM_FAILURE -11
#define CL_INVALID_MEM_OBJECT -38
#define CL_FALSE 0
#define CL_TRUE 1
#define CL_PLATFORM_NAME 0x0902
#define CL_DEVICE_TYPE_CPU (1 << 1)
#define CL_DEVICE_TYPE_GPU (1 << 2)
#define CL_DEVICE_MAX_COMPUTE_UNITS 0x1002
#define CL_DEVICE_NAME 0x102B
#define CL_DEVICE_EXTENSIONS 0x1030
#define CL_QUEUE_PROFILING_ENABLE (1 << 1)
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_SIZE 0x1102
#define CL_PROGRAM_BUILD_LOG 0x1183
#define CL_PROFILING_COMMAND_START 0x1282
#define CL_PROFILING_COMMAND_END 0x1283

cl_int clGetPlatformIDs(cl_uint, cl_platform_id *, cl_uint *);
cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info, size_t, void *, size_t *);
cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint, cl_device_id *, cl_uint *);
cl_int clGetDeviceInfo(cl_device_id, cl_device_info, size_t, void *, size_t *);
cl_context clCreateContext(const cl_context_properties *, cl_uint, const cl_device_id *,
                           void (*)(const char *, const void *, size_t, void *), void *, cl_int *);
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int *);
cl_program clCreateProgramWithSource(cl_context, cl_uint, const char **, const size_t *, cl_int *);
cl_int clBuildProgram(cl_program, cl_uint, const cl_device_id *, const char *, void (*)(cl_program, void *), void *);
cl_int clGetProgramBuildInfo(cl_program, cl_device_id, cl_program_build_info, size_t, void *, size_t *);
cl_kernel clCreateKernel(cl_program, const char *, cl_int *);
cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t, void *, cl_int *);
cl_int clSetKernelArg(cl_kernel, cl_uint, size_t, const void *);
cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel, cl_uint, const size_t *, const size_t *, const size_t *,
                              cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, void *, cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, const void *, cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueCopyBuffer(cl_command_queue, cl_mem, cl_mem, size_t, size_t, size_t, cl_uint, const cl_event *, cl_event *);
cl_int clWaitForEvents(cl_uint, const cl_event *);
cl_int clFinish(cl_command_queue);
cl_int clFlush(cl_command_queue);
cl_int clGetEventProfilingInfo(cl_event, cl_profiling_info, size_t, void *, size_t *);
cl_int clGetMemObjectInfo(cl_mem, cl_mem_info, size_t, void *, size_t *);
cl_int clReleaseMemObject(cl_mem);
cl_int clReleaseCommandQueue(cl_command_queue);
cl_int clReleaseProgram(cl_program);
cl_int clReleaseKernel(cl_kernel);
cl_int clReleaseContext(cl_context);

#ifdef __cplusplus
}
#endif
#endif

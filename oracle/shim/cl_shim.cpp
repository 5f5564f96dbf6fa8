// Forwarders from the minimal CL/cl.h of this directory to a real OpenCL
// runtime located with dlopen() at run time.  TEST INFRASTRUCTURE (see CL/cl.h).
//
// Search order: $AME_OPENCL_LIB, libOpenCL.so.1 (ICD loader; needs an .icd
// file), libnvidia-opencl.so.1 (the NVIDIA driver's implementation, used
// directly when the loader reports no platform).
//
// One deliberate behaviour: the reference passes an UNINITIALISED cl_mem as
// kernel argument 10 on the first 2-CP launch (main.cpp:506, 837).  The 2-CP
// kernels never read that argument, but a strict runtime may reject or crash on
// the garbage handle.  clSetKernelArg below therefore substitutes NULL for any
// 8-byte argument whose value is not a buffer this process created.  The
// reference source itself is compiled unmodified.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <set>

#include "CL/cl.h"

static void *g_lib = nullptr;
static std::set<cl_mem> g_buffers;

static void *try_open(const char *name) {
    void *h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (!h) return nullptr;
    typedef cl_int (*fn_t)(cl_uint, cl_platform_id *, cl_uint *);
    fn_t f = (fn_t)dlsym(h, "clGetPlatformIDs");
    cl_uint n = 0;
    if (!f || f(0, nullptr, &n) != CL_SUCCESS || n == 0) {
        fprintf(stderr, "[cl_shim] %s: no usable OpenCL platform\n", name);
        return nullptr;
    }
    fprintf(stderr, "[cl_shim] using %s (%u platform(s))\n", name, n);
    return h;
}

static void *lib() {
    if (g_lib) return g_lib;
    const char *env = getenv("AME_OPENCL_LIB");
    if (env && *env) g_lib = try_open(env);
    if (!g_lib) g_lib = try_open("libOpenCL.so.1");
    if (!g_lib) g_lib = try_open("libnvidia-opencl.so.1");
    if (!g_lib) {
        fprintf(stderr, "[cl_shim] no OpenCL runtime found\n");
        exit(97);
    }
    return g_lib;
}

template <typename F>
static F sym(const char *name) {
    void *p = dlsym(lib(), name);
    if (!p) {
        fprintf(stderr, "[cl_shim] missing symbol %s\n", name);
        exit(98);
    }
    return (F)p;
}

#define FWD(ret, name, params, args)              \
    extern "C" ret name params {                  \
        typedef ret(*fn_t) params;                \
        static fn_t f = sym<fn_t>(#name);         \
        return f args;                            \
    }

FWD(cl_int, clGetPlatformIDs, (cl_uint a, cl_platform_id *b, cl_uint *c), (a, b, c))
FWD(cl_int, clGetPlatformInfo, (cl_platform_id a, cl_platform_info b, size_t c, void *d, size_t *e), (a, b, c, d, e))
FWD(cl_int, clGetDeviceIDs, (cl_platform_id a, cl_device_type b, cl_uint c, cl_device_id *d, cl_uint *e), (a, b, c, d, e))
FWD(cl_int, clGetDeviceInfo, (cl_device_id a, cl_device_info b, size_t c, void *d, size_t *e), (a, b, c, d, e))
FWD(cl_context, clCreateContext,
    (const cl_context_properties *a, cl_uint b, const cl_device_id *c, void (*d)(const char *, const void *, size_t, void *),
     void *e, cl_int *f_),
    (a, b, c, d, e, f_))
FWD(cl_command_queue, clCreateCommandQueue, (cl_context a, cl_device_id b, cl_command_queue_properties c, cl_int *d), (a, b, c, d))
FWD(cl_program, clCreateProgramWithSource, (cl_context a, cl_uint b, const char **c, const size_t *d, cl_int *e), (a, b, c, d, e))
FWD(cl_int, clBuildProgram,
    (cl_program a, cl_uint b, const cl_device_id *c, const char *d, void (*e)(cl_program, void *), void *f_), (a, b, c, d, e, f_))
FWD(cl_int, clGetProgramBuildInfo, (cl_program a, cl_device_id b, cl_program_build_info c, size_t d, void *e, size_t *f_),
    (a, b, c, d, e, f_))
FWD(cl_kernel, clCreateKernel, (cl_program a, const char *b, cl_int *c), (a, b, c))
FWD(cl_int, clEnqueueNDRangeKernel,
    (cl_command_queue a, cl_kernel b, cl_uint c, const size_t *d, const size_t *e, const size_t *f_, cl_uint g, const cl_event *h,
     cl_event *i),
    (a, b, c, d, e, f_, g, h, i))
FWD(cl_int, clEnqueueReadBuffer,
    (cl_command_queue a, cl_mem b, cl_bool c, size_t d, size_t e, void *f_, cl_uint g, const cl_event *h, cl_event *i),
    (a, b, c, d, e, f_, g, h, i))
FWD(cl_int, clEnqueueWriteBuffer,
    (cl_command_queue a, cl_mem b, cl_bool c, size_t d, size_t e, const void *f_, cl_uint g, const cl_event *h, cl_event *i),
    (a, b, c, d, e, f_, g, h, i))
FWD(cl_int, clEnqueueCopyBuffer,
    (cl_command_queue a, cl_mem b, cl_mem c, size_t d, size_t e, size_t f_, cl_uint g, const cl_event *h, cl_event *i),
    (a, b, c, d, e, f_, g, h, i))
FWD(cl_int, clWaitForEvents, (cl_uint a, const cl_event *b), (a, b))
FWD(cl_int, clFinish, (cl_command_queue a), (a))
FWD(cl_int, clFlush, (cl_command_queue a), (a))
FWD(cl_int, clGetEventProfilingInfo, (cl_event a, cl_profiling_info b, size_t c, void *d, size_t *e), (a, b, c, d, e))
FWD(cl_int, clGetMemObjectInfo, (cl_mem a, cl_mem_info b, size_t c, void *d, size_t *e), (a, b, c, d, e))
FWD(cl_int, clReleaseCommandQueue, (cl_command_queue a), (a))
FWD(cl_int, clReleaseProgram, (cl_program a), (a))
FWD(cl_int, clReleaseKernel, (cl_kernel a), (a))
FWD(cl_int, clReleaseContext, (cl_context a), (a))

extern "C" cl_mem clCreateBuffer(cl_context a, cl_mem_flags b, size_t c, void *d, cl_int *e) {
    typedef cl_mem (*fn_t)(cl_context, cl_mem_flags, size_t, void *, cl_int *);
    static fn_t f = sym<fn_t>("clCreateBuffer");
    cl_mem m = f(a, b, c, d, e);
    if (m) g_buffers.insert(m);
    return m;
}

extern "C" cl_int clReleaseMemObject(cl_mem a) {
    typedef cl_int (*fn_t)(cl_mem);
    static fn_t f = sym<fn_t>("clReleaseMemObject");
    g_buffers.erase(a);
    return f(a);
}

extern "C" cl_int clSetKernelArg(cl_kernel k, cl_uint idx, size_t size, const void *value) {
    typedef cl_int (*fn_t)(cl_kernel, cl_uint, size_t, const void *);
    static fn_t f = sym<fn_t>("clSetKernelArg");
    if (size == sizeof(cl_mem) && value) {
        cl_mem m;
        memcpy(&m, value, sizeof m);
        if (m && !g_buffers.count(m)) {
            static int warned = 0;
            if (!warned++) fprintf(stderr, "[cl_shim] kernel arg %u is not a live buffer (uninitialised handle); passing NULL\n", idx);
            cl_mem null_mem = nullptr;
            return f(k, idx, size, &null_mem);
        }
    }
    return f(k, idx, size, value);
}

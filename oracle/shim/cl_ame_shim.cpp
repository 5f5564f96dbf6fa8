/*
 * OpenCL stand-in over libaffine_me.so -- TEST INFRASTRUCTURE (oracle/_ref only).
 *
 * Purpose: prove the drop-in boundary of include/affine_me.h against the UNMODIFIED reference host program.
 * /root/reference/main.cpp is compiled as it is, from where it lies, and linked against this file instead of an
 * OpenCL runtime: the 26 OpenCL entry points it uses are implemented here, and the one place where the reference hands
 * work to its device code -- 14 x clSetKernelArg + clEnqueueNDRangeKernel, main.cpp:827-866 / 914-953 -- is bound
 * to the library's C ABI exactly as INTEGRATION.md describes:
 *
 *     arg 0 / 1  (reference plane, current plane)   -> ame_upload_plane_ex(slot 1, REFERENCE) / (slot 0, CURRENT)
 *     arg 2 / 3 / 4 / 13 (W, H, lambda, extra iter) -> ame_create (first launch) / ame_search
 *     arg 8 / 9  (return costs / CPMVs buffers)     -> the ame_result arrays of the launched prediction type
 *     clWaitForEvents / clFinish                    -> ame_sync
 *     clGetEventProfilingInfo                       -> ame_exec_ns of that prediction type
 *
 * Everything else the reference does with OpenCL is host-visible buffer traffic (clCreateBuffer, Write / Copy / Read:
 * the reference-list rotation of main.cpp:597-699 included) and is served from plain host memory.  The four launches
 * of one (current, reference, lambda) triple -- FULL_2CP, FULL_3CP, HALF_2CP, HALF_3CP -- are ONE ame_search: the
 * first launch runs it, the other three find the result of their type ready (keyed by the contents' write counters).
 * The OpenCL C sources the reference reads and "builds" are ignored.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <set>
#include <string>
#include <vector>

#include "CL/cl.h"
#include "affine_me.h"

namespace {

struct Mem {
    std::vector<unsigned char> data;
    unsigned long long version = 0;  // bumped by every write into the buffer
};
struct Program { int nCP = 0; };
struct Kernel {
    int nCP = 0, ha = 0;
    unsigned char args[14][8];
};
struct Event { double ns = 0; };

std::set<Mem *> g_live;  // (the reference releases, and passes as kernel argument 10, handles it never created: main.cpp:506, 837)
unsigned long long g_version = 1;
ame_ctx *g_ctx = nullptr;
int g_W = 0, g_H = 0;
void *g_block = nullptr;  // pinned result block of the last search
ame_result g_res;
double g_ns[4] = {0, 0, 0, 0};
struct Key { const Mem *ref, *cur; unsigned long long vref, vcur; float lambda; int extra; } g_key = {nullptr, nullptr, 0, 0, 0.f, -1};
Event g_event;

Mem *M(cl_mem m) {
    Mem *p = reinterpret_cast<Mem *>(m);
    return g_live.count(p) ? p : nullptr;
}

void die(const char *what) {
    fprintf(stderr, "cl_ame_shim: %s: %s\n", what, ame_last_error());
    exit(3);
}

template <class T>
T arg_as(const Kernel *k, int i) {
    T v;
    memcpy(&v, k->args[i], sizeof v);
    return v;
}

cl_int info_string(const char *s, size_t size, void *out, size_t *ret) {
    const size_t n = strlen(s) + 1;
    if (ret) *ret = n;
    if (out && size) {
        const size_t c = n < size ? n : size;
        memcpy(out, s, c);
        static_cast<char *>(out)[c - 1] = 0;
    }
    return CL_SUCCESS;
}

}  // namespace

extern "C" {

cl_int clGetPlatformIDs(cl_uint n, cl_platform_id *ids, cl_uint *num) {
    if (num) *num = 1;
    if (ids && n) ids[0] = reinterpret_cast<cl_platform_id>(0x1);
    return CL_SUCCESS;
}
cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info, size_t size, void *out, size_t *ret) {
    return info_string("affine_me stand-in (CUDA, sm_100a)", size, out, ret);
}
cl_int clGetDeviceIDs(cl_platform_id, cl_device_type type, cl_uint n, cl_device_id *ids, cl_uint *num) {
    const cl_uint have = type == CL_DEVICE_TYPE_GPU ? 1 : 0;  // one GPU, no CPU device (there is no CPU fallback)
    if (num) *num = have;
    if (ids && n && have) ids[0] = reinterpret_cast<cl_device_id>(0x2);
    return CL_SUCCESS;
}
cl_int clGetDeviceInfo(cl_device_id, cl_device_info what, size_t size, void *out, size_t *ret) {
    if (what == CL_DEVICE_MAX_COMPUTE_UNITS) {
        const cl_uint v = 148;
        if (ret) *ret = sizeof v;
        if (out && size >= sizeof v) memcpy(out, &v, sizeof v);
        return CL_SUCCESS;
    }
    return info_string(what == CL_DEVICE_NAME ? "NVIDIA B200 through libaffine_me.so" : "", size, out, ret);
}
cl_context clCreateContext(const cl_context_properties *, cl_uint, const cl_device_id *, void (*)(const char *, const void *, size_t, void *), void *, cl_int *e) {
    if (e) *e = CL_SUCCESS;
    return reinterpret_cast<cl_context>(0x3);
}
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int *e) {
    if (e) *e = CL_SUCCESS;
    return reinterpret_cast<cl_command_queue>(0x4);
}
cl_program clCreateProgramWithSource(cl_context, cl_uint, const char **, const size_t *, cl_int *e) {
    if (e) *e = CL_SUCCESS;
    return reinterpret_cast<cl_program>(new Program());
}
cl_int clBuildProgram(cl_program p, cl_uint, const cl_device_id *, const char *options, void (*)(cl_program, void *), void *) {
    const char *d = options ? strstr(options, "-DnCP=") : nullptr;  // main.cpp:389-390
    reinterpret_cast<Program *>(p)->nCP = d ? atoi(d + 6) : 0;
    return reinterpret_cast<Program *>(p)->nCP == 2 || reinterpret_cast<Program *>(p)->nCP == 3 ? CL_SUCCESS : CL_BUILD_PROGRAM_FAILURE;
}
cl_int clGetProgramBuildInfo(cl_program, cl_device_id, cl_program_build_info, size_t size, void *out, size_t *ret) { return info_string("", size, out, ret); }
cl_kernel clCreateKernel(cl_program p, const char *name, cl_int *e) {
    Kernel *k = new Kernel();
    k->nCP = reinterpret_cast<Program *>(p)->nCP;
    k->ha = strcmp(name, "affine_gradient_mult_sizes_HA") == 0;  // main.cpp:440-446
    memset(k->args, 0, sizeof k->args);
    if (e) *e = (k->ha || strcmp(name, "affine_gradient_mult_sizes") == 0) ? CL_SUCCESS : -46 /* CL_INVALID_KERNEL_NAME */;
    return reinterpret_cast<cl_kernel>(k);
}
cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t size, void *host, cl_int *e) {
    Mem *m = new Mem();
    m->data.resize(size);
    if (host) memcpy(m->data.data(), host, size);
    m->version = g_version++;
    g_live.insert(m);
    if (e) *e = CL_SUCCESS;
    return reinterpret_cast<cl_mem>(m);
}
cl_int clSetKernelArg(cl_kernel k, cl_uint idx, size_t size, const void *value) {
    if (idx >= 14 || size > 8) return -50 /* CL_INVALID_ARG_VALUE */;
    memset(reinterpret_cast<Kernel *>(k)->args[idx], 0, 8);
    if (value) memcpy(reinterpret_cast<Kernel *>(k)->args[idx], value, size);
    return CL_SUCCESS;
}

// affine.cl:11 / :960 -- (ref, cur, W, H, lambda, hGrad, vGrad, eqScratch, bestCost, bestCpmvs, prevCpmvs, debug, retCU, extraIter)
cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel kk, cl_uint, const size_t *, const size_t *, const size_t *, cl_uint, const cl_event *, cl_event *ev) {
    const Kernel *k = reinterpret_cast<const Kernel *>(kk);
    const Mem *ref = M(arg_as<cl_mem>(k, 0)), *cur = M(arg_as<cl_mem>(k, 1));
    const int W = arg_as<cl_int>(k, 2), H = arg_as<cl_int>(k, 3), extra = arg_as<cl_int>(k, 13);
    const float lambda = arg_as<cl_float>(k, 4);
    Mem *costs = M(arg_as<cl_mem>(k, 8)), *cpmvs = M(arg_as<cl_mem>(k, 9));
    if (!ref || !cur || !costs || !cpmvs) return CL_INVALID_MEM_OBJECT;
    if (!g_ctx) {
        if (ame_create(&g_ctx, 0, W, H, 2, 1) != AME_OK) die("ame_create");
        g_W = W;
        g_H = H;
        g_block = ame_alloc_host(ame_result_block_bytes(g_ctx));
        if (!g_block || ame_result_bind(g_ctx, g_block, &g_res) != AME_OK) die("ame_alloc_host");
    }
    if (W != g_W || H != g_H) { fprintf(stderr, "cl_ame_shim: frame size changed\n"); exit(3); }
    const Key key = {ref, cur, ref->version, cur->version, lambda, extra};
    if (memcmp(&key, &g_key, sizeof key) != 0) {  // first launch of this (current, reference, lambda) triple: one search for all four types
        if (ame_upload_plane_ex(g_ctx, 0, reinterpret_cast<const uint16_t *>(cur->data.data()), AME_ROLE_CURRENT) != AME_OK) die("ame_upload_plane_ex(current)");
        if (ame_upload_plane_ex(g_ctx, 1, reinterpret_cast<const uint16_t *>(ref->data.data()), AME_ROLE_REFERENCE) != AME_OK) die("ame_upload_plane_ex(reference)");
        if (ame_search(g_ctx, 0, 1, lambda, extra, &g_res) != AME_OK) die("ame_search");
        if (ame_sync(g_ctx) != AME_OK) die("ame_sync");
        if (ame_exec_ns(g_ctx, g_ns, 1) != AME_OK) die("ame_exec_ns");
        g_key = key;
    }
    const int pred = k->ha * 2 + (k->nCP - 2);  // AME_FULL_2CP, AME_FULL_3CP, AME_HALF_2CP, AME_HALF_3CP
    const size_t n = (size_t)ame_result_len(g_ctx, pred);
    if (costs->data.size() < n * sizeof(int64_t) || cpmvs->data.size() < n * sizeof(ame_cpmvs)) { fprintf(stderr, "cl_ame_shim: result buffers too small\n"); exit(3); }
    memcpy(costs->data.data(), g_res.cost[pred], n * sizeof(int64_t));
    memcpy(cpmvs->data.data(), g_res.cpmvs[pred], n * sizeof(ame_cpmvs));
    costs->version = g_version++;
    cpmvs->version = g_version++;
    g_event.ns = g_ns[pred];
    if (ev) *ev = reinterpret_cast<cl_event>(&g_event);
    return CL_SUCCESS;
}

cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem m, cl_bool, size_t off, size_t size, void *dst, cl_uint, const cl_event *, cl_event *) {
    if (!M(m) || off + size > M(m)->data.size()) return CL_INVALID_MEM_OBJECT;
    memcpy(dst, M(m)->data.data() + off, size);
    return CL_SUCCESS;
}
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem m, cl_bool, size_t off, size_t size, const void *src, cl_uint, const cl_event *, cl_event *) {
    if (!M(m) || off + size > M(m)->data.size()) return CL_INVALID_MEM_OBJECT;
    memcpy(M(m)->data.data() + off, src, size);
    M(m)->version = g_version++;
    return CL_SUCCESS;
}
cl_int clEnqueueCopyBuffer(cl_command_queue, cl_mem s, cl_mem d, size_t so, size_t dof, size_t size, cl_uint, const cl_event *, cl_event *) {
    if (!M(s) || !M(d) || so + size > M(s)->data.size() || dof + size > M(d)->data.size()) return CL_INVALID_MEM_OBJECT;
    memmove(M(d)->data.data() + dof, M(s)->data.data() + so, size);
    M(d)->version = g_version++;
    return CL_SUCCESS;
}
cl_int clWaitForEvents(cl_uint, const cl_event *) { return CL_SUCCESS; }  // (the search was synchronised at launch)
cl_int clFinish(cl_command_queue) { return CL_SUCCESS; }
cl_int clFlush(cl_command_queue) { return CL_SUCCESS; }
cl_int clGetEventProfilingInfo(cl_event e, cl_profiling_info what, size_t size, void *out, size_t *ret) {
    const cl_ulong v = what == CL_PROFILING_COMMAND_END ? (cl_ulong)(reinterpret_cast<Event *>(e)->ns + 0.5) : 0;
    if (ret) *ret = sizeof v;
    if (out && size >= sizeof v) memcpy(out, &v, sizeof v);
    return CL_SUCCESS;
}
cl_int clGetMemObjectInfo(cl_mem m, cl_mem_info, size_t size, void *out, size_t *ret) {
    const size_t v = M(m) ? M(m)->data.size() : 0;
    if (ret) *ret = sizeof v;
    if (out && size >= sizeof v) memcpy(out, &v, sizeof v);
    return CL_SUCCESS;
}
cl_int clReleaseMemObject(cl_mem m) {
    Mem *p = M(m);
    if (!p) return CL_INVALID_MEM_OBJECT;
    g_live.erase(p);
    delete p;
    return CL_SUCCESS;
}
cl_int clReleaseCommandQueue(cl_command_queue) { return CL_SUCCESS; }
cl_int clReleaseProgram(cl_program p) {
    delete reinterpret_cast<Program *>(p);
    return CL_SUCCESS;
}
cl_int clReleaseKernel(cl_kernel k) {
    delete reinterpret_cast<Kernel *>(k);
    return CL_SUCCESS;
}
cl_int clReleaseContext(cl_context) {
    if (g_ctx) {
        ame_free_host(g_block);
        ame_destroy(g_ctx);
        g_ctx = nullptr;
    }
    return CL_SUCCESS;
}

}  // extern "C"

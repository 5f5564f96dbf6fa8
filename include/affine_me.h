/*
 * affine_me.h -- C ABI of the B200-native affine motion-estimation search.
 *
 * This is the drop-in boundary for the one hot path of iagostorch/VVC-Affine-GPU:
 * everything the reference host does between "frames are in host memory" and
 * "per-CU costs / CPMVs are in host memory", i.e. the OpenCL buffer + kernel
 * dispatch block of /root/reference/main.cpp:484-552 and :746-966 and the
 * readback of main_aux_functions.h:335-383.  Plain C: pointers and sizes only.
 *
 * Reference interface each entry point replaces (file:line under /root/reference):
 *
 *   ame_create          clCreateContext / clCreateCommandQueue x5 / clBuildProgram x2 /
 *                       clCreateKernel x4 / clCreateBuffer (frames, results)
 *                                                 main.cpp:223-242, 337-359, 373-447, 484-527
 *   ame_upload_plane    clEnqueueWriteBuffer of a current / reference frame
 *                                                 main.cpp:572-576, 603, 658, 711-715
 *                       (the reference-list rotation by clEnqueueCopyBuffer, :597-699,
 *                        becomes a choice of slot index on the host -- no device copies)
 *   ame_search          14x clSetKernelArg + clEnqueueNDRangeKernel for FULL_2CP, FULL_3CP,
 *                       HALF_2CP, HALF_3CP              main.cpp:827-866, 914-953
 *                       + the two blocking clEnqueueReadBuffer per kernel
 *                                                 main_aux_functions.h:335-383
 *   ame_flush/ame_sync  clWaitForEvents / clFinish      main.cpp:856-860, 943-947, 973
 *   ame_destroy         clRelease*                      main.cpp:1019-1117
 *   ame_cpmvs           Cpmvs                           typedef.h:1-8
 *   result indexing     returnArrayIdx                  affine.cl:936, 1929
 *
 * Semantics that match the reference: one search = one (current frame,
 * reference frame, lambda) triple through all four prediction types; results
 * are four cost arrays (int64) and four CPMV arrays (28-byte structs) indexed
 * ctu*201 + RETURN_STRIDE_LIST[size] + cu (aligned) and
 * ctu*284 + HA_RETURN_STRIDE_LIST[group] + cu (half-aligned).
 *
 * Threading: one ame_ctx per GPU, driven by one host thread at a time;
 * contexts share nothing.  All functions return 0 on success or a negative
 * AME_E_* code; ame_last_error() returns the message of the calling thread's
 * last failure.  There is no CPU fallback: without a CUDA device ame_create fails.
 * A CUDA error after work has been issued (ame_flush / ame_sync) makes the context unusable: the searches in flight
 * are dropped, every later call returns AME_E_CUDA with the first error, ame_destroy frees it.
 */
#ifndef AFFINE_ME_H
#define AFFINE_ME_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AME_API_VERSION 1

enum {
    AME_OK = 0,
    AME_E_INVALID = -1,  /* bad argument */
    AME_E_CUDA = -2,     /* CUDA runtime error (see ame_last_error) */
    AME_E_NOMEM = -3,
    AME_E_STATE = -4     /* e.g. too many searches in flight */
};

/* Prediction types, in the reference's order (constants.h:15-21). */
enum { AME_FULL_2CP = 0, AME_FULL_3CP = 1, AME_HALF_2CP = 2, AME_HALF_3CP = 3, AME_N_PREDS = 4 };

#define AME_ALIGNED_CUS_PER_CTU 201 /* constants.cl:118 */
#define AME_HALF_CUS_PER_CTU 284    /* constants.cl:119 */

/* == Cpmvs (typedef.h:5-8): control-point MVs in 1/16-pel units.  nCPs is never
 * written by the reference kernels; this implementation stores 0 there. */
typedef struct {
    int32_t nCPs;
    int32_t ltx, lty, rtx, rty, lbx, lby;
} ame_cpmvs;

/* Destination of one search: separate arrays, like the reference's buffers.
 * cost[p] / cpmvs[p] must hold ame_result_len(ctx, p) elements. */
typedef struct {
    int64_t *cost[AME_N_PREDS];
    ame_cpmvs *cpmvs[AME_N_PREDS];
} ame_result;

/* Options for ame_set_option. */
enum {
    AME_OPT_CVT_RULE = 1,      /* double->int rule of the delta rounding for out-of-range values:
                                  1 = NVIDIA cvt.rzi (default; what the reference does on an NVIDIA GPU),
                                  0 = x86 cvttsd2si (what it does on a CPU OpenCL device) */
    AME_OPT_FUSED_BACKSUB = 2, /* 1 (default) = back-substitution accumulates with FMA, as OpenCL's
                                  default FP_CONTRACT ON compiles affine.cl:851 */
    AME_OPT_EARLY_EXIT = 3,    /* 1 (default) = stop a CU's refinement once its CPMVs revisit an
                                  already evaluated state (results are identical either way) */
    AME_OPT_REUSE_START = 4,   /* 1 (default) = a 3-CP search whose start state moves every sub-block exactly like the best
                                  2-CP state reuses that state's SATD and normal equations instead of evaluating it again
                                  (results are identical either way) */
    AME_OPT_SHARE_FIRST = 5,   /* 1 (default) = the first evaluation of the 2-CP searches (zero motion for every CU) is
                                  computed once per 4x4 block and summed per CU instead of once per CU
                                  (results are identical either way) */
    AME_OPT_BIG_TMA = 6,       /* 1 = CUs of 256..1024 sub-blocks stage the raw search window under their MV field in shared
                                  memory with TMA (cp.async.bulk.tensor.2d) and run both interpolation stages from there;
                                  0 = they read the pre-filtered phase planes like the small CUs
                                  (results are identical either way; the default is the faster one) */
    AME_OPT_GROUP_BY_REF = 7   /* 1 (default) = the searches of a launch sequence that share a reference plane (the long-term
                                  references of main.cpp:591-707 are searched by many frames) are worked on CTU by CTU side
                                  by side, so that they share the plane's rows in L2; 0 = search after search */
};

typedef struct ame_ctx ame_ctx;

/* nCtus = ceil(W/128) * ceil(H/128)  (constants.h:73-79 holds the same numbers as a whitelist). */
int ame_num_ctus(int width, int height);

/* device: CUDA ordinal.  num_slots: frame planes resident on the GPU (>= 2).
 * max_in_flight: searches that may be queued before ame_sync (>= 1).
 * width must be a multiple of 8 and both dimensions >= 16. */
int ame_create(ame_ctx **out, int device, int width, int height, int num_slots, int max_in_flight);
void ame_destroy(ame_ctx *ctx);

int ame_result_len(const ame_ctx *ctx, int pred); /* nCtus*201 or nCtus*284 */
int ame_set_option(ame_ctx *ctx, int option, int value);

/* Asynchronous: copies a W x H plane of 10-bit samples (row-major uint16, like the
 * reference's `unsigned short` frames) into slot `slot` and prepares it for motion compensation
 * (edge replication + first interpolation stage).  `plane` must stay valid
 * until the next ame_sync; pinned memory makes the copy truly asynchronous. */
int ame_upload_plane(ame_ctx *ctx, int slot, const uint16_t *plane);

/* Same with an explicit role mask.  AME_ROLE_CURRENT: the plane will be searched FOR (original frames; kept a
 * second time in 4x4-block order); AME_ROLE_REFERENCE: the plane will be searched IN (reconstructed frames) -- this
 * runs the horizontal interpolation stage for all 16 phases once and keeps the result
 * (2 x 16 x (W+320) x (H+320) x 2 bytes of device memory, allocated on the slot's first reference upload).
 * ame_upload_plane gives both roles; ame_search fails with AME_E_STATE on a slot that lacks the role it needs. */
enum { AME_ROLE_CURRENT = 1, AME_ROLE_REFERENCE = 2 };
int ame_upload_plane_ex(ame_ctx *ctx, int slot, const uint16_t *plane, int roles);

/* Runs the preparation of ame_upload_plane_ex again on the plane that is already resident in `slot` (no host-to-device
 * copy): the block-ordered copy (AME_ROLE_CURRENT) and / or edge replication + the horizontal interpolation stage for
 * all 16 phases (AME_ROLE_REFERENCE; that stage is half of aux_functions.cl:1096-1239, run once per plane here instead
 * of once per sub-block and iteration).  bench.py uses it to time that work with the inputs already in HBM. */
int ame_prepare_plane(ame_ctx *ctx, int slot, int roles);

/* Queues one search of the plane in cur_slot against the plane in ref_slot.
 * `out` arrays are HOST memory and are valid after the next ame_sync.
 * extra_iters == --ExtraGradientIter. */
int ame_search(ame_ctx *ctx, int cur_slot, int ref_slot, float lambda, int extra_iters, const ame_result *out);

/* Same search, results left in DEVICE memory (no D2H): `out` arrays are device
 * pointers obtained from ame_device_result.  Used when the consumer is on the GPU. */
int ame_search_device(ame_ctx *ctx, int cur_slot, int ref_slot, float lambda, int extra_iters, int result_index);
int ame_device_result(ame_ctx *ctx, int result_index, ame_result *out); /* result_index < max_in_flight */

/* Result arrays in ONE block of host memory, laid out like the library's device-side result block:
 * ame_result_bind points the eight arrays of `out` into `block` (ame_result_block_bytes bytes, e.g. from
 * ame_alloc_host).  A search whose destination was bound this way comes back with one device-to-host copy
 * instead of eight (the reference reads its result buffers one by one, main_aux_functions.h:335-383). */
uint64_t ame_result_block_bytes(const ame_ctx *ctx);
int ame_result_bind(const ame_ctx *ctx, void *block, ame_result *out);

/* Launches everything queued so far without waiting. */
int ame_flush(ame_ctx *ctx);
/* ame_flush + wait for all queued uploads, searches and result copies. */
int ame_sync(ame_ctx *ctx);

/* Device time (ms, CUDA events on the context's stream) of the search kernels
 * launched by the most recent ame_flush/ame_sync, and their launch count. */
int ame_last_kernel_ms(ame_ctx *ctx, float *ms, int *launches);

/* Device time, in nanoseconds, of the searches run since the last reset, per prediction type in the reference's
 * order (kernelExecutionTime[FULL_2CP..HALF_3CP], main.cpp:862-866, 949-953, printed by reportTimingResults,
 * main_aux_functions.h:1416-1446).  The four types run fused here: the 2-CP searches of a launch sequence (aligned
 * and half-aligned CUs together) are timed as one interval on the device and so are the 3-CP searches; each
 * interval is split between FULL and HALF by the number of 4x4 sub-block evaluations either side accounted for.
 * Needs ame_sync first. */
int ame_exec_ns(ame_ctx *ctx, double ns[AME_N_PREDS], int reset);

/* Device-side stopwatch on the context's stream (CUDA events): ame_timer_start records the start event
 * behind everything issued so far, ame_timer_stop records the stop event, waits for it and returns the
 * elapsed milliseconds.  Used by bench.py; the events see uploads, kernels and result copies alike (the start
 * mark waits for whatever was issued before it on the upload and result streams, too). */
int ame_timer_start(ame_ctx *ctx);
int ame_timer_stop(ame_ctx *ctx, float *ms);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost). */
void *ame_alloc_host(uint64_t bytes);
void ame_free_host(void *p);

/* CU geometry of result element k (0..200 / 0..283) of prediction type pred:
 * out = {x, y, w, h} inside the CTU.  Returns the size-group index or -1. */
int ame_cu_geometry(int pred, int k, int out[4]);

const char *ame_last_error(void);
int ame_version(void);

#ifdef __cplusplus
}
#endif
#endif

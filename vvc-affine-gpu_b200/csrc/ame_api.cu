// C-ABI layer of include/affine_me.h: context, frame slots, queued searches, result copies.
// Replaces the OpenCL buffer / argument / enqueue / readback code of the reference
// (/root/reference/main.cpp:484-552, 746-966; main_aux_functions.h:335-383).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "ame_device.h"
#include "ame_geometry.h"

using namespace ame;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) return fail(AME_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

#ifndef AME_BIG_TMA_DEFAULT
#define AME_BIG_TMA_DEFAULT 1
#endif

struct Slot {
    uint16_t *raw = nullptr;    // W x H, as uploaded
    uint4 *blk = nullptr;       // current-frame role: the plane in 4x4-block order; allocated on first use
    uint4 *refT = nullptr;      // reference role: 2 x 16 pre-filtered planes; allocated on first use
    uint16_t *refPad = nullptr; // reference role: the edge-replicated plane (input of the phase filter; TMA source of the big-CU windows)
    bool hasRaw = false;                  // a plane has been uploaded
    bool hasCur = false, hasRef = false;  // blk / refT match the current contents of raw
};

struct ResultBlock {  // one per in-flight search, device memory (layout: ame_ctx::resOff / resBytes)
    char *base = nullptr;
    long long *cost[4];
    ame_cpmvs *cpmvs[4];
};

struct Pending {
    int resultIdx, curSlot, refSlot;
    bool toHost;
    ame_result host;
};

struct ame_ctx {
    int device = 0, W = 0, H = 0, nCtus = 0, ctuCols = 0, padStride = 0;
    int numSlots = 0, maxInFlight = 0;
    int cvtRule = 1, fusedBacksub = 1, earlyExit = 1, reuseStart = 1, shareFirst = 1, bigTma = AME_BIG_TMA_DEFAULT;
    unsigned char *dTmaps = nullptr;  // [numSlots][4] CUtensorMap over refPad (boxes of tma_box(i) x tma_box(j) samples)
    int queuedExtra = 0;
    int numSMs = 0;
    uint32_t *dSlotTab = nullptr;
    // scratch of a launch sequence (ame_device.h): search state, accumulators and work lists for maxInFlight passes
    CuState *dState = nullptr;
    CuAccum *dAccum = nullptr;
    size_t seqSlots = 0;
    WorkLists *dWork = nullptr;               // [kMaxSteps]
    uint4 *dSmallList = nullptr;              // [2][seqSlots]
    uint2 *dBigList = nullptr;                // [2][bigCap]
    unsigned *dGwOut = nullptr;               // [2 * seqSlots]
    unsigned *dRowTab = nullptr;              // [2][seqPasses * nCtus] (pass, CTU) of every row of the state array + the units of ame_iter0_kernel, per launch sequence
    int groupByRef = 1;                       // rows ordered (reference plane, CTU, pass) instead of (pass, CTU)
    size_t bigCap = 0;
    unsigned long long *dScan = nullptr;      // 4 x scan_words(seqSlots): ordered compaction of the update / phase kernels
    Telemetry *dTele = nullptr;
    int *dTab0 = nullptr;
    bool poisoned = false;                    // a CUDA call failed after work had been issued: only ame_destroy is left
    std::string poison;
    cudaStream_t stream = nullptr, side = nullptr;  // search kernels (big CUs / small CUs)
    cudaStream_t up = nullptr, down = nullptr;      // plane uploads + preparation / result copies
    cudaEvent_t evUp = nullptr, evKernels = nullptr, evAux = nullptr;
    bool kernelsRecorded = false;
    cudaEvent_t evStart = nullptr, evStop = nullptr, evFork = nullptr, evJoin = nullptr, evT0 = nullptr, evT1 = nullptr;
    bool timed = false;
    int lastLaunches = 0;
    size_t planeElems = 0;  // samples of the padded plane
    size_t planeSetRecs = 0;  // 16-byte records of the 2 x 16 tiled pre-filtered planes of one reference (refT)
    std::vector<Slot> slots;
    std::vector<ResultBlock> results;
    PassDesc *dPasses = nullptr;   // device [maxInFlight]
    PassDesc *hPasses = nullptr;   // pinned [maxInFlight]
    std::vector<Pending> queued;   // searches queued since the last flush
    std::vector<Pending> inflight; // launched, results not yet known complete
    size_t lens[4];
    size_t resOff[8] = {0}, resBytes = 0;  // byte offsets of cost[0..3], cpmvs[0..3] inside a result block
};

// A CUDA call that fails after work has been issued leaves streams, scratch and result blocks in an unknown state:
// the context keeps the error and every later call returns it (ame_destroy is the only way out).
static int poison(ame_ctx *c, const char *what, cudaError_t e) {
    c->poisoned = true;
    c->poison = std::string(what) + ": " + cudaGetErrorString(e);
    c->queued.clear();
    c->inflight.clear();
    return fail(AME_E_CUDA, "%s", c->poison.c_str());
}
#define CU_POISON(expr)                                      \
    do {                                                     \
        cudaError_t e_ = (expr);                             \
        if (e_ != cudaSuccess) return poison(c, #expr, e_);  \
    } while (0)
#define CHECK_USABLE(c, name)                                                                                       \
    do {                                                                                                            \
        if (!(c)) return fail(AME_E_INVALID, name ": ctx is NULL");                                                 \
        if ((c)->poisoned) return fail(AME_E_CUDA, name ": context unusable after an earlier CUDA error (%s)", (c)->poison.c_str()); \
    } while (0)

extern "C" {

int ame_version(void) { return AME_API_VERSION; }
const char *ame_last_error(void) { return g_err.c_str(); }

int ame_num_ctus(int width, int height) { return ((width + 127) / 128) * ((height + 127) / 128); }

int ame_cu_geometry(int pred, int k, int out[4]) {
    if (pred < 0 || pred >= AME_N_PREDS || !out) return -1;
    static const std::vector<CuDesc> tabs[2] = {ctu_cus(0), ctu_cus(1)};
    const std::vector<CuDesc> &t = tabs[pred >= 2];
    if (k < 0 || k >= (int)t.size()) return -1;
    out[0] = t[k].x; out[1] = t[k].y; out[2] = t[k].w; out[3] = t[k].h;
    return t[k].group;
}

void *ame_alloc_host(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void ame_free_host(void *p) {
    if (p) cudaFreeHost(p);
}

void ame_destroy(ame_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (cudaStream_t s : {c->up, c->stream, c->side, c->down}) if (s) cudaStreamSynchronize(s);
    for (Slot &s : c->slots) { cudaFree(s.raw); cudaFree(s.blk); cudaFree(s.refT); cudaFree(s.refPad); }
    cudaFree(c->dTmaps);
    for (ResultBlock &r : c->results) cudaFree(r.base);
    cudaFree(c->dSlotTab);
    cudaFree(c->dPasses);
    cudaFree(c->dState);
    cudaFree(c->dAccum);
    cudaFree(c->dGwOut);
    cudaFree(c->dRowTab);
    cudaFree(c->dTab0);
    cudaFree(c->dWork);
    cudaFree(c->dScan);
    cudaFree(c->dTele);
    cudaFree(c->dSmallList);
    cudaFree(c->dBigList);
    if (c->hPasses) cudaFreeHost(c->hPasses);
    if (c->evStart) cudaEventDestroy(c->evStart);
    if (c->evStop) cudaEventDestroy(c->evStop);
    if (c->evT0) cudaEventDestroy(c->evT0);
    if (c->evT1) cudaEventDestroy(c->evT1);
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->evJoin) cudaEventDestroy(c->evJoin);
    for (cudaEvent_t e : {c->evUp, c->evKernels, c->evAux}) if (e) cudaEventDestroy(e);
    if (c->up) cudaStreamDestroy(c->up);
    if (c->down) cudaStreamDestroy(c->down);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int ame_create(ame_ctx **out, int device, int width, int height, int num_slots, int max_in_flight) {
    if (!out) return fail(AME_E_INVALID, "ame_create: out is NULL");
    *out = nullptr;
    if (width < 16 || height < 16 || (width % 8) != 0) return fail(AME_E_INVALID, "ame_create: unsupported size %dx%d (width must be a multiple of 8, both >= 16)", width, height);
    if (num_slots < 2 || max_in_flight < 1) return fail(AME_E_INVALID, "ame_create: need num_slots >= 2 and max_in_flight >= 1");
    if (tiled_plane_set_recs(width + 2 * kPad, height + 2 * kPad) > 0xffffffffull || (long long)ame_num_ctus(width, height) > 0xffff)
        return fail(AME_E_INVALID, "ame_create: %dx%d is beyond the 32-bit record index of the pre-filtered planes / 16-bit CTU index of the work lists", width, height);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(AME_E_CUDA, "ame_create: no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(AME_E_INVALID, "ame_create: device %d out of range (%d devices)", device, ndev);
    CU_TRY(cudaSetDevice(device));
    ame_ctx *c = new ame_ctx();
    cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, device);
    c->device = device;
    c->W = width;
    c->H = height;
    c->ctuCols = (width + 127) / 128;
    c->nCtus = ame_num_ctus(width, height);
    c->padStride = width + 2 * kPad;
    c->numSlots = num_slots;
    c->maxInFlight = max_in_flight;
    for (int p = 0; p < 4; p++) c->lens[p] = (size_t)c->nCtus * (p < 2 ? AME_ALIGNED_CUS_PER_CTU : AME_HALF_CUS_PER_CTU);
#define CTX_TRY(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            int rc_ = fail(e_ == cudaErrorMemoryAllocation ? AME_E_NOMEM : AME_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
            ame_destroy(c);                                                                    \
            return rc_;                                                                        \
        }                                                                                      \
    } while (0)
    CTX_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CTX_TRY(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    CTX_TRY(cudaStreamCreateWithFlags(&c->up, cudaStreamNonBlocking));
    CTX_TRY(cudaStreamCreateWithFlags(&c->down, cudaStreamNonBlocking));
    CTX_TRY(cudaEventCreateWithFlags(&c->evUp, cudaEventDisableTiming));
    CTX_TRY(cudaEventCreateWithFlags(&c->evKernels, cudaEventDisableTiming));
    CTX_TRY(cudaEventCreateWithFlags(&c->evAux, cudaEventDisableTiming));
    CTX_TRY(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
    CTX_TRY(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
    CTX_TRY(cudaEventCreate(&c->evT0));
    CTX_TRY(cudaEventCreate(&c->evT1));
    CTX_TRY(cudaEventCreate(&c->evStart));
    CTX_TRY(cudaEventCreate(&c->evStop));
    c->slots.resize(num_slots);
    const size_t rawBytes = (size_t)width * height * sizeof(uint16_t);
    c->planeElems = (size_t)c->padStride * (height + 2 * kPad);
    c->planeSetRecs = tiled_plane_set_recs(c->padStride, height + 2 * kPad);
    CTX_TRY(cudaMalloc(&c->dTmaps, (size_t)num_slots * 4 * sizeof(CUtensorMap)));
    for (Slot &s : c->slots) CTX_TRY(cudaMalloc(&s.raw, rawBytes));
    size_t total = 0;
    for (int p = 0; p < 4; p++) { c->resOff[p] = total; total += (c->lens[p] * sizeof(long long) + 255) & ~(size_t)255; }
    for (int p = 0; p < 4; p++) { c->resOff[4 + p] = total; total += (c->lens[p] * sizeof(ame_cpmvs) + 255) & ~(size_t)255; }
    c->resBytes = total;
    c->results.resize(max_in_flight);
    for (ResultBlock &r : c->results) {
        CTX_TRY(cudaMalloc(&r.base, total));
        for (int p = 0; p < 4; p++) {
            r.cost[p] = reinterpret_cast<long long *>(r.base + c->resOff[p]);
            r.cpmvs[p] = reinterpret_cast<ame_cpmvs *>(r.base + c->resOff[4 + p]);
        }
    }
    {   // scratch of one launch sequence: at most min(max_in_flight, kMaxPasses) passes
        const size_t seqPasses = (size_t)(max_in_flight < kMaxPasses ? max_in_flight : kMaxPasses);
        const size_t nSlots = seqPasses * c->nCtus * kSlotsPerCtu;
        CTX_TRY(cudaMalloc(&c->dState, nSlots * sizeof(CuState)));
        c->seqSlots = nSlots;
        CTX_TRY(cudaMalloc(&c->dAccum, 2 * nSlots * sizeof(CuAccum)));
        if (nSlots >= 0x7fffffffull) {
            ame_destroy(c);
            return fail(AME_E_INVALID, "ame_create: %zu CU slots per launch sequence exceed the 31-bit index of the work lists", nSlots);
        }
        c->bigCap = seqPasses * c->nCtus * 9;
        CTX_TRY(cudaMalloc(&c->dGwOut, 2 * nSlots * sizeof(unsigned)));
        CTX_TRY(cudaMalloc(&c->dRowTab, 2 * seqPasses * c->nCtus * sizeof(unsigned)));
        CTX_TRY(cudaMalloc(&c->dTab0, (size_t)kIter0MaxCtas * c->numSMs * kTab0Ints * sizeof(int)));
        CTX_TRY(cudaMalloc(&c->dWork, sizeof(WorkLists) * kMaxSteps));
        CTX_TRY(cudaMalloc(&c->dScan, 3 * scan_words(nSlots) * sizeof(unsigned long long)));
        CTX_TRY(cudaMalloc(&c->dTele, sizeof(Telemetry)));
        CTX_TRY(cudaMemset(c->dTele, 0, sizeof(Telemetry)));
        CTX_TRY(cudaMalloc(&c->dSmallList, 2 * nSlots * sizeof(uint4)));
        CTX_TRY(cudaMalloc(&c->dBigList, 2 * c->bigCap * sizeof(uint2)));
    }
    CTX_TRY(cudaMalloc(&c->dPasses, sizeof(PassDesc) * max_in_flight));
    CTX_TRY(cudaHostAlloc(&c->hPasses, sizeof(PassDesc) * max_in_flight, cudaHostAllocDefault));
    {   // packed CU word of every slot (aligned result index 0..200, then half-aligned 0..283)
        std::vector<uint32_t> slotTab;
        for (int ha = 0; ha < 2; ha++)
            for (const CuDesc &d : ctu_cus(ha)) slotTab.push_back(pack_cu(d));
        CTX_TRY(cudaMalloc(&c->dSlotTab, sizeof(uint32_t) * slotTab.size()));
        CTX_TRY(cudaMemcpy(c->dSlotTab, slotTab.data(), sizeof(uint32_t) * slotTab.size(), cudaMemcpyHostToDevice));
    }
#undef CTX_TRY
    *out = c;
    return AME_OK;
}

int ame_result_len(const ame_ctx *c, int pred) {
    if (!c || pred < 0 || pred >= AME_N_PREDS) return fail(AME_E_INVALID, "ame_result_len: bad argument");
    return (int)c->lens[pred];
}

int ame_set_option(ame_ctx *c, int option, int value) {
    if (!c) return fail(AME_E_INVALID, "ame_set_option: ctx is NULL");
    switch (option) {
        case AME_OPT_CVT_RULE: c->cvtRule = value ? 1 : 0; return AME_OK;
        case AME_OPT_FUSED_BACKSUB: c->fusedBacksub = value ? 1 : 0; return AME_OK;
        case AME_OPT_EARLY_EXIT: c->earlyExit = value ? 1 : 0; return AME_OK;
        case AME_OPT_REUSE_START: c->reuseStart = value ? 1 : 0; return AME_OK;
        case AME_OPT_SHARE_FIRST: c->shareFirst = value ? 1 : 0; return AME_OK;
        case AME_OPT_BIG_TMA: c->bigTma = value ? 1 : 0; return AME_OK;
        case AME_OPT_GROUP_BY_REF: c->groupByRef = value ? 1 : 0; return AME_OK;
    }
    return fail(AME_E_INVALID, "ame_set_option: unknown option %d", option);
}

// The four tensor maps of a slot's edge-replicated plane (2-D, uint16, row pitch padStride; boxes of tma_box(i) x
// tma_box(j) samples, no swizzle) go to device memory, from where ame_iter_big<true> hands them to cp.async.bulk.tensor.
// cuTensorMapEncodeTiled is the one driver-API call of the library; it is looked up through the runtime.
static int encode_tensor_maps(ame_ctx *c, int slot) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return fail(AME_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    CUtensorMap maps[4];
    const cuuint64_t dims[2] = {(cuuint64_t)c->padStride, (cuuint64_t)(c->H + 2 * kPad)};
    const cuuint64_t strides[1] = {(cuuint64_t)c->padStride * sizeof(uint16_t)};
    const cuuint32_t elemStrides[2] = {1, 1};
    for (int k = 0; k < 4; k++) {
        const cuuint32_t box[2] = {(cuuint32_t)tma_box((k >> 1) != 0), (cuuint32_t)tma_box((k & 1) != 0)};
        const CUresult r = encode(&maps[k], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, c->slots[slot].refPad, dims, strides, box, elemStrides, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(AME_E_CUDA, "cuTensorMapEncodeTiled (slot %d, box %d x %d): error %d", slot, (int)box[0], (int)box[1], (int)r);
    }
    CU_TRY(cudaMemcpyAsync(c->dTmaps + (size_t)slot * sizeof maps, maps, sizeof maps, cudaMemcpyHostToDevice, c->up));  // (pageable source: staged before the call returns)
    return AME_OK;
}

// plane == nullptr: the raw plane already in the slot is prepared again (ame_prepare_plane)
static int upload_or_prepare(ame_ctx *c, int slot, const uint16_t *plane, int roles, const char *name) {
    if (slot < 0 || slot >= c->numSlots) return fail(AME_E_INVALID, "%s: slot %d out of range", name, slot);
    if (!(roles & (AME_ROLE_CURRENT | AME_ROLE_REFERENCE))) return fail(AME_E_INVALID, "%s: no role given", name);
    CU_TRY(cudaSetDevice(c->device));
    // Uploads run on their own stream so that they overlap the search kernels of earlier batches.  A slot that
    // queued searches still refer to is launched first; a slot that launched searches may still be reading makes
    // the upload stream wait for those kernels.
    bool queuedUse = false, inflightUse = false;
    for (const Pending &p : c->queued) queuedUse |= (p.curSlot == slot || p.refSlot == slot);
    if (queuedUse) { int rc = ame_flush(c); if (rc) return rc; }
    for (const Pending &p : c->inflight) inflightUse |= (p.curSlot == slot || p.refSlot == slot);
    if (inflightUse && c->kernelsRecorded) CU_TRY(cudaStreamWaitEvent(c->up, c->evKernels, 0));
    Slot &s = c->slots[slot];
    if (plane) {
        CU_TRY(cudaMemcpyAsync(s.raw, plane, (size_t)c->W * c->H * sizeof(uint16_t), cudaMemcpyHostToDevice, c->up));
        s.hasRaw = true;
        s.hasRef = s.hasCur = false;
    } else if (!s.hasRaw) {
        return fail(AME_E_STATE, "%s: slot %d holds no plane", name, slot);
    }
    if (roles & AME_ROLE_CURRENT) {
        if (!s.blk) {
            cudaError_t e = cudaMalloc(&s.blk, (size_t)(c->W / 4) * ((c->H + 3) / 4) * 2 * sizeof(uint4));
            if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? AME_E_NOMEM : AME_E_CUDA, "block-ordered plane of slot %d: %s", slot, cudaGetErrorString(e));
        }
        launch_block_plane(s.raw, s.blk, c->W, c->H, c->up);
        CU_TRY(cudaGetLastError());
        s.hasCur = true;
    }
    if (roles & AME_ROLE_REFERENCE) {
        if (!s.refT) {
            cudaError_t e = cudaMalloc(&s.refT, c->planeSetRecs * sizeof(uint4));
            if (e == cudaSuccess) e = cudaMalloc(&s.refPad, c->planeElems * sizeof(uint16_t) + 64);
            if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? AME_E_NOMEM : AME_E_CUDA, "pre-filtered planes of slot %d: %s", slot, cudaGetErrorString(e));
            int rc = encode_tensor_maps(c, slot);
            if (rc) return rc;
        }
        launch_pad(s.raw, s.refPad, c->W, c->H, c->padStride, c->up);
        launch_phase_planes(s.refPad, s.refT, c->W, c->H, c->padStride, c->up);
        CU_TRY(cudaGetLastError());
        s.hasRef = true;
    }
    CU_TRY(cudaEventRecord(c->evUp, c->up));
    return AME_OK;
}

int ame_upload_plane_ex(ame_ctx *c, int slot, const uint16_t *plane, int roles) {
    CHECK_USABLE(c, "ame_upload_plane");
    if (!plane) return fail(AME_E_INVALID, "ame_upload_plane: NULL argument");
    return upload_or_prepare(c, slot, plane, roles, "ame_upload_plane");
}

int ame_prepare_plane(ame_ctx *c, int slot, int roles) {
    CHECK_USABLE(c, "ame_prepare_plane");
    return upload_or_prepare(c, slot, nullptr, roles, "ame_prepare_plane");
}

int ame_upload_plane(ame_ctx *c, int slot, const uint16_t *plane) {
    return ame_upload_plane_ex(c, slot, plane, AME_ROLE_CURRENT | AME_ROLE_REFERENCE);
}

static int queue_search(ame_ctx *c, int cur_slot, int ref_slot, float lambda, int extra_iters, bool toHost, const ame_result *out, int resultIdx) {
    CHECK_USABLE(c, "ame_search");
    if (cur_slot < 0 || cur_slot >= c->numSlots || ref_slot < 0 || ref_slot >= c->numSlots) return fail(AME_E_INVALID, "ame_search: slot out of range");
    if (extra_iters < 0 || extra_iters > kMaxExtraIter) return fail(AME_E_INVALID, "ame_search: extra_iters %d out of range", extra_iters);
    if (!c->queued.empty() && extra_iters != c->queuedExtra) { int rc = ame_flush(c); if (rc) return rc; }  // one launch sequence = one iteration count
    c->queuedExtra = extra_iters;
    if (!c->slots[ref_slot].hasRef) return fail(AME_E_STATE, "ame_search: slot %d was not uploaded with the reference role", ref_slot);
    if (!c->slots[cur_slot].hasCur) return fail(AME_E_STATE, "ame_search: slot %d was not uploaded with the current-frame role", cur_slot);
    if ((int)(c->queued.size() + c->inflight.size()) >= c->maxInFlight) return fail(AME_E_STATE, "ame_search: %d searches already in flight; call ame_sync", c->maxInFlight);
    if (resultIdx < 0) {
        // first result block not used by a queued / in-flight search
        std::vector<char> used(c->maxInFlight, 0);
        for (const Pending &p : c->queued) used[p.resultIdx] = 1;
        for (const Pending &p : c->inflight) used[p.resultIdx] = 1;
        for (int i = 0; i < c->maxInFlight && resultIdx < 0; i++) if (!used[i]) resultIdx = i;
    } else {
        if (resultIdx >= c->maxInFlight) return fail(AME_E_INVALID, "ame_search_device: result_index out of range");
        for (const Pending &p : c->queued) if (p.resultIdx == resultIdx) return fail(AME_E_STATE, "ame_search_device: result block %d busy", resultIdx);
        for (const Pending &p : c->inflight) if (p.resultIdx == resultIdx) return fail(AME_E_STATE, "ame_search_device: result block %d busy", resultIdx);
    }
    Pending pn;
    pn.resultIdx = resultIdx;
    pn.curSlot = cur_slot;
    pn.refSlot = ref_slot;
    pn.toHost = toHost;
    if (toHost) pn.host = *out;
    PassDesc &d = c->hPasses[c->inflight.size() + c->queued.size()];  // slot stays untouched until ame_sync
    d.curBlk = c->slots[cur_slot].blk;
    d.refT = c->slots[ref_slot].refT;
    for (int p = 0; p < 4; p++) { d.cost[p] = c->results[resultIdx].cost[p]; d.cpmvs[p] = c->results[resultIdx].cpmvs[p]; }
    d.lambda = lambda;
    d.extraIter = extra_iters;
    c->queued.push_back(pn);
    return AME_OK;
}

int ame_search(ame_ctx *c, int cur_slot, int ref_slot, float lambda, int extra_iters, const ame_result *out) {
    if (!c || !out) return fail(AME_E_INVALID, "ame_search: NULL argument");
    for (int p = 0; p < 4; p++) if (!out->cost[p] || !out->cpmvs[p]) return fail(AME_E_INVALID, "ame_search: result array %d is NULL", p);
    return queue_search(c, cur_slot, ref_slot, lambda, extra_iters, true, out, -1);
}

int ame_search_device(ame_ctx *c, int cur_slot, int ref_slot, float lambda, int extra_iters, int result_index) {
    if (!c) return fail(AME_E_INVALID, "ame_search_device: ctx is NULL");
    if (result_index < 0) return fail(AME_E_INVALID, "ame_search_device: result_index out of range");
    return queue_search(c, cur_slot, ref_slot, lambda, extra_iters, false, nullptr, result_index);
}

int ame_device_result(ame_ctx *c, int result_index, ame_result *out) {
    if (!c || !out || result_index < 0 || result_index >= c->maxInFlight) return fail(AME_E_INVALID, "ame_device_result: bad argument");
    for (int p = 0; p < 4; p++) { out->cost[p] = (int64_t *)c->results[result_index].cost[p]; out->cpmvs[p] = c->results[result_index].cpmvs[p]; }
    return AME_OK;
}

int ame_flush(ame_ctx *c) {
    CHECK_USABLE(c, "ame_flush");
    if (c->queued.empty()) return AME_OK;
    CU_TRY(cudaSetDevice(c->device));
    const int n = (int)c->queued.size();
    // Descriptor slots [inflight, inflight + n) are not reused before ame_sync, so the copy can be async.
    const size_t first = c->inflight.size();
    // From here on the searches count as launched: whatever happens, they are never launched a second time.
    const std::vector<Pending> batch = c->queued;
    c->inflight.insert(c->inflight.end(), batch.begin(), batch.end());
    c->queued.clear();
    CU_POISON(cudaMemcpyAsync(c->dPasses + first, c->hPasses + first, sizeof(PassDesc) * n, cudaMemcpyHostToDevice, c->stream));
    KParams kp;
    kp.W = c->W; kp.H = c->H; kp.ctuCols = c->ctuCols; kp.nCtus = c->nCtus; kp.padStride = c->padStride;
    kp.nStrips = tile_strips(c->padStride);
    kp.cvtRule = c->cvtRule; kp.fusedBacksub = c->fusedBacksub; kp.earlyExit = c->earlyExit;
    kp.slotTab = c->dSlotTab; kp.extraIter = c->queuedExtra;
    kp.state = c->dState; kp.accum = c->dAccum; kp.accumStride = (unsigned)c->seqSlots;
    kp.reuseStart = c->reuseStart; kp.shareFirst = c->shareFirst; kp.tab0 = c->dTab0; kp.bigTma = c->bigTma;
    kp.work = c->dWork; kp.tele = c->dTele; kp.gwOut = c->dGwOut; kp.rowTab = c->dRowTab;
    for (int b = 0; b < 2; b++) {
        kp.smallList[b] = c->dSmallList + (size_t)b * c->seqSlots;
        kp.bigList[b] = c->dBigList + (size_t)b * c->bigCap;
    }
    CU_POISON(cudaStreamWaitEvent(c->stream, c->evUp, 0));  // every plane uploaded so far is ready (no-op if none)
    CU_POISON(cudaEventRecord(c->evStart, c->stream));
    // One launch sequence per chunk of at most kMaxPasses searches (they share the scratch arrays, in stream order).
    c->lastLaunches = 0;
    static thread_local PassTable pt;
    for (int k0 = 0; k0 < n; k0 += kMaxPasses) {
        const int m = n - k0 < kMaxPasses ? n - k0 : kMaxPasses;
        kp.nPasses = m;
        kp.passes = c->dPasses + first + k0;
        const size_t words = scan_words((size_t)m * c->nCtus * kSlotsPerCtu);  // the three scan arrays of this sequence, back to back
        kp.scanEmit[0] = c->dScan; kp.scanEmit[1] = c->dScan + words; kp.scanPhase = c->dScan + 2 * words;
        {   // rows of the state array: passes that search the same reference plane are interleaved CTU by CTU (stable in pass order)
            std::vector<int> order(m);
            for (int i = 0; i < m; i++) order[i] = i;
            if (c->groupByRef) std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return batch[k0 + a].refSlot < batch[k0 + b].refSlot; });
            const size_t nRows = (size_t)m * c->nCtus;
            std::vector<unsigned> rows(2 * nRows);  // rows, then the units of ame_iter0_kernel: runs of <= kIter0Unit rows of one (reference plane, CTU)
            size_t r = 0, u = nRows;
            for (int i0 = 0; i0 < m;) {
                int i1 = i0;
                while (i1 < m && (c->groupByRef ? batch[k0 + order[i1]].refSlot == batch[k0 + order[i0]].refSlot : i1 == i0)) i1++;
                for (int ctu = 0; ctu < c->nCtus; ctu++) {
                    for (int i = i0; i < i1; i += kIter0Unit) rows[u++] = (unsigned)(r + (i - i0)) | ((unsigned)std::min(kIter0Unit, i1 - i) << 24);
                    for (int i = i0; i < i1; i++) rows[r++] = (unsigned)order[i] | ((unsigned)ctu << 16);
                }
                i0 = i1;
            }
            kp.unitTab = c->dRowTab + nRows;
            kp.nUnits = (int)(u - nRows);
            CU_POISON(cudaMemcpyAsync(c->dRowTab, rows.data(), u * sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));  // (pageable source: staged before the call returns)
        }
        for (int i = 0; i < m; i++) {
            pt.p[i].curBlk = c->hPasses[first + k0 + i].curBlk;
            pt.p[i].refT = c->hPasses[first + k0 + i].refT;
            pt.p[i].refRaw = c->slots[batch[k0 + i].refSlot].raw;
            pt.p[i].tmap = c->dTmaps + (size_t)batch[k0 + i].refSlot * 4 * sizeof(CUtensorMap);
        }
        CU_POISON(launch_search(kp, pt, c->numSMs, c->stream, c->side, c->evFork, c->evJoin, &c->lastLaunches));
    }
    CU_POISON(cudaGetLastError());
    CU_POISON(cudaEventRecord(c->evStop, c->stream));
    CU_POISON(cudaEventRecord(c->evKernels, c->stream));
    c->kernelsRecorded = true;
    CU_POISON(cudaStreamWaitEvent(c->down, c->evKernels, 0));
    c->timed = true;
    for (const Pending &p : batch) {
        if (!p.toHost) continue;
        const ResultBlock &r = c->results[p.resultIdx];
        // Host arrays laid out like the device block (ame_result_bind): one copy instead of eight.
        bool contiguous = true;
        for (int k = 0; k < 4; k++)
            contiguous = contiguous && (char *)p.host.cost[k] == (char *)p.host.cost[0] + c->resOff[k] &&
                         (char *)p.host.cpmvs[k] == (char *)p.host.cost[0] + c->resOff[4 + k];
        if (contiguous) {
            CU_POISON(cudaMemcpyAsync(p.host.cost[0], r.base, c->resBytes, cudaMemcpyDeviceToHost, c->down));
            continue;
        }
        for (int k = 0; k < 4; k++) {
            CU_POISON(cudaMemcpyAsync(p.host.cost[k], r.cost[k], c->lens[k] * sizeof(long long), cudaMemcpyDeviceToHost, c->down));
            CU_POISON(cudaMemcpyAsync(p.host.cpmvs[k], r.cpmvs[k], c->lens[k] * sizeof(ame_cpmvs), cudaMemcpyDeviceToHost, c->down));
        }
    }
    return AME_OK;
}

int ame_sync(ame_ctx *c) {
    CHECK_USABLE(c, "ame_sync");
    int rc = ame_flush(c);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(c->device));
    CU_POISON(cudaStreamSynchronize(c->up));
    CU_POISON(cudaStreamSynchronize(c->stream));
    CU_POISON(cudaStreamSynchronize(c->down));
    c->inflight.clear();
    return AME_OK;
}

uint64_t ame_result_block_bytes(const ame_ctx *c) { return c ? (uint64_t)c->resBytes : 0; }

int ame_result_bind(const ame_ctx *c, void *block, ame_result *out) {
    if (!c || !block || !out) return fail(AME_E_INVALID, "ame_result_bind: NULL argument");
    for (int p = 0; p < 4; p++) {
        out->cost[p] = reinterpret_cast<int64_t *>((char *)block + c->resOff[p]);
        out->cpmvs[p] = reinterpret_cast<ame_cpmvs *>((char *)block + c->resOff[4 + p]);
    }
    return AME_OK;
}

int ame_exec_ns(ame_ctx *c, double ns[4], int reset) {
    CHECK_USABLE(c, "ame_exec_ns");
    if (!ns) return fail(AME_E_INVALID, "ame_exec_ns: NULL argument");
    if (!c->queued.empty() || !c->inflight.empty()) return fail(AME_E_STATE, "ame_exec_ns: searches in flight; call ame_sync first");
    CU_TRY(cudaSetDevice(c->device));
    Telemetry t;
    CU_TRY(cudaMemcpy(&t, c->dTele, sizeof t, cudaMemcpyDeviceToHost));
    // ns[nCP - 2] is shared between the aligned and the half-aligned CUs by the 4x4 evaluations each side stood for
    for (int k = 0; k < 2; k++) {
        const double full = (double)t.subEvals[k], half = (double)t.subEvals[2 + k], all = full + half;
        ns[k] = all > 0 ? (double)t.ns[k] * full / all : 0.0;      // AME_FULL_2CP, AME_FULL_3CP
        ns[2 + k] = all > 0 ? (double)t.ns[k] * half / all : 0.0;  // AME_HALF_2CP, AME_HALF_3CP
    }
    if (reset) CU_TRY(cudaMemset(c->dTele, 0, sizeof t));
    return AME_OK;
}

int ame_last_kernel_ms(ame_ctx *c, float *ms, int *launches) {
    if (!c || !ms) return fail(AME_E_INVALID, "ame_last_kernel_ms: NULL argument");
    if (!c->timed) return fail(AME_E_STATE, "ame_last_kernel_ms: nothing launched yet");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaEventSynchronize(c->evStop));
    CU_TRY(cudaEventElapsedTime(ms, c->evStart, c->evStop));
    if (launches) *launches = c->lastLaunches;
    return AME_OK;
}

int ame_timer_start(ame_ctx *c) {
    CHECK_USABLE(c, "ame_timer_start");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaEventRecord(c->evAux, c->up));            // the start mark follows everything issued so far on all streams
    CU_TRY(cudaStreamWaitEvent(c->stream, c->evAux, 0));
    CU_TRY(cudaEventRecord(c->evAux, c->down));
    CU_TRY(cudaStreamWaitEvent(c->stream, c->evAux, 0));
    CU_TRY(cudaEventRecord(c->evT0, c->stream));
    CU_TRY(cudaStreamWaitEvent(c->up, c->evT0, 0));    // work issued from now on starts after the start mark
    CU_TRY(cudaStreamWaitEvent(c->down, c->evT0, 0));
    return AME_OK;
}

int ame_timer_stop(ame_ctx *c, float *ms) {
    if (!c || !ms) return fail(AME_E_INVALID, "ame_timer_stop: NULL argument");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaEventRecord(c->evAux, c->up));          // the stop mark follows everything on all three streams
    CU_TRY(cudaStreamWaitEvent(c->stream, c->evAux, 0));
    CU_TRY(cudaEventRecord(c->evAux, c->down));
    CU_TRY(cudaStreamWaitEvent(c->stream, c->evAux, 0));
    CU_TRY(cudaEventRecord(c->evT1, c->stream));
    CU_TRY(cudaEventSynchronize(c->evT1));
    CU_TRY(cudaEventElapsedTime(ms, c->evT0, c->evT1));
    return AME_OK;
}

/* Not part of the public header: development counters (see ame_kernels.cu, -DAME_STATS). */
int ame_debug_stats(unsigned long long *out24, int reset) {
    debug_stats(out24, reset != 0);
    return AME_OK;
}

int ame_debug_div_check(unsigned long long n, unsigned long long seed, unsigned long long *mismatches) {
    if (!mismatches) return fail(AME_E_INVALID, "ame_debug_div_check: NULL argument");
    return debug_div_check(n, seed, mismatches) ? fail(AME_E_CUDA, "ame_debug_div_check: CUDA error") : AME_OK;
}

}  // extern "C"

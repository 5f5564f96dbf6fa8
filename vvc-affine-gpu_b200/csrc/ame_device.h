// Internal interface between the C-ABI layer (ame_api.cu) and the kernels
// (ame_kernels.cu).  Not installed; the public boundary is include/affine_me.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/affine_me.h"

namespace ame {

// Edge replication margin of the motion-compensation copy of a plane.  clipMv
// keeps CU origin + MV inside [-135, W+7] (aux_functions.cl:51-67); a sub-block
// sits up to 124 px further and the filter reaches 3/4 px beyond it, so 144 is
// the worst case; 160 keeps rows 64-byte aligned.
constexpr int kPad = 160;

// Per-CU search state and per-iteration accumulators (ame_iter_kernel / ame_update_kernel).
// Slot k of a CTU: aligned CUs 0..200 (result index), half-aligned CUs 201..484.
constexpr int kSlotsPerCtu = AME_ALIGNED_CUS_PER_CTU + AME_HALF_CUS_PER_CTU;
struct CuState {
    int cur[6];          // CPMVs evaluated by the next / current iteration (ltx, lty, rtx, rty, lbx, lby)
    int best[6];
    long long bestCost;
    int h1[6], h2[6];    // the two states evaluated before cur
    int hasMom, wbuf;    // wbuf: accumulator (0 / 1) the next evaluation writes; hasMom: accumulator wbuf ^ 1 belongs to `best`
};
struct CuAccum {
    long long mom[24];   // moments of the normal equations (numbering of kMomOf in ame_kernels.cu)
    int satd, pad;
};

// One queued search, as the kernels see it.
struct PassDesc {
    const uint4 *curBlk;     // current plane in 4x4-block order (launch_block_plane): 2 x uint4 per block
    const uint4 *refT;       // first-stage rows of the reference (launch_phase_planes): [2 copies][16 phases] planes of
                             // (H + 2*kPad) rows x padStride/8 records of eight int16, (0,0) of the frame at sample [kPad][kPad]
    long long *cost[4];
    ame_cpmvs *cpmvs[4];
    float lambda;
    int extraIter;
};

// The two plane pointers of every pass of a launch sequence, passed to the iteration kernels by value (constant bank).
constexpr int kMaxPasses = 500;
struct PassPtrs {
    const uint4 *curBlk;
    const uint4 *refT;
    const uint16_t *refRaw;  // the reference plane as uploaded (W x H), for ame_iter0_kernel
};
struct PassTable {
    PassPtrs p[kMaxPasses];
};

// Counters of the work lists (see ame_emit_kernel in ame_kernels.cu).
struct WorkLists {
    unsigned nSmall, nBig, nUpd;    // entries
    unsigned nextSmall, nextBig, pad[3];  // tickets handed out by the running ame_iter_* launch
};

constexpr int kIter0MaxCtas = 4;  // resident CTAs per SM ame_iter0_kernel may be built for (AME_ITER0_CTAS)

struct KParams {
    int W, H, ctuCols, nCtus, padStride;
    size_t planeRecs;   // (padStride / 8) * (H + 2*kPad): 16-byte records per (copy, phase) plane
    int nPasses;
    int cvtRule, fusedBacksub, earlyExit;
    const PassDesc *passes;   // device array [nPasses]
    const uint32_t *slotTab;  // device array [kSlotsPerCtu]: packed CU word of every slot
    CuState *state;           // [nPasses * nCtus * kSlotsPerCtu] search state, pass-major
    CuAccum *accum;           // [2][accumStride], same indexing: SATD and moments of the iteration in flight / of the best state
    unsigned accumStride;
    unsigned char *goFlag;    // same indexing: 1 = the CU is evaluated by the next iteration, 2 = it only takes part in the update
    uint4 *blockCnt, *blockOff;  // per 128-CU block of the state array: teams (one-warp, one-CTA, update-only) it contributes / offsets
    WorkLists *work;          // sizes and ticket counters of the lists
    uint4 *smallList;         // capacity: one entry per CU
    uint2 *bigList;
    uint2 *updList;           // CUs that skip the evaluation of the next iteration
    int *tab0;                // scratch of ame_iter0_kernel: [kIter0MaxCtas * numSMs CTAs][1024 * 45 + 1024]
    int shareFirst;           // 1: the first evaluation of all 2-CP searches is shared per sub-block (ame_iter0_kernel)
    int reuseStart;           // 1: the 3-CP search reuses the evaluation of the best 2-CP state where the motion fields agree
    int extraIter;            // --ExtraGradientIter of this batch (uniform per launch sequence)
};

// Launches the search kernels (ame_phase_kernel / ame_iter_kernel / ame_update_kernel, one iteration per launch)
// for all passes (at most kMaxPasses) on `stream`; returns launches made.
int launch_search(const KParams &kp, const PassTable &pt, int numSMs, cudaStream_t stream, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join);
// dst (padded, stride padStride) <- edge-replicated src (W x H).
void launch_pad(const uint16_t *src, uint16_t *dst, int W, int H, int padStride, cudaStream_t stream);
// refT[2][16][H + 2*kPad][padStride/8] <- first interpolation stage of the padded plane `pad`, all 16 phases, int16, in
// 16-byte records; the second copy is shifted by four columns.
void launch_phase_planes(const uint16_t *pad, uint4 *refT, int W, int H, int padStride, cudaStream_t stream);
// blk <- src (W x H) in 4x4-block order (32 bytes per block, (W/4) x ceil(H/4) blocks).
void launch_block_plane(const uint16_t *src, uint4 *blk, int W, int H, cudaStream_t stream);

// Development counters of builds with -DAME_STATS (zeros otherwise).
void debug_stats(unsigned long long *out24, bool reset);

// Development check of the shared-divisor division of the update kernel against __ddiv_rn (see ame_kernels.cu).
int debug_div_check(unsigned long long n, unsigned long long seed, unsigned long long *mismatches);

}  // namespace ame

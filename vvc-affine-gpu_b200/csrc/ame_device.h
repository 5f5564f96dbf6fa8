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
    int done, pad;
};
struct CuAccum {
    long long mom[24];   // moments of the normal equations (numbering of kMomOf in ame_kernels.cu)
    int satd, pad;
};

// One queued search, as the kernels see it.
struct PassDesc {
    const uint4 *curBlk;     // current plane in 4x4-block order (launch_block_plane): 2 x uint4 per block
    const uint2 *refT;       // first-stage rows of the reference (launch_phase_planes): [4 copies][16 phases] planes of
                             // (H + 2*kPad) rows x padStride/4 records, (0,0) of the frame at sample [kPad][kPad]
    long long *cost[4];
    ame_cpmvs *cpmvs[4];
    float lambda;
    int extraIter;
    CuState *state;  // [nCtus * kSlotsPerCtu]
    CuAccum *accum;  // [nCtus * kSlotsPerCtu]
};

struct KParams {
    int W, H, ctuCols, nCtus, padStride;
    size_t planeRecs;   // (padStride / 4) * (H + 2*kPad): 8-byte records per (copy, phase) plane
    int nPasses;
    int cvtRule, fusedBacksub, earlyExit;
    const PassDesc *passes;   // device array [nPasses]
    const uint32_t *bigTab;   // device array [nBig] packed CU words
    const uint2 *smallTab;    // device array [nSmall] (first, second) packed CU words
    int nBig, nSmall;
    const uint32_t *slotTab;  // device array [kSlotsPerCtu]: packed CU word of every slot
    int extraIter;            // --ExtraGradientIter of this batch (uniform per launch sequence)
};

// Launches the search kernels (ame_phase_kernel / ame_iter_kernel / ame_update_kernel, one iteration per launch)
// for all passes on `stream`; returns launches made.
int launch_search(const KParams &kp, cudaStream_t stream, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join);
// dst (padded, stride padStride) <- edge-replicated src (W x H).
void launch_pad(const uint16_t *src, uint16_t *dst, int W, int H, int padStride, cudaStream_t stream);
// refT[4][16][H + 2*kPad][padStride/4] <- first interpolation stage of the padded plane `pad`, all 16 phases, in four
// column alignments (8-byte records of four int16).
void launch_phase_planes(const uint16_t *pad, uint2 *refT, int W, int H, int padStride, cudaStream_t stream);
// blk <- src (W x H) in 4x4-block order (32 bytes per block, (W/4) x ceil(H/4) blocks).
void launch_block_plane(const uint16_t *src, uint4 *blk, int W, int H, cudaStream_t stream);

// Development counters of builds with -DAME_STATS (zeros otherwise).
void debug_stats(unsigned long long *out24, bool reset);

}  // namespace ame

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

// Tiled layout of the 2 x 16 pre-filtered planes of a reference (PassDesc::refT).  A tile = 128 rows (+ 8 rows
// repeated from the next tile, so that the nine rows of a sub-block window never leave the tile of their first row) x
// one strip of kStripRecs records (32 columns = 64 bytes per row: a 128-byte line holds two rows of the strip, the
// nine rows of a window are 5 lines that mostly serve this sub-block and its neighbours of the same CU); the 32 planes of
// a tile are adjacent.  What the CUs of a CTU touch under all phases is then a few contiguous pieces of 278 KB instead
// of 32 x 170 row segments that lie a plane row (4..16 KB) apart: a handful of pages for the TLBs (measured: 1080p
// +12 %, 4K +59 %, 8K +75 % against the row-major planes; 64-byte rows against 128-byte rows: +8 %).  Within a
// tile, consecutive rows are kStripRecs records apart whatever the frame width (immediate offsets for the nine loads).
#ifndef AME_STRIP_SHIFT
#define AME_STRIP_SHIFT 2
#endif
constexpr int kTileRows = 128, kTileHalo = 8;
constexpr int kStripShift = AME_STRIP_SHIFT, kStripRecs = 1 << kStripShift;  // records (of 8 columns) per tile row
constexpr int kTileRecs = (kTileRows + kTileHalo) * kStripRecs;            // records of one plane of one tile
__host__ __device__ inline int tile_strips(int padStride) { return ((padStride >> 3) + kStripRecs - 1) >> kStripShift; }
__host__ __device__ inline int tile_row_blocks(int padRows) { return (padRows + kTileRows - 1) / kTileRows; }
// record index of (plane, row, record column) for a window whose FIRST row is `row` (or for row itself)
__host__ __device__ inline unsigned tile_record(int nStrips, int plane, int row, int rec) {
    return (unsigned)(((row >> 7) * nStrips + (rec >> kStripShift)) * 32 + plane) * (unsigned)kTileRecs +
           (unsigned)((row & 127) * kStripRecs + (rec & (kStripRecs - 1)));
}
__host__ __device__ inline size_t tiled_plane_set_recs(int padStride, int padRows) {
    return (size_t)tile_row_blocks(padRows) * tile_strips(padStride) * 32 * kTileRecs;
}

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
    const uint4 *refT;       // first-stage rows of the reference (launch_phase_planes): 2 copies x 16 phases of
                             // (H + 2*kPad) rows x padStride/8 records of eight int16, (0,0) of the frame at sample
                             // [kPad][kPad], stored in tiles (see tile_record)
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
    int nStrips;        // 64-column strips per row of the pre-filtered planes (tile_record)
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
// refT (tiled, tile_record) <- first interpolation stage of the padded plane `pad`, all 16 phases, int16, in
// 16-byte records; the second copy (planes 16..31) is shifted by four columns.
void launch_phase_planes(const uint16_t *pad, uint4 *refT, int W, int H, int padStride, cudaStream_t stream);
// blk <- src (W x H) in 4x4-block order (32 bytes per block, (W/4) x ceil(H/4) blocks).
void launch_block_plane(const uint16_t *src, uint4 *blk, int W, int H, cudaStream_t stream);

// Development counters of builds with -DAME_STATS (zeros otherwise).
void debug_stats(unsigned long long *out24, bool reset);

// Development check of the shared-divisor division of the update kernel against __ddiv_rn (see ame_kernels.cu).
int debug_div_check(unsigned long long n, unsigned long long seed, unsigned long long *mismatches);

}  // namespace ame

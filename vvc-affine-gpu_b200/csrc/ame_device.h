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
    const void *tmap;        // four CUtensorMap (128 bytes each, device memory) over the edge-replicated reference plane
                             // (padStride x (H + 2*kPad) uint16): boxes of tma_box(i) x tma_box(j) samples at index 2*i + j, for ame_iter_big<true>
};
struct PassTable {
    PassPtrs p[kMaxPasses];
};

// Work lists of one step of a launch sequence.  A step = one evaluation (ame_iter_* / ame_iter0_kernel) + one
// ame_update_kernel; its lists are written by ame_phase_kernel (first step of a search) or by the ame_emit_kernel
// launch behind the update of the step before, and every step has its own counters, zeroed once per sequence, so
// nothing has to be reset between launches.  Entry formats: see "Work lists" in ame_kernels.cu.
struct WorkLists {
    unsigned nSmall, nBig;                 // entries of the step's two lists
    unsigned nextSmall, nextBig;           // tickets handed out by the step's ame_iter_* launches
    unsigned nextChunk;                    // chunk numbers of the ame_emit_kernel launch that reads the step's lists
    unsigned nextPhaseChunk;               // chunk numbers of the ame_phase_kernel launch that writes the step's lists
    unsigned pad[2];
};
constexpr int kMaxExtraIter = 64;
constexpr int kMaxSteps = (6 + kMaxExtraIter) + (5 + kMaxExtraIter) + 1;

// Device-side record of where the time of the launch sequences went (ame_exec_ns): %globaltimer at the start of the
// three ame_phase_kernel launches of a sequence gives the duration of its 2-CP and 3-CP searches; the update kernels
// count the 4x4 sub-block evaluations per prediction type, by which the host splits the two durations between the
// aligned and the half-aligned CUs (all four types run fused; main_aux_functions.h:1416-1446 reports them one by one).
struct Telemetry {
    unsigned long long mark[3];
    unsigned long long ns[2];       // accumulated: 2-CP searches, 3-CP searches
    unsigned long long subEvals[4]; // accumulated: FULL_2CP, FULL_3CP, HALF_2CP, HALF_3CP
};

constexpr int kIter0MaxCtas = 4;  // resident CTAs per SM ame_iter0_kernel may be built for (AME_ITER0_CTAS)
// scratch of one CTA of ame_iter0_kernel (ints): [sub-block][case][sum] + [sub-block] SATD + [CU][18] int64 moments shared by the searches of a unit
constexpr int kTab0Ints = 1024 * 45 + 1024 + (AME_ALIGNED_CUS_PER_CTU + AME_HALF_CUS_PER_CTU) * 18 * 2;
constexpr int kIter0Unit = 8;     // searches of one (reference plane, CTU) ame_iter0_kernel evaluates in one turn

struct KParams {
    int W, H, ctuCols, nCtus, padStride;
    int nStrips;        // 64-column strips per row of the pre-filtered planes (tile_record)
    int nPasses;
    int cvtRule, fusedBacksub, earlyExit;
    const PassDesc *passes;   // device array [nPasses]
    const uint32_t *slotTab;  // device array [kSlotsPerCtu]: packed CU word of every slot
    const unsigned *rowTab;   // device array [nPasses * nCtus]: pass | ctu << 16 of every row of the state array.  The order of the rows
                              // is the order of the work lists: passes that search the SAME reference plane are interleaved CTU by
                              // CTU, so that the warps resident at one time work on one region of one reference plane
    const unsigned *unitTab;  // device array [nUnits]: first row | number of rows << 24 of the runs of up to kIter0Unit rows of rowTab
                              // that belong to one CTU and one reference plane (ame_iter0_kernel)
    int nUnits;
    CuState *state;           // [nPasses * nCtus][kSlotsPerCtu] search state, rows in rowTab order
    CuAccum *accum;           // [2][accumStride], same indexing: SATD and moments of the iteration in flight / of the best state
    unsigned accumStride;
    WorkLists *work;          // [kMaxSteps] sizes and ticket counters of the lists of every step
    uint4 *smallList[2];      // lists of step s: buffer s & 1; capacity: one entry per CU
    uint2 *bigList[2];
    unsigned *gwOut;          // [2 * CUs] per list position of the step: the CU's entry word for the next step, or kNone (ame_update_kernel -> ame_emit_kernel)
    unsigned long long *scanEmit[2];  // ordered compaction (decoupled look-back) of ame_emit_kernel: one word per 128-CU chunk, buffer s & 1
    unsigned long long *scanPhase;    // same for ame_phase_kernel
    Telemetry *tele;
    int *tab0;                // scratch of ame_iter0_kernel: [kIter0MaxCtas * numSMs CTAs][kTab0Ints]
    int shareFirst;           // 1: the first evaluation of all 2-CP searches is shared per sub-block (ame_iter0_kernel)
    int reuseStart;           // 1: the 3-CP search reuses the evaluation of the best 2-CP state where the motion fields agree
    int extraIter;            // --ExtraGradientIter of this batch (uniform per launch sequence)
    int bigTma;               // 1: big CUs stage their search window in shared memory with TMA (ame_iter_big<true>)
};
// box extents (samples) of the tensor maps of a reference plane, for a CU extent of 64 / 128: see ame_iter_big<true>
constexpr int kTmaBoxSmall = 96, kTmaBoxBig = 160;
__host__ __device__ constexpr int tma_box(bool extent128) { return extent128 ? kTmaBoxBig : kTmaBoxSmall; }

// Launches the search kernels (ame_phase_kernel / ame_iter_* / ame_update_kernel, one iteration per launch) for all
// passes (at most kMaxPasses) on `stream`.  Returns cudaSuccess or the first error of a launch / attribute call;
// *launches += kernels launched.
cudaError_t launch_search(const KParams &kp, const PassTable &pt, int numSMs, cudaStream_t stream, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join,
                          int *launches);
// Number of 128-CU chunks the ordered compaction of a sequence of nSlots CU slots may use: words of each of the three
// KParams::scan* arrays, which launch_search expects back to back starting at scanEmit[0].
// (a one-warp CU takes two places of the update kernel's index space, hence 2 x)
inline size_t scan_words(size_t nSlots) { return (2 * nSlots + 127) / 128 + 1; }
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

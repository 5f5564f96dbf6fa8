// CU geometry of one 128x128 CTU: the 201 aligned and 284 half-aligned CUs the
// reference evaluates (constants.cl:73-141 and :207-435), generated from grid
// rules instead of position tables, and the order in which this implementation
// schedules them.  Host-side code; the kernels only see the packed task table.
#pragma once
#include <stdint.h>

#include <vector>

namespace ame {

struct CuDesc {
    int x, y, w, h;  // inside the CTU
    int ha;          // 0 = aligned, 1 = half-aligned
    int group;       // size-group index (0..11 / 0..23)
    int idx;         // result index inside the CTU (0..200 / 0..283)
};

// A grid axis: n positions start, start+step, ...
struct Axis { int start, step, n; };

struct HaGroup { int w, h; Axis ax, ay; };

// Aligned sizes in the reference's result order (constants.cl:74-113).
static const int kAlignedW[12] = {128, 128, 64, 64, 64, 32, 32, 64, 16, 32, 16, 16};
static const int kAlignedH[12] = {128, 64, 128, 64, 32, 64, 32, 16, 64, 16, 32, 16};

// Half-aligned groups 0..22 are regular grids (x fastest, then y), group 23 is irregular.
static const HaGroup kHaGroups[23] = {
    {64, 32, {0, 64, 2}, {16, 64, 2}},  {32, 64, {16, 64, 2}, {0, 64, 2}},  {64, 16, {0, 64, 2}, {8, 32, 4}},
    {64, 16, {0, 64, 2}, {24, 64, 2}},  {16, 64, {8, 32, 4}, {0, 64, 2}},   {16, 64, {24, 64, 2}, {0, 64, 2}},
    {32, 32, {16, 64, 2}, {0, 32, 4}},  {32, 32, {0, 32, 4}, {16, 64, 2}},  {32, 16, {0, 32, 4}, {8, 32, 4}},
    {32, 16, {0, 32, 4}, {24, 64, 2}},  {32, 16, {16, 64, 2}, {0, 16, 8}},  {16, 32, {8, 32, 4}, {0, 32, 4}},
    {16, 32, {24, 64, 2}, {0, 32, 4}},  {16, 32, {0, 16, 8}, {16, 64, 2}},  {16, 16, {0, 16, 8}, {8, 32, 4}},
    {16, 16, {8, 32, 4}, {0, 16, 8}},   {16, 16, {0, 16, 8}, {24, 64, 2}},  {16, 16, {24, 64, 2}, {0, 16, 8}},
    {32, 32, {16, 64, 2}, {16, 64, 2}}, {32, 16, {16, 64, 2}, {8, 32, 4}},  {32, 16, {16, 64, 2}, {24, 64, 2}},
    {16, 32, {8, 32, 4}, {16, 64, 2}},  {16, 32, {24, 64, 2}, {16, 64, 2}}};

// Group 23 ("16x16 U123"): rows of 6 / 4 / 6 CUs in each CTU half.
inline void ha_group23(std::vector<CuDesc> &out, int &idx) {
    static const int x6[6] = {8, 24, 40, 72, 88, 104};
    static const int x4[4] = {8, 40, 72, 104};
    static const int ys[6] = {8, 24, 40, 72, 88, 104};
    for (int r = 0; r < 6; r++) {
        const bool four = (r % 3) == 1;
        const int n = four ? 4 : 6;
        for (int c = 0; c < n; c++) out.push_back({four ? x4[c] : x6[c], ys[r], 16, 16, 1, 23, idx++});
    }
}

// All CUs of a CTU for one alignment class, in result-index order.
inline std::vector<CuDesc> ctu_cus(int ha) {
    std::vector<CuDesc> out;
    int idx = 0;
    if (!ha) {
        for (int g = 0; g < 12; g++) {
            const int w = kAlignedW[g], h = kAlignedH[g];
            for (int y = 0; y < 128; y += h)
                for (int x = 0; x < 128; x += w) out.push_back({x, y, w, h, 0, g, idx++});
        }
    } else {
        for (int g = 0; g < 23; g++) {
            const HaGroup &G = kHaGroups[g];
            for (int j = 0; j < G.ay.n; j++)
                for (int i = 0; i < G.ax.n; i++)
                    out.push_back({G.ax.start + i * G.ax.step, G.ay.start + j * G.ay.step, G.w, G.h, 1, g, idx++});
        }
        ha_group23(out, idx);
    }
    return out;
}

// Packed CU descriptor handed to the kernels (one 32-bit word):
//   bits 0-3 x/8, 4-7 y/8, 8-9 log2(w)-4, 10-11 log2(h)-4, 12 ha, 13-21 idx, 31 valid
inline uint32_t pack_cu(const CuDesc &c) {
    auto lg = [](int v) { int r = 0; while (v > 1) { v >>= 1; r++; } return r; };
    return (uint32_t)(c.x / 8) | ((uint32_t)(c.y / 8) << 4) | ((uint32_t)(lg(c.w) - 4) << 8) |
           ((uint32_t)(lg(c.h) - 4) << 10) | ((uint32_t)c.ha << 12) | ((uint32_t)c.idx << 13) | 0x80000000u;
}

// Schedule of one CTU:
//   big   : CUs with >= 256 4x4 sub-blocks, one 256-thread CTA each (9 per CTU)
//   small : everything else, one warp each, largest first; the smallest CUs are paired
//           (two CUs per warp, one per half-warp).  second == 0 means "no partner".
struct SmallTask { uint32_t first, second; };
struct CtuSchedule {
    std::vector<uint32_t> big;
    std::vector<SmallTask> small;
};

inline CtuSchedule build_schedule(int pairMaxArea = 256) {
    CtuSchedule s;
    std::vector<CuDesc> all = ctu_cus(0);
    std::vector<CuDesc> h = ctu_cus(1);
    all.insert(all.end(), h.begin(), h.end());
    for (int area = 128 * 128; area >= 256; area >>= 1) {
        // CUs of one area class; those up to pairMaxArea are paired (two CUs of the same shape per warp, one per
        // half warp) with their successor of the same alignment class and size group, i.e. a geometric neighbour.
        std::vector<CuDesc> cls;
        for (const CuDesc &c : all)
            if (c.w * c.h == area) cls.push_back(c);
        for (size_t i = 0; i < cls.size(); i++) {
            const CuDesc &c = cls[i];
            if (area >= 64 * 64) {
                s.big.push_back(pack_cu(c));
            } else if (area <= pairMaxArea && i + 1 < cls.size() && cls[i + 1].ha == c.ha && cls[i + 1].group == c.group) {
                s.small.push_back({pack_cu(c), pack_cu(cls[i + 1])});
                i++;
            } else {
                s.small.push_back({pack_cu(c), 0u});
            }
        }
    }
    return s;
}

}  // namespace ame

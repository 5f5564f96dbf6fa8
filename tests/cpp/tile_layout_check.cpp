// CPU check of the tiled layout of the pre-filtered reference planes (vvc-affine-gpu_b200/csrc/ame_device.h):
// what phase_kernel writes (every row at its own position, rows 0..7 of a tile a second time as rows 128..135 of the
// tile above) is what a sub-block window reads (first row through tile_record, the next eight rows kStripRecs records
// further each).  Usage: tile_layout_check W H  -> exit 0 and "ok ..." or a message and exit 1.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../vvc-affine-gpu_b200/csrc/ame_device.h"

using namespace ame;

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    const int W = atoi(argv[1]), H = atoi(argv[2]);
    const int padStride = W + 2 * kPad, padRows = H + 2 * kPad, recsPerRow = padStride >> 3;
    const int nStrips = tile_strips(padStride);
    const size_t total = tiled_plane_set_recs(padStride, padRows);
    if (total > 0xffffffffull) { printf("plane set exceeds the 32-bit record index\n"); return 1; }
    // owner[idx] = 1 + (plane, row, rec) packed, as written by the phase kernel
    std::vector<unsigned long long> owner(total, 0);
    auto key = [&](int plane, int row, int rec) { return 1ull + ((unsigned long long)plane * padRows + row) * recsPerRow + rec; };
    for (int plane = 0; plane < 32; plane += 5)  // (a sample of the planes keeps the check fast; the formula is linear in plane)
        for (int row = 0; row < padRows; row++)
            for (int rec = 0; rec < recsPerRow; rec++) {
                const size_t idx = tile_record(nStrips, plane, row, rec);
                if (idx >= total) { printf("index out of range: plane %d row %d rec %d\n", plane, row, rec); return 1; }
                if (owner[idx]) { printf("two records share index %zu\n", idx); return 1; }
                owner[idx] = key(plane, row, rec);
                if ((row & 127) < kTileHalo && row >= kTileRows) {  // the halo copy (phase_kernel: h0 / h1)
                    const size_t h = (size_t)tile_record(nStrips, plane, row - kTileRows, rec) + (size_t)kTileRows * kStripRecs;
                    if (h >= total || owner[h]) { printf("bad halo index for row %d\n", row); return 1; }
                    owner[h] = key(plane, row, rec);
                }
            }
    // every window: first row r0 (all rows a window can start at), nine rows
    size_t windows = 0;
    for (int plane = 0; plane < 32; plane += 5)
        for (int r0 = 0; r0 + 8 < padRows; r0++)
            for (int rec = 0; rec < recsPerRow; rec += 3) {
                const size_t base = tile_record(nStrips, plane, r0, rec);
                for (int j = 0; j < 9; j++)
                    if (owner[base + (size_t)j * kStripRecs] != key(plane, r0 + j, rec)) {
                        printf("window at plane %d row %d rec %d: row %d reads the wrong record\n", plane, r0, rec, j);
                        return 1;
                    }
                windows++;
            }
    printf("ok %dx%d: %zu records, %zu windows checked, %d strips, tile rows of %d records\n", W, H, total, windows, nStrips, kStripRecs);
    return 0;
}

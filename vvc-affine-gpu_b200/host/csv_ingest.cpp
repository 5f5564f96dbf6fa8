#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <thread>

#include "ame_host.h"

namespace host {

// The reference parses with getline + stoi per sample (main.cpp:311-326).  Here the file is mapped, line
// starts are located with memchr and row ranges are parsed in parallel straight into the destination
// plane memory (pinned by the caller).  Same accepted format: H lines per frame, W decimal samples per
// line separated by commas; anything after the W-th value of a line is ignored, like the reference.
int read_csv_frames(const std::string &path, int nFrames, int W, int H, uint16_t *dst, int threads, std::string &err) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) { err = "error while opening samples files: " + path; return -1; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size == 0) { close(fd); err = "empty or unreadable file: " + path; return -1; }
    const size_t size = (size_t)st.st_size;
    const char *data = (const char *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (data == MAP_FAILED) { err = "mmap failed: " + path; return -1; }
    madvise((void *)data, size, MADV_SEQUENTIAL);
    const size_t rows = (size_t)nFrames * H;
    std::vector<size_t> start(rows + 1);
    size_t pos = 0, r = 0;
    for (; r < rows && pos < size; r++) {
        start[r] = pos;
        const char *nl = (const char *)memchr(data + pos, '\n', size - pos);
        pos = nl ? (size_t)(nl - data) + 1 : size;
    }
    if (r < rows) { munmap((void *)data, size); err = "file has fewer than FramesToBeEncoded*height lines: " + path; return -1; }
    start[rows] = pos;
    std::atomic<int> bad(0);
    auto work = [&](size_t r0, size_t r1) {
        for (size_t row = r0; row < r1; row++) {
            const char *p = data + start[row], *end = data + start[row + 1];
            uint16_t *out = dst + row * (size_t)W;
            int col = 0;
            while (col < W && p < end) {
                while (p < end && (*p == ' ' || *p == '\t')) p++;
                unsigned v = 0;
                bool any = false;
                while (p < end && *p >= '0' && *p <= '9') { v = v < 100000u ? v * 10 + (unsigned)(*p - '0') : v; p++; any = true; }
                if (!any || v > 65535u) { bad++; break; }  // (not a number / beyond the reference's unsigned short)
                out[col++] = (uint16_t)v;
                while (p < end && *p != ',' && *p != '\n') p++;
                if (p < end && *p == ',') p++;
            }
            if (col < W) bad++;
        }
    };
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    const size_t chunk = (rows + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        const size_t r0 = t * chunk, r1 = std::min(rows, r0 + chunk);
        if (r0 < r1) pool.emplace_back(work, r0, r1);
    }
    for (auto &t : pool) t.join();
    munmap((void *)data, size);
    if (bad.load()) { err = "malformed sample rows in " + path; return -1; }
    return 0;
}

int read_raw_frames(const std::string &path, int nFrames, int W, int H, uint16_t *dst, std::string &err) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { err = "error while opening samples files: " + path; return -1; }
    const size_t want = (size_t)nFrames * W * H;
    const size_t got = fread(dst, sizeof(uint16_t), want, f);
    fclose(f);
    if (got != want) { err = "file has fewer than FramesToBeEncoded*height*width samples: " + path; return -1; }
    return 0;
}

}  // namespace host

#include <string.h>

#include "ame_host.h"

namespace host {

static const char *kTypeTag[4] = {"_FULL_2CPs_", "_FULL_3CPs_", "_HALF_2CPs_", "_HALF_3CPs_"};
static const char *kHeader = "POC,List,Ref,CTU,idx,X,Y,Cost,LT_X,LT_Y,RT_X,RT_Y,LB_X,LB_Y\n";

// Size groups in result order with their first result index, taken from the library's own geometry.
struct Group { int w, h, first, n; };
static std::vector<Group> groups_of(int pred) {
    std::vector<Group> g;
    const int total = pred < 2 ? AME_ALIGNED_CUS_PER_CTU : AME_HALF_CUS_PER_CTU;
    int last = -1;
    for (int k = 0; k < total; k++) {
        int geo[4];
        const int grp = ame_cu_geometry(pred, k, geo);
        if (grp != last) { g.push_back({geo[2], geo[3], k, 0}); last = grp; }
        g.back().n++;
    }
    return g;
}

LogWriter::LogWriter(const std::string &prefix, int W, int H) : prefix_(prefix), W_(W), H_(H) {
    ctuCols_ = (W + 127) / 128;
    nCtus_ = ame_num_ctus(W, H);
}

LogWriter::~LogWriter() { close(); }

void LogWriter::close() {
    for (auto &v : files_)
        for (File &f : v)
            if (f.f) { fclose(f.f); f.f = nullptr; }
}

// One FILE per (prediction type, size string), opened on first use with the header written once (the
// reference truncates and writes headers at poc==1 && ref==0, main_aux_functions.h:431-456, then appends).
LogWriter::File &LogWriter::file_for(int pred, int w, int h) {
    char nm[32];
    snprintf(nm, sizeof nm, "%dx%d", w, h);
    for (File &f : files_[pred]) if (f.name == nm) return f;
    File f;
    f.name = nm;
    const std::string path = prefix_ + kTypeTag[pred] + nm + ".csv";
    f.f = fopen(path.c_str(), "w");
    if (f.f) {
        setvbuf(f.f, nullptr, _IOFBF, 1 << 20);
        fputs(kHeader, f.f);
    } else {
        fprintf(stderr, "cannot open %s for writing\n", path.c_str());
    }
    files_[pred].push_back(f);
    return files_[pred].back();
}

static inline char *put_int(char *p, long long v) {
    char tmp[24];
    int n = 0;
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

size_t LogWriter::write_pass(int poc, int ref, const ame_result &res) {
    if (!enabled()) return 0;
    size_t rows = 0;
    static const std::vector<Group> groups[4] = {groups_of(0), groups_of(1), groups_of(2), groups_of(3)};
    for (int pred = 0; pred < 4; pred++) {
        printf("Reporting results POC=%d refIdx=%d PredType=%d\n", poc, ref, pred);
        if (poc == 1 && ref == 0) printf("Writing headers\n");
        const int perCtu = pred < 2 ? AME_ALIGNED_CUS_PER_CTU : AME_HALF_CUS_PER_CTU;
        for (const Group &g : groups[pred]) {
            File &f = file_for(pred, g.w, g.h);
            if (!f.f) continue;
            buf_.resize((size_t)nCtus_ * g.n * 160);
            char *p = buf_.data();
            for (int ctu = 0; ctu < nCtus_; ctu++) {
                const int ctuX = (ctu % ctuCols_) * 128, ctuY = (ctu / ctuCols_) * 128;
                for (int i = 0; i < g.n; i++) {
                    int geo[4];
                    ame_cu_geometry(pred, g.first + i, geo);
                    const size_t k = (size_t)ctu * perCtu + g.first + i;
                    const ame_cpmvs &m = res.cpmvs[pred][k];
                    const long long vals[14] = {poc, 0, ref, ctu, i, ctuX + geo[0], ctuY + geo[1], (long long)res.cost[pred][k],
                                                m.ltx, m.lty, m.rtx, m.rty, m.lbx, m.lby};
                    for (int c = 0; c < 14; c++) {
                        p = put_int(p, vals[c]);
                        *p++ = c == 13 ? '\n' : ',';
                    }
                }
            }
            fwrite(buf_.data(), 1, (size_t)(p - buf_.data()), f.f);
            rows += (size_t)nCtus_ * g.n;
        }
    }
    return rows;
}

}  // namespace host

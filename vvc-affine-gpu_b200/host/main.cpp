// affine_b200 -- drop-in for the reference's `./main` (same flags, CSV inputs, decision logs and stdout
// markers; /root/reference/main.cpp:53-1123) with the OpenCL device code replaced by the CUDA library behind
// include/affine_me.h.  No OpenCL, no CPU fallback.
//
// Frame pipeline: both CSV files are parsed into pinned planes; frames are processed in batches of
// --BatchFrames; batch b runs on GPU (DeviceIndex + b % NumDevices), every GPU with its own ame_ctx and host thread
// (frames are mutually independent given the input files); the log writer consumes batches in
// POC order, so the logs are byte-identical to the single-GPU order.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include <condition_variable>
#include <iostream>
#include <map>
#include <mutex>
#include <set>
#include <thread>

#include "ame_host.h"

using namespace host;

struct PassOut {  // pinned result arrays of one (poc, ref) pass
    ame_result res;
    void *block = nullptr;
};

// One pinned block per pass, laid out like the library's result block: one device-to-host copy per search.
static bool alloc_pass(PassOut &p, const ame_ctx *ctx) {
    p.block = ame_alloc_host(ame_result_block_bytes(ctx));
    return p.block && ame_result_bind(ctx, p.block, &p.res) == AME_OK;
}

struct Batch {
    int firstFrame = 0, nFrames = 0;       // frame index = poc - 1
    std::vector<PassOut> passes;           // in (poc, ref) order
    std::vector<std::pair<int, int>> ids;  // (poc, ref)
    bool done = false;
    double execNs[4] = {0, 0, 0, 0};  // device time of the batch's searches per prediction type (its share of the group's ame_exec_ns)
    double tDone = 0;         // host clock when the results of the batch were complete
    int device = 0;
};

static size_t file_size(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return 0;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fclose(f);
    return n > 0 ? (size_t)n : 0;
}

static double now_s() {
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return tv.tv_sec + tv.tv_usec * 1e-6;
}

int main(int argc, char **argv) {
    Options o;
    const int prc = parse_options(argc, argv, o);
    if (prc) return prc - 1000;
    if (o.help) { print_help(); return 1; }
    if (check_report_parameters(o) > 0) {
        std::cout << "Exiting after finding errors in input parameters" << std::endl;
        return 1;
    }
    print_timestamp("START HOST");

    int W = 0, H = 0;
    {
        const size_t x = o.resolution.find('x');
        char *e1 = nullptr, *e2 = nullptr;
        if (x != std::string::npos) {
            W = (int)strtol(o.resolution.c_str(), &e1, 10);
            H = (int)strtol(o.resolution.c_str() + x + 1, &e2, 10);
        }
        if (x == std::string::npos || o.resolution.find('x', x + 1) != std::string::npos || W <= 0 || H <= 0) {
            std::cout << "  [!] ERROR: Input resolution \"" << o.resolution << "\" not set properly" << std::endl;
            return 0;
        }
    }
    // The reference accepts five whitelisted sizes (constants.h:73-79); any size the kernels can address is
    // accepted here (nCtus = ceil(W/128)*ceil(H/128) gives the same counts for the five).
    if (W % 8 != 0 || W < 16 || H < 16) {
        printf("[!] ERROR: Unsupported resolution %dx%d\n", W, H);
        printf("Supported resolutions are: any WxH with W a multiple of 8 and W,H >= 16\n");
        return 0;
    }
    const int N = o.nFrames;
    if (N < 1) { std::cout << "  [!] ERROR: FramesToBeEncoded must be positive" << std::endl; return 1; }
    if (o.extraGradIter < 0 || o.extraGradIter > 64) {
        std::cout << "  [!] ERROR: ExtraGradientIter must be between 0 and 64" << std::endl;
        return 1;
    }
    print_reference_plan(N, o.qp);

    // ---- CSV ingest (main.cpp:293-328) ----
    const size_t plane = (size_t)W * H;
    uint16_t *orig = (uint16_t *)ame_alloc_host(plane * N * sizeof(uint16_t));
    uint16_t *recon = (uint16_t *)ame_alloc_host(plane * N * sizeof(uint16_t));
    if (!orig || !recon) {
        fprintf(stderr, "cannot allocate pinned frame memory (%s); a CUDA device is required, there is no CPU fallback\n", ame_last_error());
        return 1;
    }
    print_timestamp("START READ .csv");
    const double tRead0 = now_s();
    size_t csvBytes = 0;
    {
        std::string e1, e2;
        const int threads = std::max(1u, std::thread::hardware_concurrency());
        int r1 = 0, r2 = 0;
        if (o.rawFrames) {
            r1 = read_raw_frames(o.origFile, N, W, H, orig, e1);
            r2 = read_raw_frames(o.refFile, N, W, H, recon, e2);
        } else {
            std::thread t1([&] { r1 = read_csv_frames(o.origFile, N, W, H, orig, std::max(1, threads / 2), e1); });
            r2 = read_csv_frames(o.refFile, N, W, H, recon, std::max(1, threads / 2), e2);
            t1.join();
        }
        if (r1 || r2) {
            fprintf(stderr, "%s\n", (r1 ? e1 : e2).c_str());
            return 1;
        }
        csvBytes = file_size(o.origFile) + file_size(o.refFile);
    }
    const double readSeconds = now_s() - tRead0;
    print_timestamp("FINISHED READ .csv");

    print_timestamp("START BUILD KERNELS");  // kernels are compiled ahead of time; markers kept for the energy scripts
    print_timestamp("FINISH BUILD KERNELS");

    // ---- contexts, one per GPU ----
    // A GPU works on GROUPS of kGroup batches: every batch of a group is uploaded, searched and flushed with its own plane slots,
    // result blocks and pinned host blocks, then ONE ame_sync waits for the group -- the uploads of a batch and the result
    // copies of the batch before it run beside the kernels.  All device and pinned memory is allocated here, once (the
    // reference allocates its buffers in this stage too, main.cpp:330-359, 484-552): pinned allocation and release synchronise
    // the device, and 250 of each per 64 frames cost seven times the search itself.
    print_timestamp("START ALLOCATE MEMORY");
    const int nDev = std::max(1, o.numDevices);
    // frames per batch: --BatchFrames, but never so many that a GPU is left without a batch
    int B = std::max(1, std::min(o.batchFrames, (N + nDev - 1) / nDev));
    // Reference slots dominate the device memory (2 x 16 pre-filtered planes each: 0.2 GB at 1080p, 2.6 GB at 8K): batches per
    // group, then frames per batch, are cut back until they fit a budget of 64 GB.
    const double slotBytes = 32.0 * (W + 320.0) * (H + 320.0) * 2.0, budget = 64e9;
    while (B > 1 && (B + 8) * slotBytes > budget) B--;
    const int kGroup = (int)std::max(1.0, std::min(4.0, budget / ((B + 8) * slotBytes)));
    const int slotsPerBatch = 2 * B + 8;  // B current planes + up to B+3 reference planes (+ slack)
    const int passesPerBatch = 4 * B;
    std::vector<ame_ctx *> ctxs(nDev, nullptr);
    std::vector<std::vector<PassOut>> pool(nDev);  // [device][2 sets x kGroup batches x passesPerBatch]
    for (int d = 0; d < nDev; d++) {
        if (ame_create(&ctxs[d], o.deviceIndex + d, W, H, kGroup * slotsPerBatch, kGroup * passesPerBatch) != AME_OK) {
            fprintf(stderr, "ame_create(device %d) failed: %s\n", o.deviceIndex + d, ame_last_error());
            return 1;
        }
        bool ok = true;
        for (int s = 0; ok && s < kGroup * slotsPerBatch; s++)  // the device-side copies of every slot, in the role it will have
            ok = ame_upload_plane_ex(ctxs[d], s, s % slotsPerBatch < B ? orig : recon, s % slotsPerBatch < B ? AME_ROLE_CURRENT : AME_ROLE_REFERENCE) == AME_OK;
        ok = ok && ame_sync(ctxs[d]) == AME_OK;
        pool[d].resize((size_t)2 * kGroup * passesPerBatch);
        for (PassOut &p : pool[d]) ok = ok && alloc_pass(p, ctxs[d]);
        double dummy[4];
        ok = ok && ame_exec_ns(ctxs[d], dummy, 1) == AME_OK;
        if (!ok) {
            fprintf(stderr, "allocation on device %d failed: %s\n", o.deviceIndex + d, ame_last_error());
            return 1;
        }
    }
    print_timestamp("FINISH ALLOCATE MEMORY");

    const auto lists = reference_lists(N);
    const int nBatches = (N + B - 1) / B;
    std::vector<Batch> batches(nBatches);
    for (int b = 0; b < nBatches; b++) {
        batches[b].firstFrame = b * B;
        batches[b].nFrames = std::min(B, N - b * B);
        for (int f = batches[b].firstFrame; f < batches[b].firstFrame + batches[b].nFrames; f++)
            for (int r = 0; r < (int)lists[f].size(); r++) batches[b].ids.push_back({f + 1, r});
    }

    std::mutex mu;
    std::condition_variable cv;
    int written = 0;  // batches consumed by the writer
    bool failed = false;

    print_timestamp("START GPU KERNEL");
    const double t0 = now_s();

    auto worker = [&](int d) {
        ame_ctx *ctx = ctxs[d];
        // batches of this GPU: d, d + nDev, ...; group gi = kGroup consecutive ones of them, pinned set gi & 1
        for (int gi = 0;; gi++) {
            std::vector<int> mine;
            for (int j = 0; j < kGroup; j++) {
                const int b = d + nDev * (gi * kGroup + j);
                if (b < nBatches) mine.push_back(b);
            }
            if (mine.empty()) return;
            {   // the pinned set was last used by group gi - 2: the writer must be through with it
                const int lastOld = d + nDev * ((gi - 2) * kGroup + kGroup - 1);
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || gi < 2 || written > std::min(lastOld, nBatches - 1); });
                if (failed) return;
            }
            bool ok = true;
            size_t groupPasses = 0;
            for (size_t j = 0; ok && j < mine.size(); j++) {
                Batch &bt = batches[mine[j]];
                bt.device = o.deviceIndex + d;
                bt.passes.resize(bt.ids.size());
                const size_t poolBase = ((size_t)(gi & 1) * kGroup + j) * passesPerBatch;
                for (size_t k = 0; k < bt.ids.size(); k++) bt.passes[k] = pool[d][poolBase + k];
                // planes: slot i < B of the batch's slot set holds current frame firstFrame+i; reference POCs get the following slots
                const int slot0 = (int)j * slotsPerBatch;
                std::map<int, int> refSlot;
                for (int i = 0; ok && i < bt.nFrames; i++) {
                    const int f = bt.firstFrame + i;
                    ok = ok && ame_upload_plane_ex(ctx, slot0 + i, orig + plane * f, AME_ROLE_CURRENT) == AME_OK;
                    for (int rp : lists[f]) {
                        if (refSlot.count(rp)) continue;
                        const int s = slot0 + B + (int)refSlot.size();
                        refSlot[rp] = s;
                        ok = ok && ame_upload_plane_ex(ctx, s, recon + plane * rp, AME_ROLE_REFERENCE) == AME_OK;
                    }
                }
                for (size_t k = 0; ok && k < bt.ids.size(); k++) {
                    const int poc = bt.ids[k].first, r = bt.ids[k].second, f = poc - 1;
                    ok = ok && ame_search(ctx, slot0 + f - bt.firstFrame, refSlot[lists[f][r]], lambda_for(o.qp, poc), o.extraGradIter, &bt.passes[k].res) == AME_OK;
                }
                ok = ok && ame_flush(ctx) == AME_OK;
                groupPasses += bt.ids.size();
            }
            ok = ok && ame_sync(ctx) == AME_OK;
            double ns[4] = {0, 0, 0, 0};
            ok = ok && ame_exec_ns(ctx, ns, 1) == AME_OK;
            const double tDone = now_s();
            std::lock_guard<std::mutex> lk(mu);
            if (!ok) {
                fprintf(stderr, "GPU %d: %s\n", o.deviceIndex + d, ame_last_error());
                failed = true;
            }
            // the device times of the group go to its batches by their share of the passes, laid out back to back before tDone
            double tail = 0;
            for (size_t j = mine.size(); j-- > 0;) {
                Batch &bt = batches[mine[j]];
                const double share = groupPasses ? (double)bt.ids.size() / (double)groupPasses : 0.0;
                double batchNs = 0;
                for (int k = 0; k < 4; k++) { bt.execNs[k] = ns[k] * share; batchNs += bt.execNs[k]; }
                bt.tDone = tDone - tail;
                tail += batchNs * 1e-9;
                bt.done = true;
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> threads;
    for (int d = 0; d < nDev; d++) threads.emplace_back(worker, d);

    // ---- writer: consumes batches in POC order (main.cpp:746-748, 980-1003) ----
    LogWriter log(o.cpmvLogFile, W, H);
    double execNs[4] = {0, 0, 0, 0}, logSeconds = 0;
    std::vector<double> devNs(nDev, 0.0);
    size_t logRows = 0;
    static const char *kExecName[4] = {"FULL 2 CPs", "FULL 3 CPs", "HALF 2 CPs", "HALF 3 CPs"};
    for (int b = 0; b < nBatches; b++) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return failed || batches[b].done; });
            if (failed) break;
        }
        Batch &bt = batches[b];
        double batchNs = 0;
        for (int k = 0; k < 4; k++) { execNs[k] += bt.execNs[k]; batchNs += bt.execNs[k]; }
        devNs[bt.device - o.deviceIndex] += batchNs;
        // The searches of a batch run fused (all passes and all four prediction types in one launch sequence), so the
        // reference's per-launch markers (main.cpp:764-959) are laid out inside the measured interval of the batch:
        // it ended at tDone and lasted batchNs; every pass gets an equal share, split by the batch's per-type times.
        const double t0b = bt.tDone - batchNs * 1e-9, perPass = batchNs * 1e-9 / (double)std::max<size_t>(1, bt.ids.size());
        for (size_t k = 0; k < bt.ids.size(); k++) {
            printf("POC   %d  RefIdx  %d  -> lambda %f\n", bt.ids[k].first, bt.ids[k].second, lambda_for(o.qp, bt.ids[k].first));
            double t = t0b + perPass * (double)k;
            for (int pr = 0; pr < 4; pr++) {
                char name[48];
                snprintf(name, sizeof name, "START EXEC %s", kExecName[pr]);
                print_timestamp_at(name, t);
                t += batchNs > 0 ? perPass * bt.execNs[pr] / batchNs : 0.0;
                snprintf(name, sizeof name, "FINISH EXEC %s", kExecName[pr]);
                print_timestamp_at(name, t);
            }
            const double tw = now_s();
            logRows += log.write_pass(bt.ids[k].first, bt.ids[k].second, bt.passes[k].res);
            logSeconds += now_s() - tw;
        }
        std::lock_guard<std::mutex> lk(mu);
        written = b + 1;
        cv.notify_all();
    }
    for (auto &t : threads) t.join();
    log.close();
    print_timestamp("FINISH GPU KERNEL");
    const double overall = now_s() - t0;

    // reportTimingResults (main_aux_functions.h:1416-1446), same keys.  The per-type figures are device time of the 2-CP /
    // 3-CP searches split between aligned and half-aligned CUs by evaluated 4x4 blocks (ame_exec_ns); lines after OVERALL
    // are extensions (per-GPU device time, host-side ingest and log throughput).
    printf("=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=\n");
    printf("TIMING RESULTS (nanoseconds)\n");
    printf("FULL_2CP_EXEC,%f\n", execNs[0]);
    printf("FULL_3CP_EXEC,%f\n", execNs[1]);
    printf("HALF_2CP_EXEC,%f\n", execNs[2]);
    printf("HALF_3CP_EXEC,%f\n", execNs[3]);
    printf("TOTAL_EXEC_TIME(%dx),%f\n", N, execNs[0] + execNs[1] + execNs[2] + execNs[3]);
    printf("OVERALL(%dx),%f\n", N, overall);
    for (int d = 0; d < nDev; d++) printf("GPU%d_EXEC,%f\n", o.deviceIndex + d, devNs[d]);
    printf("CSV_INGEST,%.1f MB/s,%.0f samples/s\n", readSeconds > 0 ? csvBytes / readSeconds / 1e6 : 0.0, readSeconds > 0 ? 2.0 * plane * N / readSeconds : 0.0);
    if (log.enabled()) printf("LOG_WRITE,%.0f rows/s\n", logSeconds > 0 ? logRows / logSeconds : 0.0);
    printf("=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=-=\n\n");

    for (auto &v : pool)
        for (PassOut &p : v) ame_free_host(p.block);
    for (ame_ctx *c : ctxs) ame_destroy(c);
    ame_free_host(orig);
    ame_free_host(recon);
    print_timestamp("FINISH HOST");
    return failed ? 1 : 0;
}

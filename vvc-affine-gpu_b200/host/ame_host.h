// Host side of the drop-in CLI: the reference's user-visible surface (flags, CSV frame ingest, lambda /
// reference-list schedule, per-CU decision logs, stdout markers) re-implemented over the C ABI of
// include/affine_me.h.  Mirrors /root/reference/main.cpp and main_aux_functions.h; each function cites
// the lines it replaces.
#pragma once
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "affine_me.h"

namespace host {

// ---- cli.cpp: flags of main.cpp:58-69 and the validation echo of main_aux_functions.h:77-145
struct Options {
    int deviceIndex = 0;        // --DeviceIndex
    int qp = -1;                // -q / --QP
    int nFrames = -1;           // -f / --FramesToBeEncoded
    int extraGradIter = 0;      // --ExtraGradientIter
    std::string resolution;     // -s / --Resolution  WxH
    std::string origFile;       // -o / --OriginalFrames
    std::string refFile;        // -r / --ReferenceFrames
    std::string cpmvLogFile;    // -l / --CpmvLogFile (default "": no files)
    int numDevices = 1;         // --NumDevices (extension: shard frames over GPUs DeviceIndex..+N-1)
    int batchFrames = 8;        // --BatchFrames (extension: frames queued per launch)
    bool rawFrames = false;     // --RawFrames (extension: inputs are raw uint16 planes, not CSV text)
    bool deviceIndexSet = false, logSet = false, qpSet = false, framesSet = false, extraSet = false, resSet = false,
         origSet = false, refSet = false, help = false;
};
// Returns 0 to continue, otherwise the process exit code + 1000 (so that 0 can be an exit code).
int parse_options(int argc, char **argv, Options &o);
int check_report_parameters(const Options &o);  // prints the reference's echo, returns the error count
void print_help();

// ---- schedule.cpp
int compute_delta_qp(int inputQp, int poc);             // main_aux_functions.h:1482-1497
float lambda_for(int inputQp, int poc);                 // main.cpp:585 + constants.h:94-103
// Reference POC list per frame (newest first), frames poc = 1..n: main.cpp:584, 591-707.
std::vector<std::vector<int>> reference_lists(int nFrames);
void print_reference_plan(int nFrames, int inputQp);    // testReferences, main_aux_functions.h:1499-1545

// ---- csv_ingest.cpp: main.cpp:303-328.  Parses nFrames*H lines of W comma-separated samples into dst
// (nFrames*W*H uint16, e.g. pinned memory).  Returns 0 or -1 (message in err).
int read_csv_frames(const std::string &path, int nFrames, int W, int H, uint16_t *dst, int threads, std::string &err);
// Same destination from a file of raw little-endian uint16 samples (nFrames * H * W of them; --RawFrames).
int read_raw_frames(const std::string &path, int nFrames, int W, int H, uint16_t *dst, std::string &err);

// ---- log_writer.cpp: reportAffineResultsMaster_new, main_aux_functions.h:387-525
class LogWriter {
public:
    LogWriter(const std::string &prefix, int W, int H);
    ~LogWriter();
    bool enabled() const { return !prefix_.empty(); }
    // Appends the rows of one (poc, ref) pass for all four prediction types, in the reference's order.
    // Returns the number of rows written.
    size_t write_pass(int poc, int ref, const ame_result &res);
    void close();
private:
    struct File { FILE *f = nullptr; std::string name; };
    File &file_for(int pred, int w, int h);
    std::string prefix_;
    int W_, H_, nCtus_, ctuCols_;
    std::vector<File> files_[4];
    std::vector<char> buf_;
};

void print_timestamp(const char *prefix);  // main_aux_functions.h:59-68
void print_timestamp_at(const char *prefix, double epochSeconds);  // same line for a given instant

}  // namespace host

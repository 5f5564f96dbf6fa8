#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>

#include <iostream>

#include "ame_host.h"

namespace host {

// Same option names and short forms as the boost::program_options table of main.cpp:58-69.
struct OptSpec { const char *longName; char shortName; bool takesArg; const char *help; };
static const OptSpec kOpts[] = {
    {"help", 'h', false, "produce help message"},
    {"DeviceIndex", 0, true, "Index of the GPU device (CUDA ordinal)"},
    {"QP", 'q', true, "Quantization parameter"},
    {"FramesToBeEncoded", 'f', true, "Number of frames to be processed"},
    {"ExtraGradientIter", 0, true, "Number of extra iterations during Gradient-based Affine ME"},
    {"Resolution", 's', true, "Resolution of the video, in the format 1920x1080"},
    {"OriginalFrames", 'o', true, "Input file for original frames samples"},
    {"ReferenceFrames", 'r', true, "Input file for reference frames samples"},
    {"CpmvLogFile", 'l', true, "Output files preffix with produced CPMVs"},
    {"NumDevices", 0, true, "(extension) shard frames over this many GPUs starting at DeviceIndex"},
    {"BatchFrames", 0, true, "(extension) frames queued per kernel launch"},
    {"RawFrames", 0, false, "(extension) the two input files hold raw little-endian 16-bit samples (frames stacked) instead of CSV text"},
};

void print_help() {
    std::cout << "Allowed options:\n";
    for (const OptSpec &o : kOpts) {
        std::string left = "  ";
        if (o.shortName) left += std::string("-") + o.shortName + " [ --" + o.longName + " ]";
        else left += std::string("--") + o.longName;
        if (o.takesArg) left += " arg";
        std::cout << left << "  " << o.help << "\n";
    }
    std::cout << "\n";
}

static bool to_int(const std::string &s, int &v) {
    char *end = nullptr;
    const long x = strtol(s.c_str(), &end, 10);
    if (end == s.c_str() || *end != 0) return false;
    v = (int)x;
    return true;
}

int parse_options(int argc, char **argv, Options &o) {
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        const OptSpec *spec = nullptr;
        std::string val;
        bool haveVal = false;
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            std::string n = a.substr(2);
            const size_t eq = n.find('=');
            if (eq != std::string::npos) { val = n.substr(eq + 1); n = n.substr(0, eq); haveVal = true; }
            for (const OptSpec &s : kOpts) if (n == s.longName) spec = &s;
        } else if (a.size() >= 2 && a[0] == '-') {
            for (const OptSpec &s : kOpts) if (s.shortName && a[1] == s.shortName) spec = &s;
            if (a.size() > 2) { val = a.substr(2); haveVal = true; }
        }
        if (!spec) { std::cerr << "unrecognised option '" << a << "'\n"; return 1000 + 1; }
        if (spec->takesArg && !haveVal) {
            if (i + 1 >= argc) { std::cerr << "the required argument for option '" << a << "' is missing\n"; return 1000 + 1; }
            val = argv[++i];
        }
        const std::string n = spec->longName;
        bool ok = true;
        if (n == "help") o.help = true;
        else if (n == "DeviceIndex") { ok = to_int(val, o.deviceIndex); o.deviceIndexSet = true; }
        else if (n == "QP") { ok = to_int(val, o.qp); o.qpSet = true; }
        else if (n == "FramesToBeEncoded") { ok = to_int(val, o.nFrames); o.framesSet = true; }
        else if (n == "ExtraGradientIter") { ok = to_int(val, o.extraGradIter); o.extraSet = true; }
        else if (n == "Resolution") { o.resolution = val; o.resSet = true; }
        else if (n == "OriginalFrames") { o.origFile = val; o.origSet = true; }
        else if (n == "ReferenceFrames") { o.refFile = val; o.refSet = true; }
        else if (n == "CpmvLogFile") { o.cpmvLogFile = val; o.logSet = true; }
        else if (n == "NumDevices") ok = to_int(val, o.numDevices);
        else if (n == "BatchFrames") ok = to_int(val, o.batchFrames);
        else if (n == "RawFrames") o.rawFrames = true;
        if (!ok) { std::cerr << "the argument ('" << val << "') for option '--" << n << "' is invalid\n"; return 1000 + 1; }
    }
    return 0;
}

// Same lines as checkReportParameters (main_aux_functions.h:77-145).
int check_report_parameters(const Options &o) {
    using std::cout;
    using std::endl;
    int errors = 0;
    cout << "-=-= INPUT PARAMETERS =-=-" << endl;
    if (!o.deviceIndexSet) cout << "  Device index not set. Using standard value of " << o.deviceIndex << "." << endl;
    else cout << "  Device Index=" << o.deviceIndex << endl;
    if (!o.logSet) cout << "  CPMVs log file not set. The output will not be written to any file." << endl;
    else cout << "  CpmvLogFile=" << o.cpmvLogFile << endl;
    if (o.qpSet) cout << "  QP=" << o.qp << endl;
    else { cout << "  [!] ERROR: QP not set." << endl; errors++; }
    if (o.framesSet) cout << "  FramesToBeEncoded=" << o.nFrames << endl;
    else { cout << "  [!] ERROR: FramesToBeEncoded not set." << endl; errors++; }
    if (!o.extraSet) cout << "  ExtraGradientIter not specified. Using zero extra gradients (i.e., 5 iterations for 2 CPs and 4 iterations for 3 CPs)." << endl;
    else cout << "  ExtraGradientIter=" << o.extraGradIter << ". Using a total of " << 5 + o.extraGradIter << " iterations for 2 CPs and " << 4 + o.extraGradIter << " iterations for 3 CPs." << endl;
    if (o.resSet) cout << "  Resolution=" << o.resolution << endl;
    else { cout << "  [!] ERROR: Resolution not set." << endl; errors++; }
    if (o.origSet) cout << "  InputOriginalFrame=" << o.origFile << endl;
    else { cout << "  [!] ERROR: Input original frames not set." << endl; errors++; }
    if (o.refSet) cout << "  InputReferenceFrame=" << o.refFile << endl;
    else { cout << "  [!] ERROR: Input reference frames not set." << endl; errors++; }
    return errors;
}

void print_timestamp_at(const char *prefix, double epochSeconds) {
    const time_t sec = (time_t)epochSeconds;
    int ms = (int)((epochSeconds - (double)sec) * 1000.0);
    if (ms < 0) ms = 0;
    if (ms > 999) ms = 999;
    struct tm tmv;
    localtime_r(&sec, &tmv);
    printf("%s @ %02d:%02d:%02d.%03d\n", prefix, tmv.tm_hour, tmv.tm_min, tmv.tm_sec, ms);
}

void print_timestamp(const char *prefix) {
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    print_timestamp_at(prefix, (double)tv.tv_sec + (double)tv.tv_usec * 1e-6);
}

}  // namespace host

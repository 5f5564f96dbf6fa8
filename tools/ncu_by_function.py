"""Joins an ncu report's per-SASS-instruction counters with nvdisasm line info of the same kernel build and
prints dynamic instruction counts / stall samples per source function.
usage: ncu_by_function.py report.ncu-rep [kernel# in the report] [kernel function name, default ame_iter_small]
The name selects the .text section by substring: give the mangled instance of a template (ame_update_kernelILi3E,
ame_iter_bigILb1E), or the listing holds the instructions of all its instances."""
import bisect, collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kname = sys.argv[3] if len(sys.argv) > 3 else "ame_iter_small"
src_path = os.environ.get("AME_SRC") or os.path.join(ROOT, "vvc-affine-gpu_b200", "csrc", "ame_kernels.cu")
tmp = tempfile.mkdtemp()
obj = os.path.join(tmp, "k.o")
subprocess.check_call(["nvcc", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-I", os.path.join(ROOT, "vvc-affine-gpu_b200", "csrc"), "-c", src_path, "-o", obj])
subprocess.check_call(["cuobjdump", "-xelf", "all", obj], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# instruction list of the search kernel with source lines
lines, cur, infn = [], None, False
for ln in txt.split("\n"):
    if ln.startswith(".text."):
        infn = kname in ln
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        cur = int(m.group(2)) if os.path.basename(m.group(1)) == os.path.basename(src_path) else -1
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and infn:
        lines.append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
secs, curs = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        curs = []
        secs.append(curs)
    elif curs is not None:
        curs.append(row)
sec = secs[which]
hdr = sec[0]
iex, ismp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
rows = sec[1:]
print("kernel section", which, "sass rows", len(rows), "nvdisasm rows", len(lines))
assert len(rows) == len(lines), "SASS listing and nvdisasm disagree: rebuild from the profiled source"
src = open(src_path).read().split("\n")
marks = [(i + 1, l) for i, l in enumerate(src) if re.match(r"^(__device__|__global__)", l)]
starts = [m[0] for m in marks]
agg, smp, stat = collections.Counter(), collections.Counter(), collections.Counter()
STALLS = ["stall_wait", "stall_long_sb", "stall_short_sb", "stall_no_inst", "stall_branch_resolving", "stall_math", "stall_not_selected", "stall_barrier", "stall_selected"]
sidx = [hdr.index(c) if c in hdr else -1 for c in STALLS]
stl = collections.defaultdict(lambda: [0] * len(STALLS))
ops = collections.Counter()
for (line, ins), r in zip(lines, rows):
    n, s = int(r[iex] or 0), int(r[ismp] or 0)
    if line is None or line < 0:
        key = "<intrinsics: dp2a/shf/prmt/shfl/ldg>"
    else:
        j = bisect.bisect_right(starts, line) - 1
        key = re.sub(r"__device__ |__forceinline__ |__noinline__ |__global__ ", "", marks[j][1])[:70] if j >= 0 else "?"
    agg[key] += n
    smp[key] += s
    for q, ix in enumerate(sidx):
        if ix >= 0 and r[ix]:
            stl[key][q] += int(r[ix])
    stat[key] += 1
    op = ins.split()[0] if not ins.startswith("@") else ins.split()[1]
    ops[op.split(".")[0]] += n
tot, tots = sum(agg.values()), sum(smp.values())
print("%-72s %8s %7s %7s %6s" % ("function", "Minst", "inst%", "stall%", "static"))
for k, v in agg.most_common():
    print("%-72s %8.1f %6.1f%% %6.1f%% %6d" % (k, v / 1e6, 100.0 * v / tot, 100.0 * smp[k] / max(tots, 1), stat[k]))
print("total %.1f M warp instructions" % (tot / 1e6))
print("\nstall samples by function (%% of all samples): " + " ".join(c.replace("stall_", "") for c in STALLS))
for k, v in sorted(smp.items(), key=lambda kv: -kv[1])[:14]:
    print("%-60s %5.1f%% | " % (k[:60], 100.0 * v / max(tots, 1)) + " ".join("%5.1f" % (100.0 * x / max(tots, 1)) for x in stl[k]))
print("by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in ops.most_common(24)))
if len(sys.argv) > 3:
    pat = sys.argv[3]
    print("\n--- instructions of functions matching", pat)
    for (line, ins), r in zip(lines, rows):
        if line is None or line < 0:
            continue
        j = bisect.bisect_right(starts, line) - 1
        if j >= 0 and pat in marks[j][1]:
            print("%5d %10s %6s  %s" % (line, r[iex], r[ismp], ins[:90]))

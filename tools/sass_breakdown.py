"""SASS instruction count per source function of the search kernel (uses nvdisasm -g line info)."""
import bisect, collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_path = os.path.join(ROOT, "vvc-affine-gpu_b200", "csrc", "ame_kernels.cu")
tmp = tempfile.mkdtemp()
obj = os.path.join(tmp, "k.o")
subprocess.check_call(["nvcc", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-c", src_path, "-o", obj])
subprocess.check_call(["cuobjdump", "-xelf", "all", obj], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cnt, cur = collections.Counter(), None
for ln in txt.split("\n"):
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith("ame_kernels.cu") else -1
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln) and cur:
        cnt[cur] += 1
src = open(src_path).read().split("\n")
marks = [(i + 1, l) for i, l in enumerate(src) if re.match(r"^(__device__|__global__)", l)]
starts = [m[0] for m in marks]
agg = collections.Counter()
for line, c in cnt.items():
    if line < 0:
        agg[(0, "<cuda intrinsics headers>")] += c
        continue
    j = bisect.bisect_right(starts, line) - 1
    agg[(marks[j][0], marks[j][1][:100]) if j >= 0 else (0, "?")] += c
for (l, n), c in sorted(agg.items()):
    print("%5d  line %4d  %s" % (c, l, n))
print("total", sum(cnt.values()), "instructions =", sum(cnt.values()) * 16 // 1024, "KB")

"""Opcode histogram of every kernel in libaffine_me.so (cuobjdump -sass): instruction count, top opcodes, local-memory
instructions (LDL / STL: spills) and the TMA / mbarrier instructions.  usage: python tools/sass_histogram.py [lib]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vvc-affine-gpu_b200", "libaffine_me.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, per, lines = None, collections.OrderedDict(), collections.defaultdict(list)
for ln in txt.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        per[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", ln)
    if m and kern:
        per[kern][m.group(1)] += 1
        if m.group(1) in ("UTMALDG", "SYNCS", "LDL", "STL", "UBLKCP"):
            lines[kern].append(re.sub(r"\s+/\*[0-9a-fx]+\*/\s*$", "", ln.strip()))
print("cuobjdump -sass %s" % os.path.relpath(lib, ROOT))
for k, c in per.items():
    tot = sum(c.values())
    print("\n%s: %d instructions; LDL %d, STL %d (local memory); UTMALDG %d, SYNCS %d" % (k, tot, c["LDL"], c["STL"], c["UTMALDG"], c["SYNCS"]))
    print("  " + ", ".join("%s %d" % kv for kv in c.most_common(16)))
    for l in lines[k][:12]:
        print("    " + l)

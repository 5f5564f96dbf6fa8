"""Small workload for compute-sanitizer: border-heavy golden cases + one 832x480 pass through the C ABI."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from conftest import load_pkg
import oracle_binding as ob
pkg = load_pkg()
bad = 0
for name in ("bigmotion_416x240", "stress_noise_416x240", "affine_832x480_f1_q27"):
    d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    n, H, W = d["orig"].shape
    ctx = pkg.AffineME(W, H)
    costs, cp = ctx.ref_pass(d["recon"][0], d["orig"][0], ob.lambda_for(int(d["qp"]), 1), int(d["extra_iter"]))
    for p in range(4):
        bad += int((costs[p] != d["cost_0_%d" % p]).sum())
    ctx.close()
print("mismatching costs:", bad)
import ctypes
if hasattr(pkg.lib(), "ame_debug_stats"):
    st = (ctypes.c_ulonglong * 24)()
    pkg.lib().ame_debug_stats(st, 0)
    print("out-of-window accesses counted by an AME_STATS build:", st[20], "(searches counted: %d)" % sum(st[0:16]))

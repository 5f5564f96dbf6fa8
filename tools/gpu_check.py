"""Quick GPU parity check of the CUDA path against the golden fixtures (development helper)."""
import glob
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from conftest import load_pkg  # noqa: E402
import ame_logs  # noqa: E402
import oracle_binding as ob  # noqa: E402

pkg = load_pkg()
FLD = ("LTx", "LTy", "RTx", "RTy", "LBx", "LBy")
total_bad = 0
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
    d = np.load(path)
    orig, recon, qp, extra = d["orig"], d["recon"], int(d["qp"]), int(d["extra_iter"])
    n, H, W = orig.shape
    ctx = pkg.AffineME(W, H)
    
    lists = ob.ref_lists(n)
    bad = tot = 0
    t0 = time.time()
    for k, (poc, r) in enumerate(ame_logs.pass_list(n)):
        lam = ob.lambda_for(qp, poc)
        costs, cp = ctx.ref_pass(recon[lists[poc - 1][r]], orig[poc - 1], lam, extra)
        for p in range(4):
            m = d["cost_%d_%d" % (k, p)] != costs[p]
            for f in FLD:
                m |= d["cpmv_%d_%d" % (k, p)][f] != cp[p][f]
            if m.any() and bad < 5:
                i = int(np.flatnonzero(m)[0])
                print("  first mismatch pred %d idx %d: golden %s %s  got %s %s" % (
                    p, i, d["cost_%d_%d" % (k, p)][i], d["cpmv_%d_%d" % (k, p)][i], costs[p][i], cp[p][i]))
            bad += int(m.sum())
            tot += m.size
    ms, nl = ctx.last_kernel_ms()
    print("%s: %d/%d mismatching CUs (%.2fs, last pass kernels %.3f ms in %d launches)" % (
        os.path.basename(path), bad, tot, time.time() - t0, ms, nl), flush=True)
    total_bad += bad
    ctx.close()
print("TOTAL MISMATCHES", total_bad)
sys.exit(1 if total_bad else 0)

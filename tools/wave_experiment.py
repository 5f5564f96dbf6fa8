"""How the size of a launch sequence changes the per-pass time of the bench workload (1080p, 64 frames, 250
reference passes, planes resident): the same 250 searches are launched as sequences of K passes each
(ame_flush per K), K from --ks; also prints the device time of every sequence for one K (--detail) to show
how the per-pass time moves along the sequence (late frames search longer-term references).  Not a benchmark."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
from conftest import load_pkg  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ks", default="250,125,64,32,16,8,4")
ap.add_argument("--detail", type=int, default=16)
ap.add_argument("--qp", type=int, default=32)
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
pkg = load_pkg()
orig, recon = bench.make_sequences(bench.W, bench.H, a.frames, (a.qp,))
N = a.frames
lists = bench.ref_lists(N)
passes = [(poc, r, lists[poc - 1][r]) for poc in range(1, N + 1) for r in range(len(lists[poc - 1]))]
ctx = pkg.AffineME(bench.W, bench.H, num_slots=2 * N, max_in_flight=len(passes))
for f in range(N):
    ctx.upload(f, orig[f], pkg.ROLE_CURRENT)
    ctx.upload(N + f, recon[a.qp][f], pkg.ROLE_REFERENCE)
ctx.sync()


def run(order, K, detail=False):
    per = []
    ctx.timer_start()
    t_prev = 0.0
    for k0 in range(0, len(order), K):
        for k in order[k0:k0 + K]:
            poc, r, rp = passes[k]
            ctx.search_device(poc - 1, N + rp, bench.lambda_for(a.qp, poc), k)
        ctx.flush()
        if detail:
            ms = ctx.last_kernel_ms()[0]
            per.append(ms)
    total = ctx.timer_stop()
    ctx.sync()
    return total, per


by_poc = list(range(len(passes)))
by_ref = sorted(by_poc, key=lambda k: (passes[k][2], passes[k][0]))
for name, order in (("(poc, ref) order", by_poc), ("grouped by reference plane", by_ref)):
    for K in [int(x) for x in a.ks.split(",")]:
        best = min(run(order, K)[0] for _ in range(a.reps))
        print("%-28s K=%3d passes per sequence: %8.2f ms -> %7.1f passes/s" % (name, K, best, 1000.0 * len(passes) / best), flush=True)
if a.detail:
    total, per = run(by_poc, a.detail, True)
    print("per-sequence device ms at K=%d (poc order):" % a.detail, " ".join("%.2f" % x for x in per))
    print("  -> ms per pass:", " ".join("%.3f" % (x / min(a.detail, len(passes) - i * a.detail)) for i, x in enumerate(per)))
ctx.close()

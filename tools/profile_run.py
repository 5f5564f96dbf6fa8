"""Small fixed workload for ncu: N 1080p frames (default 4 -> 10 reference passes) in one launch sequence,
repeated --reps times.  Prints device ms per rep.  Not a benchmark."""
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
import synth_frames as sf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--size", default="1920x1080")
ap.add_argument("--early-exit", type=int, default=1)
a = ap.parse_args()
W, H = [int(x) for x in a.size.split("x")]
pkg = load_pkg()
orig, recon = sf.sequences(a.frames, W, H, 32)
lists = bench.ref_lists(a.frames)
passes = [(poc, r, lists[poc - 1][r]) for poc in range(1, a.frames + 1) for r in range(len(lists[poc - 1]))]
ctx = pkg.AffineME(W, H, num_slots=2 * a.frames, max_in_flight=len(passes))
ctx.set_option(pkg.OPT_EARLY_EXIT, a.early_exit)
for f in range(a.frames):
    ctx.upload(f, orig[f])
    ctx.upload(a.frames + f, recon[f])
ctx.sync()
for rep in range(a.reps):
    for k, (poc, r, rp) in enumerate(passes):
        ctx.search_device(poc - 1, a.frames + rp, bench.lambda_for(32, poc), k)
    ctx.timer_start()
    ctx.flush()
    ms = ctx.timer_stop()
    ctx.sync()
    print("rep %d: %d passes in %.3f ms -> %.1f passes/s" % (rep, len(passes), ms, 1000.0 * len(passes) / ms), flush=True)
import ctypes
st = (ctypes.c_ulonglong * 24)()
if hasattr(pkg.lib(), "ame_debug_stats"):
    pkg.lib().ame_debug_stats(st, 0)
    v = list(st)
    if sum(v):
        tot = a.reps
        print("2-CP searches by states evaluated:", [x // tot for x in v[0:8]])
        print("3-CP searches by states evaluated:", [x // tot for x in v[8:16]])
        print("exit: fixed point %d, 2-cycle %d, 3-cycle %d, iteration limit %d" % tuple(x // tot for x in v[16:20]))
ctx.close()

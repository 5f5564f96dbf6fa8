"""Runs the drop-in CLI on a golden input with --NumDevices 1 and 2 (needs two GPUs) and checks that the 40 log
files are byte-identical; also prints the CLI's own timing block on a 1080p input."""
import filecmp
import glob
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth_frames as sf  # noqa: E402

CLI = os.path.join(ROOT, "vvc-affine-gpu_b200", "bin", "affine_b200")
tmp = tempfile.mkdtemp(prefix="ame_cli_")
d = np.load(os.path.join(ROOT, "tests", "golden", "affine_416x240_f3_q32.npz"))
n, H, W = d["orig"].shape
sf.write_csv(os.path.join(tmp, "o.csv"), d["orig"])
sf.write_csv(os.path.join(tmp, "r.csv"), d["recon"])
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for k in (1, ndev):
    r = subprocess.run([CLI, "-f", str(n), "-s", "%dx%d" % (W, H), "-q", "32", "-o", os.path.join(tmp, "o.csv"), "-r", os.path.join(tmp, "r.csv"),
                        "-l", os.path.join(tmp, "log%d" % k), "--NumDevices", str(k), "--BatchFrames", "1"], capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:])
        sys.exit("CLI failed with --NumDevices %d" % k)
a = sorted(glob.glob(os.path.join(tmp, "log1_*.csv")))
bad = [f for f in a if not filecmp.cmp(f, f.replace("log1_", "log%d_" % ndev), shallow=False)]
print("%d log files, %d differ between --NumDevices 1 and %d" % (len(a), len(bad), ndev))
# 1080p timing through the CLI (8 frames, no logs)
o, r8 = sf.sequences(8, 1920, 1080, 32)
sf.write_csv(os.path.join(tmp, "o8.csv"), o)
sf.write_csv(os.path.join(tmp, "r8.csv"), r8)
for k in (1, ndev):
    r = subprocess.run([CLI, "-f", "8", "-s", "1920x1080", "-q", "32", "-o", os.path.join(tmp, "o8.csv"), "-r", os.path.join(tmp, "r8.csv"),
                        "--NumDevices", str(k)], capture_output=True, text=True)
    lines = [l for l in r.stdout.splitlines() if "READ .csv" in l or ("_EXEC" in l) or "OVERALL" in l or "CSV_INGEST" in l]
    print("--NumDevices %d: rc=%d" % (k, r.returncode), " | ".join(lines))
    if r.returncode != 0 or sum(1 for l in lines if l.startswith("GPU") and "_EXEC," in l) != k:
        print(r.stdout[-1500:], r.stderr[-1500:])
        sys.exit("1080p CLI run with --NumDevices %d failed or lacks its per-GPU lines" % k)
sys.exit(1 if bad else 0)

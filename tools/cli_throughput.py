"""The drop-in CLI on the bench workload: 64 frames of 1080p (250 reference passes, QP 32) from raw 16-bit input files
(--RawFrames; the same frames as CSV text would be 1.2 GB), without log files, on 1 .. N GPUs (--NumDevices).
Prints the CLI's OVERALL wall time of its GPU section and the frames/s it amounts to.  usage: cli_throughput.py [N] [frames]"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402

CLI = os.path.join(ROOT, "vvc-affine-gpu_b200", "bin", "affine_b200")
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 1
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
orig, recon = bench.make_sequences(bench.W, bench.H, frames, (32,))
tmp = tempfile.mkdtemp(prefix="ame_cli_tp_")
orig.tofile(os.path.join(tmp, "o.raw"))
recon[32].tofile(os.path.join(tmp, "r.raw"))
n = 1
while n <= ndev:
    for rep in range(2):
        r = subprocess.run([CLI, "-f", str(frames), "-s", "%dx%d" % (bench.W, bench.H), "-q", "32", "-o", os.path.join(tmp, "o.raw"), "-r", os.path.join(tmp, "r.raw"),
                            "--RawFrames", "--NumDevices", str(n)], capture_output=True, text=True)
        if r.returncode != 0:
            sys.exit("CLI failed: " + r.stdout[-500:] + r.stderr[-500:])
        overall = float(re.search(r"^OVERALL\(\d+x\),([0-9.]+)", r.stdout, re.M).group(1))
        total = float(re.search(r"^TOTAL_EXEC_TIME\(\d+x\),([0-9.]+)", r.stdout, re.M).group(1))
    print("--NumDevices %d: %d frames, OVERALL %.3f s -> %.1f frames/s through the CLI (device time of the searches %.1f ms)" % (
        n, frames, overall, frames / overall, total / 1e6), flush=True)
    n *= 2

"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (oracle/_ref/affine_ref, built by `make -C oracle ref`) on a GPU box
through NVIDIA's OpenCL, on synthetic inputs from tools/synth_frames.py.

Run on the GPU box (needs an OpenCL runtime):
    python tools/make_golden.py --out gpurun_out/golden
It also runs the CPU oracle on the same inputs under every FP option
combination and prints the mismatch counts, which is how the oracle's default
options (fused back-substitution, double->int rule) were settled.

Each fixture <case>.npz holds: orig, recon (n,H,W uint16), qp, extra_iter, and
for every pass k (in reference order) cost_{k}_{pred}, cpmv_{k}_{pred}.
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ame_logs  # noqa: E402
import synth_frames as sf  # noqa: E402

REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def run_reference(orig, recon, qp, extra_iter=0, keep_stdout=None):
    n, H, W = orig.shape
    tmp = tempfile.mkdtemp(prefix="ame_ref_")
    try:
        sf.write_csv(os.path.join(tmp, "original_frames.csv"), orig)
        sf.write_csv(os.path.join(tmp, "reconstructed_frames.csv"), recon)
        cmd = [os.path.join(REF_DIR, "affine_ref"), "-f", str(n), "-s", "%dx%d" % (W, H), "-q", str(qp),
               "-o", os.path.join(tmp, "original_frames.csv"), "-r", os.path.join(tmp, "reconstructed_frames.csv"),
               "-l", os.path.join(tmp, "log")]
        if extra_iter:
            cmd += ["--ExtraGradientIter", str(extra_iter)]
        t = time.time()
        p = subprocess.run(cmd, cwd=REF_DIR, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        dt = time.time() - t
        if keep_stdout:
            with open(keep_stdout, "w") as f:
                f.write(p.stdout)
        if p.returncode != 0:
            raise RuntimeError("reference failed rc=%d\n%s" % (p.returncode, p.stdout[-3000:]))
        res = ame_logs.read_logs(os.path.join(tmp, "log"), W, H, n)
        return res, p.stdout, dt
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cases():
    out = []
    o, r = sf.sequences(3, 416, 240, 32)
    out.append(("affine_416x240_f3_q32", o, r, 32, 0))
    o, r = sf.sequences(1, 416, 240, 22)
    out.append(("affine_416x240_f1_q22", o, r, 22, 0))
    o, r = sf.sequences(1, 416, 240, 37, seed=sf.SEED + 5)
    out.append(("affine_416x240_f1_q37_x2", o, r, 37, 2))
    o, r = sf.sequences(1, 832, 480, 27)
    out.append(("affine_832x480_f1_q27", o, r, 27, 0))
    for name, cur, ref in sf.stress_frames(416, 240):
        out.append(("stress_%s_416x240" % name, cur[None], ref[None], 32, 0))
    # large motion near picture borders: strong zoom + shift
    o = sf.frame(40, 416, 240)[None]
    r = sf.frame(0, 416, 240)[None]
    out.append(("bigmotion_416x240", o, r, 32, 0))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    ap.add_argument("--compare-oracle", type=int, default=1)
    ap.add_argument("--time-1080p", type=int, default=0, help="also time the reference on N 1080p frames")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    if a.compare_oracle:
        import oracle_binding as ob
    summary = []
    for name, orig, recon, qp, extra in cases():
        res, stdout, dt = run_reference(orig, recon, qp, extra, keep_stdout=os.path.join(a.out, name + ".stdout.txt"))
        n = orig.shape[0]
        blob = {"orig": orig, "recon": recon, "qp": np.int32(qp), "extra_iter": np.int32(extra)}
        passes = ame_logs.pass_list(n)
        for k, p in enumerate(passes):
            for pred in range(4):
                blob["cost_%d_%d" % (k, pred)] = res[p][pred][0]
                blob["cpmv_%d_%d" % (k, pred)] = res[p][pred][1]
        np.savez_compressed(os.path.join(a.out, name + ".npz"), **blob)
        line = "%s: reference ran %.1fs, %d passes" % (name, dt, len(passes))
        if a.compare_oracle:
            lists = ob.ref_lists(n)
            for fused in (1, 0):
                for cvt in (1, 0):
                    bad = tot = 0
                    for (poc, r) in passes:
                        lam = ob.lambda_for(qp, poc)
                        refpoc = lists[poc - 1][r]
                        costs, cp = ob.ref_pass(recon[refpoc], orig[poc - 1], lam,
                                                ob.default_opts(extra_iter=extra, fused_backsub=fused, cvt_rule=cvt))
                        for pred in range(4):
                            gc, gm = res[(poc, r)][pred]
                            m = (gc != costs[pred])
                            for fld in ("LTx", "LTy", "RTx", "RTy", "LBx", "LBy"):
                                m |= gm[fld] != cp[pred][fld]
                            bad += int(m.sum())
                            tot += m.size
                    line += " | fused=%d cvt=%d: %d/%d mismatching CUs" % (fused, cvt, bad, tot)
        print(line, flush=True)
        summary.append(line)
    if a.time_1080p:
        o, r = sf.sequences(a.time_1080p, 1920, 1080, 32)
        res, stdout, dt = run_reference(o, r, 32, 0, keep_stdout=os.path.join(a.out, "timing_1080p.stdout.txt"))
        tail = [l for l in stdout.splitlines() if "EXEC" in l or "OVERALL" in l or "TIMING" in l]
        summary.append("1080p x%d frames: wall %.1fs\n" % (a.time_1080p, dt) + "\n".join(tail[-12:]))
        print(summary[-1], flush=True)
    with open(os.path.join(a.out, "SUMMARY.txt"), "w") as f:
        f.write("\n".join(summary) + "\n")


if __name__ == "__main__":
    main()

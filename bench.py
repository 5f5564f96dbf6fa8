#!/usr/bin/env python
"""bench.py -- throughput of the affine-ME hot path on synthetic 1080p 10-bit frames.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): 1920x1080, 64 frames with global zoom/rotation/
translation (tools/synth_frames.py); one STEP = the whole 64-frame sequence = 250 reference
passes (frame poc runs min(4, poc) passes, main.cpp:584,746) at QP = (22, 27, 32, 37)[step % 4].
`value` = 1080p frames/s with the reference's multi-reference semantics; ref-passes/s is reported
beside it.  N > 1 (torchrun, one rank per GPU): every rank runs the same amount of work on its
own GPU (weak scaling, no inter-GPU traffic), value = all ranks' frames / max-over-ranks time.

  value : planes resident in HBM before the timed region, results left in HBM;
          device time from CUDA events on the context's stream.
  e2e   : same step through the C ABI with HOST buffers: pinned planes uploaded and every
          search's costs/CPMVs copied back inside the timed region.
  roofline : SURVEY.md 8(d): the path is bound by the INT32 issue rate (algorithmic DRAM traffic
          is ~11 MB per pass); achieved = as-written op model x passes/s, peak = 148 SMs x 4 x 32
          lanes x the SM clock sampled during the run.
  cpu_baseline : the CPU oracle (a port of the reference's OpenCL kernels, oracle/ame_oracle.c)
          on the host cores, rank 0, on a bounded sample.

--impl reference times that CPU port alone (the reference's own OpenCL kernels cannot run on
the host: there is no CPU OpenCL runtime in the image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

W, H, N_FRAMES = 1920, 1080, 64
QPS = (22, 27, 32, 37)
WORKLOAD = "synthetic 1080p 10-bit, 64 frames with global zoom/rotation, QP sweep 22/27/32/37 (one QP per step)"
N_CTUS = 135
# S = sum over in-frame CUs of w*h; OPS = S * (11*40.2 + 5*45 + 4*71) int32 lane-ops per pass
S_1080P = 42585600
OPS_PER_PASS = S_1080P * (11 * 40.2 + 5 * (13 + 32) + 4 * (13 + 58))
ALGO_BYTES_PER_PASS = 2 * W * H * 2 + 2 * N_CTUS * (201 + 284) * (8 + 28)


# ----------------------------------------------------------------------------- host-side schedule
def ref_lists(n):
    """Reference POCs per frame, newest first (restates main.cpp:591-707 as a label simulation)."""
    refs, lt, out = [-1] * 4, [0] * 4, []
    for poc in range(1, n + 1):
        num = min(4, poc)
        if poc < 5:
            a = refs[0]
            refs[0] = poc - 1
            b = None
            if num > 1:
                b, refs[1] = refs[1], a
            if num > 2:
                a, refs[2] = refs[2], b
            if num > 3:
                refs[3] = a
            lt[3] = 1 if refs[3] % 8 == 0 else 0
        else:
            a = refs[0]
            refs[0] = poc - 1
            if lt[1] == 0 or (a % 8 == 0 and a != refs[0]):
                b, refs[1] = refs[1], a
                if lt[2] == 0 or (b % 8 == 0 and b != refs[1]):
                    a, refs[2] = refs[2], b
                    if lt[3] == 0 or (a % 8 == 0 and a != refs[3]):
                        refs[3] = a
            lt[3] = 1 if refs[3] % 8 == 0 else 0
            lt[2] = 1 if (refs[2] % 8 == 0 and lt[3]) else 0
            lt[1] = 1 if (refs[1] % 8 == 0 and lt[2]) else 0
        out.append(refs[:num])
    return out


_FULL_LAMBDAS = [0.0] * 11 + [2.769291, 3.108425, 3.489089, 3.916370, 4.395976, 4.934316, 5.538583, 6.216849, 6.978177,
                              7.832739, 8.791952, 9.868633, 11.077166, 12.433698, 13.956355, 15.665478, 17.583905,
                              19.737266, 22.154332, 24.867397, 27.912709, 31.330957, 35.167810, 39.474532, 44.308664,
                              49.734793, 55.825418, 62.661913, 70.335619, 78.949063, 88.617327, 99.469587, 111.650836,
                              125.323826, 140.671239, 157.898127, 177.234655, 198.939174, 223.301672, 250.647653,
                              281.342477, 315.796254, 354.469310, 397.878347, 446.603345, 501.295305, 562.684955,
                              631.592507, 708.938619]


def lambda_for(qp, poc):
    """main_aux_functions.h:1482-1497 + constants.h:94-103."""
    q = qp + (1, 5, 4, 5, 4, 5, 4, 5)[poc % 8]
    if poc % 8:
        q += int(np.floor(min(3.0, max(0.0, q * 0.259 + -6.5 + 0.5))))
    return float(np.float32(_FULL_LAMBDAS[q]))


def _gen_frame(t):
    import synth_frames as sf
    return sf.frame(t, W, H)


def make_sequences():
    """frames 0..64 (shared by all QPs) and the per-QP reconstructed (noisy) sets."""
    import multiprocessing as mp
    import synth_frames as sf
    with mp.get_context("fork").Pool(min(16, os.cpu_count() or 1)) as pool:
        frames = pool.map(_gen_frame, range(N_FRAMES + 1))
    frames = np.stack(frames)
    recon = {}
    for qp in QPS:
        a = {22: 1, 27: 2, 32: 3, 37: 5}[qp]
        rng = np.random.Generator(np.random.PCG64(sf.SEED + 1000 * qp))
        noise = rng.integers(-a, a + 1, size=(N_FRAMES, H, W), dtype=np.int16)
        recon[qp] = np.clip(frames[:-1].astype(np.int16) + noise, 0, 1023).astype(np.uint16)
    return frames[1:], recon


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU port (oracle)
def cpu_port_rate(orig, recon, qp, rows, threads=0):
    """Times the CPU oracle on the first `rows` CTU rows of one 1080p pass (poc 1, ref 0).
    Returns (frames/s extrapolated to a full pass, seconds, cores)."""
    import oracle_binding as ob
    hh = min(H, rows * 128)
    cur = np.ascontiguousarray(orig[0][:hh])
    ref = np.ascontiguousarray(recon[0][:hh])
    lam = ob.lambda_for(qp, 1)
    t = time.perf_counter()
    ob.ref_pass(ref, cur, lam, ob.default_opts(threads=threads))
    dt = time.perf_counter() - t
    frac = (rows * 15) / float(N_CTUS) if hh < H else 1.0
    passes_per_s = frac / dt
    cores = threads if threads > 0 else (os.cpu_count() or 1)
    return passes_per_s * N_FRAMES / 250.0, passes_per_s, dt, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import synth_frames as sf
    frames = [sf.frame(t, W, H) for t in (0, 1)]
    orig = np.stack(frames[1:])
    rng = np.random.Generator(np.random.PCG64(sf.SEED + 1000 * 32))
    recon = np.clip(np.stack(frames[:1]).astype(np.int16) + rng.integers(-3, 4, size=(1, H, W), dtype=np.int16), 0, 1023).astype(np.uint16)
    rows = 3
    for _ in range(args.warmup):
        cpu_port_rate(orig, recon, 32, rows)
    t0 = time.perf_counter()
    fps = []
    for _ in range(args.steps):
        f, p, dt, cores = cpu_port_rate(orig, recon, 32, rows)
        fps.append(f)
    total = time.perf_counter() - t0
    v = float(np.mean(fps))
    sample = "CTU rows 0-%d of one 1080p reference pass (poc 1, ref 0, QP 32) per step, extrapolated by CTU count" % (rows - 1)
    line = {"impl": "reference", "metric": "1080p frames/sec affine ME", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": N_FRAMES, "ref_passes_per_step": 250,
                       "note": "CPU arm: each step times a bounded sample of this workload (see cpu_baseline.sample)"},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ref_passes_per_s": v * 250.0 / N_FRAMES}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- the B200 arm
def run_b200(args, rank, world, local_rank):
    import torch
    from conftest import load_pkg
    pkg = load_pkg()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    orig, recon = make_sequences()
    lists = ref_lists(N_FRAMES)
    passes = [(poc, r, lists[poc - 1][r]) for poc in range(1, N_FRAMES + 1) for r in range(len(lists[poc - 1]))]
    assert len(passes) == 250
    n_pass = len(passes)

    # slots: 0..63 = original frames (poc-1), 64..127 = reconstructed frames 0..63
    ctx = pkg.AffineME(W, H, device=local_rank, num_slots=2 * N_FRAMES, max_in_flight=n_pass)
    pin_orig = pkg.PinnedArray(orig.shape, np.uint16)
    pin_orig.array[...] = orig
    pin_recon = {}
    for qp in QPS:
        pin_recon[qp] = pkg.PinnedArray(recon[qp].shape, np.uint16)
        pin_recon[qp].array[...] = recon[qp]
    host_res = [pkg.HostResult(ctx) for _ in range(n_pass)]

    def upload_all(qp):
        for f in range(N_FRAMES):
            ctx.upload(f, pin_orig.array[f], pkg.ROLE_CURRENT)
            ctx.upload(N_FRAMES + f, pin_recon[qp].array[f], pkg.ROLE_REFERENCE)

    def queue_all(qp, to_host):
        for k, (poc, r, refpoc) in enumerate(passes):
            lam = lambda_for(qp, poc)
            if to_host:
                ctx.search(poc - 1, N_FRAMES + refpoc, lam, host_res[k])
            else:
                ctx.search_device(poc - 1, N_FRAMES + refpoc, lam, k)

    # ---- device-resident timing ----
    def step_resident(step):
        qp = QPS[step % 4]
        upload_all(qp)          # untimed: inputs resident before the timed region
        ctx.sync()
        barrier()
        queue_all(qp, False)    # host-side bookkeeping only; nothing is launched before flush()
        ctx.timer_start()
        ctx.flush()
        ms = ctx.timer_stop()
        ctx.sync()
        barrier()
        return ms

    # frames per launch sequence in the end-to-end path: uploads of chunk k+1 and result copies of chunk k-1 overlap
    # the kernels of chunk k; the first upload and the last copy overlap nothing, so the first and last chunks are short
    CHUNK = 8
    CHUNKS = [int(x) for x in os.environ.get("AME_BENCH_CHUNKS", "").split(",") if x] or [2, 6] + [CHUNK] * ((N_FRAMES - 16) // CHUNK) + [6, 2]
    assert sum(CHUNKS) == N_FRAMES

    def step_e2e(step):
        qp = QPS[step % 4]
        barrier()
        t0 = time.perf_counter()
        ctx.timer_start()
        k = 0
        f0 = 0
        for chunk in CHUNKS:
            for f in range(f0, f0 + chunk):
                ctx.upload(f, pin_orig.array[f], pkg.ROLE_CURRENT)
                ctx.upload(N_FRAMES + f, pin_recon[qp].array[f], pkg.ROLE_REFERENCE)
            f0 += chunk
            while k < n_pass and passes[k][0] - 1 < f0:
                poc, r, refpoc = passes[k]
                ctx.search(poc - 1, N_FRAMES + refpoc, lambda_for(qp, poc), host_res[k])
                k += 1
            ctx.flush()
        ms = ctx.timer_stop()
        ctx.sync()
        wall = (time.perf_counter() - t0) * 1000.0
        barrier()
        return max(ms, wall)

    for s in range(args.warmup):
        step_resident(s)
    sampler = ClockSampler(local_rank)
    sampler.start()
    times = [step_resident(s) for s in range(args.steps)]
    kernel_ms, launches_per_flush = ctx.last_kernel_ms()
    sampler.stop_flag = True
    sampler.join()
    clocks = sampler.summary()
    total_ms = max_over_ranks(float(np.sum(times)))
    frames_per_s = world * N_FRAMES * args.steps / (total_ms / 1000.0)
    passes_per_s = frames_per_s * n_pass / N_FRAMES

    for s in range(min(args.warmup, 1)):
        step_e2e(s)
    e2e_times = [step_e2e(s) for s in range(args.steps)]
    e2e_ms = max_over_ranks(float(np.sum(e2e_times)))
    e2e_fps = world * N_FRAMES * args.steps / (e2e_ms / 1000.0)
    h2d = 2 * N_FRAMES * W * H * 2
    d2h = n_pass * sum(ctx.result_len(p) * (8 + 28) for p in range(4))

    # spot check of the last e2e step against... nothing on the CPU here (tests do that); sanity only
    assert int(host_res[0].cost[0][0]) > 0

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_mhz = clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
        peak_tops = 148 * 4 * 32 * sm_mhz * 1e6 / 1e12
        per_gpu_passes = passes_per_s / world
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["dram_bytes_per_ref_pass"] * n_pass
        except Exception:
            pass
        achieved = OPS_PER_PASS * per_gpu_passes / 1e12
        # the CPU baseline is timed at N = 1 only (the other ranks would spin on the barrier and take its cores)
        cb = cpu_port_rate(orig, recon[32], 32, N_CTUS // 15) if world == 1 else None
        line = {
            "metric": "1080p frames/sec affine ME", "value": frames_per_s, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "frames_per_step": N_FRAMES, "ref_passes_per_step": n_pass, "per_gpu": "same sequence on every rank",
                       "l2": "inputs (1.35 GB of planes per step) larger than L2; no flush"},
            "ref_passes_per_s": passes_per_s,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ref_passes_per_s": e2e_fps * n_pass / N_FRAMES},
            "gpu_launches": args.steps * launches_per_flush,
            "gpu_launches_e2e_per_step": len(CHUNKS) * launches_per_flush + 3 * N_FRAMES,
            "clocks": clocks,
            "roofline": {"bound": "int32_issue", "achieved": achieved, "peak": peak_tops, "unit": "Tlane-op/s",
                         "frac": achieved / peak_tops, "traffic": traffic,
                         "traffic_note": "DRAM bytes per launch sequence (= per step of 250 passes), extrapolated from the ncu launch list with DRAM counters "
                                         "recorded in profiles/r01_traffic.json; algorithmic bytes per step = %d" % (ALGO_BYTES_PER_PASS * n_pass),
                         "note": "SURVEY 8(d): compute-bound on INT32 issue; achieved = as-written op model (%.1f G lane-ops/pass) x "
                                 "passes/s per GPU; peak = 148 SM x 4 x 32 lanes x %.0f MHz sampled during the run; algorithmic DRAM "
                                 "bytes/pass = %d (%.1f GB/s, vs %.0f GB/s measured HBM peak)" % (
                                     OPS_PER_PASS / 1e9, sm_mhz, ALGO_BYTES_PER_PASS, ALGO_BYTES_PER_PASS * per_gpu_passes / 1e9,
                                     peaks.get("hbm_gbs", 6650.0)),
                         "frac_note": "the op model counts the work as the reference writes it; the exact shortcuts (early exit on a "
                                      "revisited state, shared first 2-CP evaluation, 3-CP start reuse) skip part of it, so frac can "
                                      "exceed 1 -- ncu_issue_active_pct is the hardware-side figure of the dominant kernel",
                         "ncu_issue_active_pct": ncu_issue_active(),
                         "kernel_ms_per_step": kernel_ms},
            "cpu_baseline": None if cb is None else {
                "value": cb[0], "unit": "frames/s", "cores": cb[3], "kind": "port",
                "sample": "one full 1080p reference pass (poc 1, ref 0, QP 32), %.1f s" % cb[2], "ref_passes_per_s": cb[1]},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def ncu_issue_active():
    """smsp__issue_active of ame_iter_small (60 % of the step) from the committed ncu --set full capture, or None."""
    try:
        for ln in open(os.path.join(ROOT, "profiles", "r01_final_ncu_ame_iter_small.txt")):
            if ln.startswith("smsp__issue_active.avg.pct_of_peak_sustained_active"):
                return float(ln.split()[-1])
    except OSError:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

"""Turns what tools/collect_evidence.sh left in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_evidence.py [tag]        (tag defaults to r01_final)

Writes  profiles/<tag>_bench.json, <tag>_bench_reference.json, <tag>_pytest_gpu.log,
        <tag>_ncu_launch_list_bench.csv + _summary.txt, <tag>_ncu_launch_list_sequence.txt (one launch sequence: time and
        DRAM bytes per kernel), <tag>_ncu_<kernel>.txt (selected raw metrics + per-function instruction / stall split) and
        profiles/r01_traffic.json (DRAM bytes per reference pass, read by bench.py)."""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_final"

RAW = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
       "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active"]


def launch_rows(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, out = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            out.append(dict(zip(hdr, r)))
    return out


def kernel_name(d):
    return d["Kernel Name"].split("(")[0]


def summarize_launch_list(src, dst, title):
    rows = launch_rows(src)
    per = collections.OrderedDict()
    ids = {}
    for d in rows:
        key = (kernel_name(d), d["Block Size"])
        v = float(d["Metric Value"].replace(",", ""))
        e = per.setdefault(key, collections.Counter())
        e[d["Metric Name"]] += v
        ids.setdefault(key, set()).add(d["ID"])
    tot = sum(e["gpu__time_duration.sum"] for e in per.values())
    with open(dst, "w") as f:
        f.write(title + "\n")
        f.write("launches captured: %d, total device time %.3f ms (per-launch times are cold-cache and serialised)\n" % (
            sum(len(v) for v in ids.values()), tot / 1e6))
        for key, e in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
            line = "%-22s block %-14s launches %4d  total %10.3f ms  share %5.1f%%" % (
                key[0], key[1], len(ids[key]), e["gpu__time_duration.sum"] / 1e6, 100.0 * e["gpu__time_duration.sum"] / tot)
            if "dram__bytes_read.sum" in e:
                line += "  DRAM read %8.3f GB  write %7.3f GB" % (e["dram__bytes_read.sum"] / 1e9, e["dram__bytes_write.sum"] / 1e9)
            f.write(line + "\n")
    return per, tot


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


def main():
    os.makedirs(PROF, exist_ok=True)
    # bench lines
    for src, dst in (("ev_bench.log", "_bench.json"), ("ev_bench_reference.log", "_bench_reference.json")):
        p = os.path.join(OUT, src)
        if os.path.exists(p):
            lines = [l for l in open(p) if l.startswith("{")]
            if lines:
                open(os.path.join(PROF, tag + dst), "w").write(lines[-1])
    for src, dst in (("ev_pytest_gpu.log", "_pytest_gpu.log"), ("ev_profile_run16.log", "_profile_run_16frames.log"),
                     ("ev_profile_run4.log", "_profile_run_4frames.log")):
        p = os.path.join(OUT, src)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(PROF, tag + dst))
    # launch list of the bench command
    p = os.path.join(OUT, "ev_launches_bench.csv")
    if os.path.exists(p):
        shutil.copy(p, os.path.join(PROF, tag + "_ncu_launch_list_bench.csv"))
        summarize_launch_list(p, os.path.join(PROF, tag + "_ncu_launch_list_bench_summary.txt"),
                              "command: ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 1 --warmup 3\n"
                              "(untimed uploads = pad / phase / block kernels; one step = one launch sequence over 250 searches)")
    # one launch sequence with DRAM bytes
    p = os.path.join(OUT, "ev_launches_seq16.csv")
    if os.path.exists(p):
        per, tot = summarize_launch_list(p, os.path.join(PROF, tag + "_ncu_launch_list_sequence.txt"),
                                         "command: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                                         "python tools/profile_run.py --frames 16 --reps 1\n(16 frames = 58 searches: 32 plane uploads + ONE launch sequence)")
        rows = launch_rows(p)
        dram = 0.0
        for d in rows:
            if d["Metric Name"].startswith("dram__bytes") and kernel_name(d).startswith("ame_"):
                dram += float(d["Metric Value"].replace(",", "")) * unit_scale(d["Metric Unit"])
        W, H, nct = 1920, 1080, 135
        algo = 2 * W * H * 2 + 2 * nct * (201 + 284) * (8 + 28)
        json.dump({"dram_bytes_per_ref_pass": int(dram / 58), "algorithmic_bytes_per_ref_pass": algo,
                   "source": "profiles/%s_ncu_launch_list_sequence.txt: dram__bytes_read.sum + dram__bytes_write.sum over every ame_* launch of one "
                             "launch sequence of 58 searches (1080p, 16 frames), cold caches, divided by 58" % tag,
                   "why_above_algorithmic": "a reference plane is kept as 2 x 16 pre-filtered int16 planes (200 MB at 1080p) and every iteration is its "
                                            "own set of launches, so the rows a search touches are fetched again per iteration; the per-CU state "
                                            "and moments (312 B per CU and iteration) travel through global memory as well"},
                  open(os.path.join(PROF, "r01_traffic.json"), "w"), indent=1)
    # full captures
    for k in ("ame_iter_small", "ame_iter_big", "ame_update_kernel", "ame_iter0_kernel"):
        rep = os.path.join(OUT, "ev_%s.ncu-rep" % k)
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        with open(os.path.join(PROF, "%s_ncu_%s.txt" % (tag, k)), "w") as f:
            f.write("ncu --set full --clock-control none --import-source on -k regex:%s (one launch of tools/profile_run.py --frames 16)\n" % k)
            f.write("kernel: %s\n" % vals[hdr.index("Kernel Name")])
            for m in RAW:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("%-70s %-12s %s\n" % (m, units[i], vals[i]))
            f.write("warp stall reasons (warp cycles per issued instruction):\n")
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                    if float(vals[i]) >= 0.05:
                        f.write("  %-30s %s\n" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), vals[i]))
            by = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_function.py"), rep, "0", k], capture_output=True, text=True)
            f.write("\nper source function (tools/ncu_by_function.py):\n" + by.stdout.split("--- instructions")[0])
            if by.returncode:
                f.write("(ncu_by_function failed: %s)\n" % by.stderr[-300:])
    print("summaries written to", PROF)


if __name__ == "__main__":
    main()

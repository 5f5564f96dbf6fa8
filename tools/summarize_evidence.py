"""Turns what tools/collect_evidence.sh left in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_evidence.py [tag]        (tag defaults to r02_final)

Writes  profiles/<tag>_bench*.json, <tag>_pytest_gpu.log, <tag>_ncu_launch_list_bench_summary.txt,
        <tag>_ncu_launch_list_sequence_<config>.txt (one launch sequence: time, DRAM bytes and executed instructions per
        kernel), <tag>_ncu_<kernel>.txt (selected raw metrics + per-function instruction / stall split),
        <tag>_sass_opcodes.txt and profiles/r02_inst_table.json (executed warp instructions and DRAM bytes per reference
        pass with the hash of the kernel sources they were measured on; read by bench.py, which refuses a stale table)."""
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
tag = sys.argv[1] if len(sys.argv) > 1 else "r02_final"

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


def inst_table_entry(path, n_passes, what):
    """Per-pass totals of one launch sequence captured with time / DRAM / instruction metrics."""
    import launch_summary
    per, ids = launch_summary.summarize(path)
    search = {k: v for k, v in per.items() if k.startswith("ame_") or k.startswith("ame::ame_")}
    prep = {k: v for k, v in per.items() if k not in search}
    inst = sum(e["smsp__inst_executed.sum"] for e in per.values())
    dram = sum(e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"] for e in per.values())
    return {"passes": n_passes, "warp_inst_per_pass": inst / n_passes, "dram_bytes_per_pass": dram / n_passes,
            "search_kernel_ms_serialised": sum(e["gpu__time_duration.sum"] for e in search.values()) / 1e6,
            "plane_preparation_ms_serialised": sum(e["gpu__time_duration.sum"] for e in prep.values()) / 1e6,
            "per_kernel": {k: {"launches": len(ids[k]), "ms": e["gpu__time_duration.sum"] / 1e6, "warp_inst": e["smsp__inst_executed.sum"],
                               "dram_bytes": e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"]} for k, e in per.items()},
            "source": what}


def main():
    import hashlib
    import launch_summary
    sys.path.insert(0, ROOT)
    os.makedirs(PROF, exist_ok=True)
    for src, dst in (("ev_bench.log", "_bench.json"), ("ev_bench_reference.log", "_bench_reference.json"), ("ev_bench_4k.log", "_bench_4k.json"),
                     ("ev_bench_8k.log", "_bench_8k.json")):
        p = os.path.join(OUT, src)
        if os.path.exists(p):
            lines = [l for l in open(p) if l.startswith("{")]
            if lines:
                open(os.path.join(PROF, tag + dst), "w").write(lines[-1])
    for src, dst in (("ev_pytest_gpu.log", "_pytest_gpu.log"), ("ev_profile_run16.log", "_profile_run_16frames.log")):
        p = os.path.join(OUT, src)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(PROF, tag + dst))
    # launch list of the bench command
    p = os.path.join(OUT, "ev_launches_bench.csv")
    if os.path.exists(p):
        open(os.path.join(PROF, tag + "_ncu_launch_list_bench_summary.txt"), "w").write(launch_summary.text(
            p, "command: ncu --metrics gpu__time_duration.sum --clock-control none -c 700 python bench.py --steps 1 --warmup 3\n"
               "(plane preparation = pad / phase / block kernels; one step = one launch sequence over 250 searches)") + "\n")
    # launch sequences with DRAM bytes and executed instructions -> summaries + the table bench.py reads
    import bench
    table = {"kernel_source_sha": bench.kernel_source_sha(), "head": (open(os.path.join(OUT, "ev_head.txt")).read().strip() if os.path.exists(os.path.join(OUT, "ev_head.txt")) else None),
             "metrics": "gpu__time_duration.sum, dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed.sum per launch (ncu --clock-control none), "
                        "summed over every kernel of one launch sequence incl. its plane preparation, divided by its passes", "configs": {}}
    cmd = "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none python tools/profile_run.py"
    for src, cfg, n, args in (("ev_launches_seq64.csv", "1080p", 250, "--frames 64"), ("ev_launches_seq64_4k.csv", "4k", 250, "--size 3840x2160 --frames 64"),
                              ("ev_launches_seq16_8k.csv", "8k", 58, "--size 7680x4320 --frames 16"), ("ev_launches_seq16.csv", "1080p_16frames", 58, "--frames 16")):
        p = os.path.join(OUT, src)
        if not os.path.exists(p):
            continue
        what = "%s %s --reps 1 (%d searches: plane uploads + ONE launch sequence)" % (cmd, args, n)
        open(os.path.join(PROF, "%s_ncu_launch_list_sequence_%s.txt" % (tag, cfg)), "w").write(launch_summary.text(p, "command: " + what) + "\n")
        table["configs"][cfg] = inst_table_entry(p, n, what)
    if table["configs"]:
        json.dump(table, open(os.path.join(PROF, "r02_inst_table.json"), "w"), indent=1)
    # full captures
    for k, sel in (("ame_iter_small", "ame_iter_small"), ("ame_iter_big", "ame_iter_bigILb1E"), ("ame_update_kernel2", "ame_update_kernelILi2E"),
                   ("ame_update_kernel3", "ame_update_kernelILi3E"), ("ame_iter0_kernel", "ame_iter0_kernel"), ("ame_emit_kernel", "ame_emit_kernel")):
        rep = os.path.join(OUT, "ev_%s.ncu-rep" % k)
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        with open(os.path.join(PROF, "%s_ncu_%s.txt" % (tag, k)), "w") as f:
            f.write("ncu --set full --clock-control none --import-source on -k regex:%s (one launch of tools/profile_run.py --frames 16)\n" % k.rstrip("23"))
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
            by = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_function.py"), rep, "0", sel], capture_output=True, text=True)
            f.write("\nper source function (tools/ncu_by_function.py):\n" + by.stdout.split("--- instructions")[0])
            if by.returncode:
                f.write("(ncu_by_function failed: %s)\n" % by.stderr[-300:])
    # SASS opcode histogram of the library that ran
    h = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_histogram.py")], capture_output=True, text=True).stdout
    open(os.path.join(PROF, tag + "_sass_opcodes.txt"), "w").write(h)
    print("summaries written to", PROF)


if __name__ == "__main__":
    main()

"""Per-kernel totals of an ncu launch-list CSV (--csv --log-file ...): launches, device time, DRAM bytes, executed
warp instructions.  usage: python tools/launch_summary.py file.csv [title]"""
import collections
import csv
import sys


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


SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9, "inst": 1.0}


def summarize(path):
    per, ids = collections.OrderedDict(), {}
    for d in launch_rows(path):
        key = d["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(d["Metric Value"].replace(",", "")) * SCALE.get(d["Metric Unit"], 1.0)
        per.setdefault(key, collections.Counter())[d["Metric Name"]] += v
        ids.setdefault(key, set()).add(d["ID"])
    return per, ids


def text(path, title=""):
    per, ids = summarize(path)
    tot = sum(e["gpu__time_duration.sum"] for e in per.values())
    out = [title] if title else []
    out.append("launches captured: %d, total device time %.3f ms (per-launch times are cold-cache and serialised)" % (sum(len(v) for v in ids.values()), tot / 1e6))
    for key, e in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        line = "%-26s launches %4d  total %10.3f ms  share %5.1f%%" % (key, len(ids[key]), e["gpu__time_duration.sum"] / 1e6, 100.0 * e["gpu__time_duration.sum"] / tot)
        if "dram__bytes_read.sum" in e:
            line += "  DRAM read %8.3f GB  write %7.3f GB" % (e["dram__bytes_read.sum"] / 1e9, e["dram__bytes_write.sum"] / 1e9)
        if "smsp__inst_executed.sum" in e:
            line += "  warp inst %9.1f M" % (e["smsp__inst_executed.sum"] / 1e6)
        out.append(line)
    return "\n".join(out)


if __name__ == "__main__":
    print(text(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""))

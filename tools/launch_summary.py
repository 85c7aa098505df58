"""Per-kernel summary of an `ncu --csv` launch list (gpu__time_duration.sum and, when present, dram__bytes_read/write.sum):
    python tools/launch_summary.py gpurun_out/<tag>_launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, gi, mi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Metric Name"), hdr.index("ID")
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("ss::", "")
    k = (name[:60], r[gi])
    a = agg.setdefault(k, {"ids": set(), "t": 0.0, "rd": 0.0, "wr": 0.0})
    a["ids"].add(r[ii])
    v = float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum":
        a["t"] += v / 1e3
    elif r[mi] == "dram__bytes_read.sum":
        a["rd"] += v
    elif r[mi] == "dram__bytes_write.sum":
        a["wr"] += v
tot = sum(a["t"] for a in agg.values())
print("%-62s %-14s %5s %10s %8s %6s %s" % ("kernel", "grid", "count", "total us", "mean us", "share", "DRAM read / written per launch"))
for k, a in sorted(agg.items(), key=lambda x: -x[1]["t"]):
    n = len(a["ids"])
    print("%-62s %-14s %5d %10.1f %8.2f %5.1f%% %s" % (k[0], k[1], n, a["t"], a["t"] / n, 100 * a["t"] / tot,
                                                        ("%.2f MB / %.2f MB" % (a["rd"] / n / 1e6, a["wr"] / n / 1e6)) if a["rd"] else ""))

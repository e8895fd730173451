"""Summarises an ncu launch-list CSV (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv):
   python experiments/launch_summary.py launches.csv [--list]"""
import csv, gzip, sys
from collections import defaultdict
path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(op(path, "rt")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
recs = {}
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    d = dict(zip(H, r))
    k = int(d["ID"])
    recs.setdefault(k, {"name": d["Kernel Name"], "grid": d["Grid Size"]})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
agg = defaultdict(lambda: [0.0, 0, 0.0])
for k in sorted(recs):
    r = recs[k]
    name = r["name"].split("(")[0].replace("void ", "").replace("b200::", "")[:70]
    a = agg[name]
    a[0] += r["gpu__time_duration.sum"] / 1e6
    a[1] += 1
    a[2] += (r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)) / 1e9
    if "--list" in sys.argv:
        print(k, name[:44], r["grid"], "%.1f us" % (r["gpu__time_duration.sum"] / 1e3),
              "%.0f MB" % ((r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)) / 1e6))
tot = sum(a[0] for a in agg.values())
print("| share | total ms | launches | avg us | DRAM GB | kernel |\n|---:|---:|---:|---:|---:|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| {100 * a[0] / tot:.2f}% | {a[0]:.3f} | {a[1]} | {1e3 * a[0] / a[1]:.1f} | {a[2]:.2f} | `{name}` |")
print(f"\nTotal device time {tot:.2f} ms over {sum(a[1] for a in agg.values())} launches")

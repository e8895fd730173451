"""DRAM bytes per training step of the tcgen05 launches, from an ncu launch list with dram__bytes_read/write
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`, experiments/r02_profile3.sh).

    python experiments/traffic_from_launches.py profiles/r02/r02_final_launches_cfg3.csv.gz 3

A step is delimited by the optimizer launch (adam_pack_kernel); the first steps of the capture (warm-up, packs) are
dropped and the median step is reported.  Writes profiles/r02_traffic_cfg<N>.json, which bench.py reads for
`roofline.traffic`."""
import csv
import gzip
import json
import os
import statistics
import sys

path, cfg = sys.argv[1], int(sys.argv[2])
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(op(path, "rt")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
iname, imet, iunit, ival, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}  # -> bytes / microseconds
launches = {}
for r in rows[hi + 1:]:
    if len(r) <= ival:
        continue
    d = launches.setdefault(int(r[iid]), {"name": r[iname], "bytes": 0.0, "us": 0.0})
    v = float(r[ival].replace(",", "")) * scale.get(r[iunit], 1.0)
    if r[imet].startswith("dram__bytes"):
        d["bytes"] += v
    elif r[imet].startswith("gpu__time_duration"):
        d["us"] += v
steps, cur = [], []
for k in sorted(launches):
    cur.append(launches[k])
    if "adam_pack_kernel" in launches[k]["name"]:
        steps.append(cur)
        cur = []
tc = lambda l: "umma_conv_kernel" in l["name"] or "wgrad_umma_kernel" in l["name"]
per_step = [(sum(l["bytes"] for l in s if tc(l)), sum(l["us"] for l in s if tc(l)), sum(1 for l in s if tc(l)),
             sum(l["bytes"] for l in s), sum(l["us"] for l in s)) for s in steps]
full = [p for p in per_step if p[2] == max(q[2] for q in per_step)]   # complete training steps only
med = lambda i: statistics.median(p[i] for p in full)
out = {"source": f"{path} (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, batch of the configuration; median of "
                 f"{len(full)} complete steps; under ncu every kernel runs alone with a cold L2)",
       "dram_bytes_per_step_tcgen05": med(0), "tcgen05_us_per_step_under_ncu": med(1), "tcgen05_launches_per_step": med(2),
       "dram_bytes_per_step_all_kernels": med(3), "us_per_step_all_kernels_under_ncu": med(4), "steps_in_capture": len(steps)}
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"r02_traffic_cfg{cfg}.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))

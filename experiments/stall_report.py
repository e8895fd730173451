"""Per-instruction stall samples of an .ncu-rep (source page): waits on mbarriers and the hottest instructions."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot)
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 150
for i, r in enumerate(data):
    n = int(r[isamp] or 0)
    nxt = int(data[i + 1][isamp] or 0) if i + 1 < len(data) else 0
    if "TRYWAIT" in r[isrc]:
        print(f"{i:5d} wait  {n + nxt:6d} ({100.0 * (n + nxt) / tot:5.1f}%) exec {r[iex]:>9s} {r[isrc][:90]}")
    elif n > thr and "BRA" not in r[isrc]:
        print(f"{i:5d} hot   {n:6d} ({100.0 * n / tot:5.1f}%) exec {r[iex]:>9s} {r[isrc][:90]}")

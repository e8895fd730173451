"""Per-instruction stall breakdown of one launch of an .ncu-rep (source page).
usage: stall_report2.py file.ncu-rep <launch index> [min samples]"""
import csv, subprocess, sys
rep, launch = sys.argv[1], int(sys.argv[2])
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in data)
print(rows[0][1][:100], "total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for i, h in stalls:
        agg[h] = agg.get(h, 0) + int(r[i] or 0)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot})
for idx, r in enumerate(data):
    n = int(r[isamp] or 0)
    if n >= thr:
        top = sorted(((int(r[i] or 0), h) for i, h in stalls), reverse=True)[:3]
        print(f"{idx:5d} {n:6d} ({100.0 * n / tot:4.1f}%) exec {r[iex]:>8s}  {r[isrc].strip()[:70]:70s} {[(h[6:], c) for c, h in top if c]}")

"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): last `n` launches, or totals per kernel name."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, out = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        out.append((d["Kernel Name"], d.get("Grid Size", ""), float(d["Metric Value"].replace(",", ""))))
tot = 0.0
for name, grid, v in out[-n:]:
    tot += v
    print(f"{name[:70]:72s} {grid:18s} {v / 1000:9.1f} us")
print(f"sum of the last {n}: {tot / 1000:.1f} us")

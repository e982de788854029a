"""Print one frame's kernel sequence (device time per launch) from an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
cols = rows[h]
ki, vi, gi = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Grid Size")
seq = [(r[ki][:70], float(r[vi].replace(",", "")), r[gi]) for r in rows[h + 2:] if len(r) > vi]
starts = [i for i, (k, _, _) in enumerate(seq) if "k_crop" in k or "k_front" in k]
a, b = starts[-2], starts[-1]
tot = 0.0
agg = {}
for k, v, g in seq[a:b]:
    print(f"{v / 1000:8.1f} us  {g:16s} {k}")
    tot += v
    key = k.split("(")[0].split("<")[0]
    agg[key] = agg.get(key, 0.0) + v
print("frame total us", round(tot / 1000, 1))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"   {v / 1000:8.1f} us  {k}")

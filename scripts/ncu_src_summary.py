"""Summarise an `ncu --page source --csv` export: opcode histogram by executed instructions and stall samples,
shared-memory wavefront excess, and the hottest SASS lines.  usage: ncu_src_summary.py file.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {k: i for i, k in enumerate(hdr)}
data = []
for r in rows[h + 1:]:          # first kernel section only
    if r and r[0] in ("Kernel Name", "Address"):
        break
    if len(r) == len(hdr):
        data.append(r)
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data) or 1
print("warp instructions", tot_inst, "samples", tot_samp, "SASS lines", len(data))
op, samp = collections.Counter(), collections.Counter()
for r in data:
    s = r[ix["Source"]].strip().split()
    o = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
    op[o] += int(r[ix["Instructions Executed"]])
    samp[o] += int(r[ix["# Samples"]])
for o, c in op.most_common(top):
    print(f"  {o:14s} {c:11d} {100 * c / tot_inst:5.1f}%   stall samples {100 * samp[o] / tot_samp:5.1f}%")
if "L1 Wavefronts Shared" in ix:
    tw = sum(int(r[ix["L1 Wavefronts Shared"]]) for r in data)
    exc = sum(int(r[ix["L1 Wavefronts Shared Excessive"]]) for r in data)
    print("shared wavefronts", tw, "excessive", exc)
print("hottest lines by samples:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
    print(f"  {int(r[ix['# Samples']]):6d}  {int(r[ix['Instructions Executed']]):9d}  {r[ix['Source']].strip()[:110]}")

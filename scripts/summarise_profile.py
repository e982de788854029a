"""Summarise an `ncu --set full` report of one frame's ReID kernels: per launch duration, tensor-pipe and issue activity,
DRAM bytes; per class totals -> profiles/traffic.json (read by bench.py for the roofline's `traffic` / `hbm` entries) and a
markdown table.   python scripts/summarise_profile.py gpurun_out/r2h_reid_full.ncu-rep profiles/r2h_ncu_reid.md [crops]"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

rep, out_md = sys.argv[1], sys.argv[2]
crops = float(sys.argv[3]) if len(sys.argv) > 3 else 208.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    try:
        return float(r[ix[name]].replace(",", ""))
    except Exception:
        return float("nan")


def cls_of(name):
    for key, c in (("k_front_tc", "front (crop + stem + pool)"), ("k_chain_tc", "chain (LightConv branches)"), ("k_gemm_tc", "gemm (1x1 / combine)"),
                   ("k_gates_tc", "gates"), ("k_head", "head")):
        if key in name:
            return c
    return "other"


# one frame = the launches from k_front_tc up to the next k_front_tc: rotate the capture so that it starts there and
# drop what exceeds one frame (the capture window may overlap the neighbouring frame by a launch)
body = rows[2:]
first = next((i for i, r in enumerate(body) if "k_front_tc" in r[ix["Kernel Name"]]), 0)
per_frame = int(sys.argv[4]) if len(sys.argv) > 4 else 26
extra = max(0, len(body) - per_frame)          # launches of the previous frame's tail that the window repeats
body = (body[first:] + body[extra:first])[:per_frame]
per = {}
lines = ["| # | kernel | grid | us | tensor pipe % | issue active % | warps active % | DRAM read MB | DRAM write MB |", "|---:|---|---|---:|---:|---:|---:|---:|---:|"]
for i, r in enumerate(body):
    name = r[ix["Kernel Name"]]
    c = cls_of(name)
    us = val(r, "gpu__time_duration.sum")
    if "msecond" in rows[1][ix["gpu__time_duration.sum"]]:
        us *= 1e3
    elif "nsecond" in rows[1][ix["gpu__time_duration.sum"]]:
        us /= 1e3
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    for k, u in (("dram__bytes_read.sum", None), ("dram__bytes_write.sum", None)):
        unit = rows[1][ix[k]]
        f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if k.endswith("read.sum"):
            rd *= f
        else:
            wr *= f
    tp = val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    ia = val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    wa = val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    lines.append(f"| {i} | {name.split('(')[0][-48:]} | {r[ix['Grid Size']]} | {us:.1f} | {tp:.1f} | {ia:.1f} | {wa:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} |")
    p = per.setdefault(c, {"us": 0.0, "dram_bytes": 0.0, "launches": 0, "tensor_weighted": 0.0})
    p["us"] += us
    p["dram_bytes"] += rd + wr
    p["launches"] += 1
    p["tensor_weighted"] += tp * us
tot_b = sum(p["dram_bytes"] for p in per.values())
tot_us = sum(p["us"] for p in per.values())
summ = ["", "| class | launches | us (serialised, under ncu) | DRAM MB | mean tensor pipe % |", "|---|---:|---:|---:|---:|"]
for c, p in per.items():
    summ.append(f"| {c} | {p['launches']} | {p['us']:.1f} | {p['dram_bytes'] / 1e6:.1f} | {p['tensor_weighted'] / max(p['us'], 1e-9):.1f} |")
summ.append(f"| all ReID kernels of one frame | {sum(p['launches'] for p in per.values())} | {tot_us:.1f} | {tot_b / 1e6:.1f} | "
            f"{sum(p['tensor_weighted'] for p in per.values()) / max(tot_us, 1e-9):.1f} |")
Path(out_md).write_text("\n".join(lines + summ) + "\n")
tj_path = Path(__file__).resolve().parents[1] / "profiles" / "traffic.json"
tj = json.loads(tj_path.read_text()) if tj_path.exists() else {}
tj["config2"] = {"crops_per_step": crops, "dram_bytes_per_step": tot_b, "source": f"ncu --set full, {Path(rep).name}, one frame, REID_SPLIT=1",
                 "per_class": {c: {"dram_bytes": p["dram_bytes"], "us_under_ncu": p["us"], "launches": p["launches"]} for c, p in per.items()}}
tj_path.write_text(json.dumps(tj, indent=1))
print("\n".join(summ))

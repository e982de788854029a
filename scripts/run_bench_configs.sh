#!/bin/bash
# default bench line (with cpu_baseline + parity) and short runs of BASELINE configs 3 / 4 / 5 on one GPU
mkdir -p gpurun_out
(time timeout 600 python bench.py > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err) 2>&1 | tail -3
tail -2 gpurun_out/r2g_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2g_bench_default.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "dtype")})
print("e2e", d["e2e"]["value"], d["e2e"]["api"], "pinned", d["e2e_pinned"]["value"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "parity", d.get("parity"))
print("config5", d.get("config5"))
print("cpu", d["cpu_baseline"]["value"], d.get("speedup_e2e_vs_cpu"))
PY
for c in "$@"; do
  timeout 500 python bench.py --config $c --steps 30 --warmup 5 --skip-cpu > gpurun_out/r2g_bench_c$c.json 2> gpurun_out/r2g_bench_c$c.err
  tail -1 gpurun_out/r2g_bench_c$c.err | cut -c1-300
  python - "$c" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2g_bench_c{sys.argv[1]}.json"))
print("config", sys.argv[1], {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "TF/s", d["roofline"]["achieved"],
      {k: round(v["ms_per_step"], 2) for k, v in d["kernel_classes"].items()})
PY
done

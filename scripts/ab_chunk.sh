mkdir -p gpurun_out
for c in 64 104 128 208 256; do
  BOXMOT_B200_REID_CHUNK=$c python bench.py --skip-cpu --steps 60 --warmup 10 > gpurun_out/ab_chunk$c.json 2> gpurun_out/ab_chunk$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab_chunk*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["e2e"]["value"],1), d["launches_per_step"], {k:round(v["ms_per_step"],3) for k,v in d["kernel_classes"].items()})
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-300:])
PY

# A/B runs of the ReID kernel generations through bench.py --skip-cpu (kernel experiments only):
#   BOXMOT_B200_LIGHT_CHAIN=0 per-level LightConv launches, BOXMOT_B200_CHAIN_VAR=1 stage-2 chain tiles of 16 rows,
#   BOXMOT_B200_LIGHT_V1=1 / BOXMOT_B200_PW_V1=1 first-generation kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_reid.py tests/test_gpu_pointwise_tc.py -x -q 2>&1 | tail -2
python bench.py --skip-cpu --steps 60 --warmup 10 > gpurun_out/ab_cur.json 2> gpurun_out/ab_cur.err
for v in ${VARIANTS:-BOXMOT_B200_CHAIN_VAR=1 BOXMOT_B200_LIGHT_CHAIN=0}; do
  env $v python bench.py --skip-cpu --steps 60 --warmup 10 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["e2e"]["value"],1), d["launches_per_step"], {k:round(v["ms_per_step"],3) for k,v in d["kernel_classes"].items()})
    except Exception as e: print(f, "ERR", e)
PY

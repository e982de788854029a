#!/bin/bash
# quick GPU confidence run after a ReID kernel change: stage taps vs oracle, ReID tests, short bench (no CPU leg),
# then the phase clocks of the instrumented build
python scripts/tc_stage_check.py osnet_x0_25 37 2>&1 | tail -4
python -m pytest tests/test_gpu_reid.py -x -q 2>&1 | tail -2
python bench.py --steps 150 --warmup 15 --skip-cpu --no-extra > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
python - <<'PY'
import json
a = json.load(open('gpurun_out/quick_bench.json'))
print('value', a['value'], 'e2e', a['e2e']['value'], {k: round(v['ms_per_step'], 4) for k, v in a['kernel_classes'].items() if v['ms_per_step']})
PY
if [ -f boxmot_b200/libboxmot_b200_clocks.so ]; then
  cp boxmot_b200/libboxmot_b200_clocks.so boxmot_b200/libboxmot_b200.so
  BOXMOT_B200_REID_SPLIT=1 python scripts/tc_clock_run.py 2>&1 | sed -n "/second pass/,\$p" > gpurun_out/quick_clocks.txt
fi

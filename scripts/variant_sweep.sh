#!/bin/bash
# on the GPU box: short bench of every boxmot_b200/libboxmot_b200_v*.so variant (copied over the product library in the
# box's scratch copy of the repo)
cp boxmot_b200/libboxmot_b200.so /tmp/orig.so
for f in boxmot_b200/libboxmot_b200_v*.so; do
  cp $f boxmot_b200/libboxmot_b200.so
  python bench.py --steps 150 --warmup 15 --skip-cpu --no-extra > /tmp/v.json 2>/tmp/v.err
  python - "$f" <<'PY'
import json, sys
try:
    a = json.load(open('/tmp/v.json'))
    print(sys.argv[1].split('_')[-1], 'value %.1f' % a['value'], 'e2e %.1f' % a['e2e']['value'], {k: round(v['ms_per_step'], 4) for k, v in a['kernel_classes'].items() if v['ms_per_step']}, flush=True)
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('/tmp/v.err').read()[-400:])
PY
done 2>&1 | tee gpurun_out/variant_sweep.txt
cp /tmp/orig.so boxmot_b200/libboxmot_b200.so

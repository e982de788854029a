#!/bin/bash
# bench value / e2e over ReID chunk sizes and slice counts (L2 residency of the inter-kernel tensors vs wave fill)
for ch in "$@"; do
 for sp in 1 2 3; do
  BOXMOT_B200_REID_CHUNK=$ch BOXMOT_B200_REID_SPLIT=$sp timeout 150 python bench.py --steps 100 --warmup 10 --skip-cpu --no-extra 2>/dev/null > /tmp/sw.json
  python - "$ch" "$sp" <<'PY'
import json, sys
d = json.load(open("/tmp/sw.json"))
print("chunk", sys.argv[1], "split", sys.argv[2], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
PY
 done
done

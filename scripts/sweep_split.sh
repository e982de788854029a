#!/bin/bash
# bench value / e2e for several ReID slice counts (BOXMOT_B200_REID_SPLIT)
for sp in "$@"; do
  BOXMOT_B200_REID_SPLIT=$sp timeout 150 python bench.py --steps 100 --warmup 10 --skip-cpu 2>/dev/null > /tmp/sw_$sp.json
  python - "$sp" <<'PY'
import json, sys
d = json.load(open(f"/tmp/sw_{sys.argv[1]}.json"))
print("split", sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
PY
done

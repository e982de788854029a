#!/bin/bash
# e2e (pageable frames) against the number of staging threads (BOXMOT_B200_STAGE_THREADS, caller included)
for t in "$@"; do
  BOXMOT_B200_STAGE_THREADS=$t timeout 150 python bench.py --steps 100 --warmup 10 --skip-cpu --no-extra 2>/dev/null > /tmp/st_$t.json
  python - "$t" <<'PY'
import json, sys
d = json.load(open(f"/tmp/st_{sys.argv[1]}.json"))
print("stage threads", sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["e2e"]["ms_per_step"], 3), "pinned", round(d["e2e_pinned"]["value"], 1))
PY
done

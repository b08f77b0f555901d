#!/bin/bash
# dev: timing-only ablations of fine_tc2 (outputs invalid): 1 no gather, 2 no weight stream, 8 no TMEM stores, 16 MMA-only skeleton (+32 TS form)
for ab in ${ABL:-0 1 2 3 8 16 48}; do
  echo -n "ablate=$ab : "
  EDN_TC2_ABLATE=$ab timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('fine ms', round(d['kernels_ms']['fine'], 4))
"
done

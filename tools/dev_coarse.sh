#!/bin/bash
# dev: coarse kernel timing under ablations (outputs invalid), run under gpurun
for ab in ${ABL:-0 1 2 3}; do
  EDN_COARSE_ABLATE=$ab timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ablate=$ab coarse ms', round(d['kernels_ms']['coarse'], 4))
"
done
EDN_COARSE_V1=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('v1 coarse ms', round(d['kernels_ms']['coarse'], 4))
"

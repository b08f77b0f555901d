#!/bin/bash
# Timing-only ablations of the fine tcgen05 kernel (profiles/r1_fine_tc_ablation.txt).  Rebuilds the library WITH the ablation
# branches compiled in, runs bench.py per setting, then restores the normal build.  Outputs of ablated runs are invalid.
set -e
cd "$(dirname "$0")/.."
EDN_NVCC_EXTRA=-DEDN_TC_ABLATE_BUILD=1 python evdeblurnerf_b200/csrc/build.py --force > /dev/null 2>&1
for ab in 0 1 2 4 8 3 5 6 7 15; do
  EDN_TC_ABLATE=$ab python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ablate', $ab, 'fine_ms', round(d['kernels_ms']['fine'],4), 'step_ms', round(d['ms_per_step'],4))"
done
python evdeblurnerf_b200/csrc/build.py --force > /dev/null 2>&1

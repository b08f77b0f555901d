#!/bin/bash
# Timing-only ablation of the VM gradient scatter (field_bwd.cu): which reds bound it?  Rebuilds field_bwd.cu with
# -DEDN_SCATTER_ABLATE=1 (no line reds) / 2 (no plane reds), prints the field-backward times, restores the normal build.
set -e
cd "$(dirname "$0")/.."
for ab in 0 1 2; do
  touch evdeblurnerf_b200/csrc/field_bwd.cu
  EDN_NVCC_EXTRA=-DEDN_SCATTER_ABLATE=$ab python evdeblurnerf_b200/csrc/build.py > /dev/null 2>&1
  python tools/bench_train_step.py --precision bf16 --steps 5 --warmup 2 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('scatter_ablate', $ab, 'step_ms', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernels_ms'].items() if 'bwd' in k})"
done
touch evdeblurnerf_b200/csrc/field_bwd.cu
python evdeblurnerf_b200/csrc/build.py > /dev/null 2>&1

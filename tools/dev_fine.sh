#!/bin/bash
# dev loop for the tensor-core fine kernel (run under gpurun): parity tests of the bf16 mode, then a short bench
timeout 600 python -m pytest tests/test_render_tc_gpu.py -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline 2>gpurun_out/dev_bench.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'kernels', {k: round(v, 4) for k, v in d['kernels_ms'].items()}, 'frac', round(d['roofline']['frac'], 4), 'errs', d['numerical_errors'])
"
tail -3 gpurun_out/dev_bench.err

#!/bin/bash
# Round-end verification on a B200 box: GPU tests, smoke, bench lines (ours bf16 / fp32, reference arm), training-step variants,
# memcheck over the test suite.  Outputs under gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 500 gpurun_out/bench_final.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref.json
rm -f gpurun_out/train_final.jsonl
for f in "" "--awp" "--events 2048" "--awp --events 2048"; do python tools/bench_train_step.py --precision bf16 $f 2>/dev/null | tail -1 >> gpurun_out/train_final.jsonl; done
cut -c1-260 gpurun_out/train_final.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_train_final.csv python tools/bench_train_step.py --steps 1 --warmup 1 --precision bf16 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_train_awp_final.csv python tools/bench_train_step.py --steps 1 --warmup 1 --precision bf16 --awp > /dev/null 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q -x --deselect tests/test_train_gpu.py::test_trainer_reduces_loss_on_a_fixed_batch --deselect tests/test_train_gpu.py::test_trainer_with_awp_reduces_loss > gpurun_out/memcheck_final.log 2>&1; tail -4 gpurun_out/memcheck_final.log

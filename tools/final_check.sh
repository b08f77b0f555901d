#!/bin/bash
# Round-end evidence on ONE B200 (outputs under gpurun_out/, the judged copies are committed under profiles/ as r2_*):
# GPU tests, smoke, bench lines (bf16 / tc32 / fp32, reference arm), training-step variants, ncu launch lists + full captures of the
# tensor-core kernels, in-kernel phase trace + ablations of the fine kernel, memcheck over smoke.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
cp $O/parity_achieved.json $O/r2_parity_achieved.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/r2_bench_bf16.json 2> $O/r2_bench_bf16.err; tail -c 300 $O/r2_bench_bf16.json
python bench.py --precision tc32 --no-train --no-cpu-baseline > $O/r2_bench_tc32.json 2>/dev/null
python bench.py --precision fp32 --no-train --no-cpu-baseline --steps 5 > $O/r2_bench_fp32.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2>/dev/null; cut -c1-200 $O/r2_bench_reference_arm.json
rm -f $O/r2_train_step.jsonl
for f in "" "--awp" "--events 2048" "--awp --events 2048"; do python tools/bench_train_step.py --precision bf16 $f 2>/dev/null | tail -1 >> $O/r2_train_step.jsonl; done
cut -c1-200 $O/r2_train_step.jsonl
python tools/bench_awp_forward.py --steps 5 2>/dev/null | tail -1 > $O/r2_awp_forward.json; cat $O/r2_awp_forward.json
# launch lists (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_bf16.csv python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches_train_step_awp_bf16.csv python tools/bench_train_step.py --steps 1 --warmup 1 --precision bf16 --awp > /dev/null 2>&1
python tools/ncu_launch_agg.py $O/r2_launches_bench_bf16.csv 12 > $O/r2_bench_kernel_shares.txt; cat $O/r2_bench_kernel_shares.txt
python tools/ncu_launch_agg.py $O/r2_launches_train_step_awp_bf16.csv 45 > $O/r2_train_step_awp_kernel_shares.txt
# full captures of the tensor-core kernels (one launch each)
ncu --set full --clock-control none --import-source on -k regex:"fine_fwd_tc2|coarse_fwd_tc" -s 6 -c 2 -o $O/r2_tc_kernels python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fine_fwd_tc3" -s 3 -c 1 -o $O/r2_tc3_kernel python bench.py --precision tc32 --steps 2 --warmup 3 --no-train --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py $O/r2_tc_kernels.ncu-rep > $O/r2_tc_kernels_ncu_summary.txt; python tools/ncu_summary.py $O/r2_tc3_kernel.ncu-rep >> $O/r2_tc_kernels_ncu_summary.txt
# fine kernel: phase trace + timing ablations
python tools/dev_trace.py 2>&1 | grep trace2 > $O/r2_fine_tc2_phase_trace.txt
ABL="0 1 2 3 8 16 48" bash tools/dev_ablate.sh > $O/r2_fine_tc2_ablation.txt 2>&1; cat $O/r2_fine_tc2_ablation.txt
# coarse kernel: producer / epilogue time line of CTA 0 and timing ablations (1 no gather, 2 no PE / bias; outputs invalid)
EDN_COARSE_TRACE=1 python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline 2>/dev/null | grep coarse2 | tail -16 > $O/r2_coarse_tc2_phase_trace.txt
ABL="0 1 2 3" bash tools/dev_coarse.sh > $O/r2_coarse_tc2_ablation.txt 2>&1; cat $O/r2_coarse_tc2_ablation.txt
# sweep (config 5), one GPU
python tools/sweep.py > $O/r2_sweep.jsonl 2>/dev/null; wc -l $O/r2_sweep.jsonl
# memcheck over smoke (the three precisions + the backward)
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_memcheck_smoke.log 2>&1; tail -3 $O/r2_memcheck_smoke.log

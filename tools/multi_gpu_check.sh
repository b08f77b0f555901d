#!/bin/bash
# Round-end multi-GPU evidence (run under `gpurun --gpus N`): NCCL training equivalence tests (N >= 2), bench.py under torchrun at N
# (forward weak scaling + train_step with the gradient all-reduce + the strong-scaling leg), config-5 sweep at N.
N=${1:-2}
O=gpurun_out
mkdir -p $O
set -x
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q -k "nccl or sync" 2>&1 | tail -5 > $O/r2_tests_nccl_n2.log; cat $O/r2_tests_nccl_n2.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
  2> $O/r2_bench_n$N.err | grep '^{' > $O/r2_bench_n$N.json
tail -3 $O/r2_bench_n$N.err
python - <<PY
import json
d = json.loads(open("$O/r2_bench_n$N.json").read().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "train_step", "strong", "shipped_forward")})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/sweep.py --quick \
  2> $O/r2_sweep_n$N.err | grep '^{' > $O/r2_sweep_n$N.jsonl
wc -l $O/r2_sweep_n$N.jsonl

for ab in 0 1 2 4 8 3 5 6 7 15; do
  EDN_TC_ABLATE=$ab python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ablate', $ab, 'fine_ms', round(d['kernels_ms']['fine'],4), 'step_ms', round(d['ms_per_step'],4))"
done

#!/bin/bash
# 1 GPU: host threads of the index expansion in the end-to-end arm
set -u
OUT=gpurun_out
mkdir -p $OUT
nproc
for t in 0 4 6 12; do
  python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-sweep --no-full --option host_threads=$t 2>/dev/null | tail -1 > $OUT/r02y_bench_threads$t.json
  python - $OUT/r02y_bench_threads$t.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); e = d['e2e']
print(sys.argv[1], 'e2e %.1f ms (median %.1f)' % (e['ms_per_step'], e['step_ms_median_rank0']), e['step_ms_rank0'])
PY
done

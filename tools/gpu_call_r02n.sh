#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-full 2>/dev/null | tail -1 > $OUT/r02n_bench_sweep.json
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-full --option pipeline_ramp=0 2>/dev/null | tail -1 > $OUT/r02n_bench_sweep_ramp0.json
python - <<'PY'
import json
for f in ('r02n_bench_sweep', 'r02n_bench_sweep_ramp0'):
    d = json.load(open(f'gpurun_out/{f}.json'))
    for s in d['sweep']:
        print(f, s['faces'], s['dtype'], 'ms/step %.2f trace %.2f' % (s['ms_per_step'], s['trace_ms_per_launch']), s.get('step_ms_wall_rank0'))
PY

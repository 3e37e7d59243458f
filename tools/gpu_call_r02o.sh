#!/bin/bash
# Round 2, 1 GPU: emit kernel with 2 / 3 / 4 entries per lane; column loads of the trace kernel through L2 only; sweep without re-allocations
set -u
OUT=gpurun_out
mkdir -p $OUT
for v in default e3 e4 c1 c2; do
  so=fluxpy_b200/libfluxb200_$v.so
  [ $v = default ] && so=fluxpy_b200/libfluxb200.so
  echo "== $v"
  FLUXB200_SO=$PWD/$so PROF_ONE_REPS=5 python tools/prof_one.py 4096 317 2>&1 | grep "^rep" | tail -3
done | tee $OUT/r02o_variants.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-full 2>/dev/null | tail -1 > $OUT/r02o_bench_sweep.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02o_bench_sweep.json'))
print('value %.4e ms/step %.2f e2e %.1f ms' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']), d['e2e']['step_ms_rank0'])
for s in d['sweep']:
    print(s['faces'], s['dtype'], 'ms/step %.2f trace %.2f' % (s['ms_per_step'], s['trace_ms_per_launch']), s.get('step_ms_wall_rank0'))
PY

#!/bin/bash
# 1 GPU, last call of round 2: the committed tree as the driver will run it (smoke, gpu tier, both bench arms)
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200 | tee $OUT/r02w_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $OUT/r02w_tests.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > $OUT/r02w_bench_reference.json
python bench.py --gpus 1 --steps 20 --warmup 5 2>$OUT/r02w_bench.err | tail -1 > $OUT/r02w_bench_full.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02w_bench_full.json')); r = json.load(open('gpurun_out/r02w_bench_reference.json'))
print('value %.4e ms/step %.2f frac %.3f traffic %s | e2e %.4e %.1f ms %s | parity %s | clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['step_ms_rank0'], d['parity_check']['ok'], d['clocks']))
print('reference %.4e (%d threads): ratio %.0f e2e ratio %.0f' % (r['value'], r['cpu_baseline']['omp_threads_used'], d['value']/r['value'], d['e2e']['value']/r['value']))
for s in d['sweep']: print(s['faces'], s['dtype'], '%.3e' % s['pairs_per_s'], '%.2f ms' % s['ms_per_step'], s.get('step_ms_wall_rank0'))
print(d['full_matrix']['t_total_s'], d['gpu_launches'])
PY

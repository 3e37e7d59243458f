#!/bin/bash
# Round 2, 1 GPU: packed pairs in phase C only (shipped), pipeline ramp + fill-before-trace, emit without fp64 maxima.
set -u
OUT=gpurun_out
mkdir -p $OUT
T=${TAG:-r02l}
echo "== gpu tier"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${T}_tests_default.log
echo "== bench"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sweep --no-full 2>$OUT/${T}_bench_ramp1.err | tail -1 > $OUT/${T}_bench_ramp1.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sweep --no-full --option pipeline_ramp=0 2>/dev/null | tail -1 > $OUT/${T}_bench_ramp0.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sweep --no-full --option sub_rows=256 2>/dev/null | tail -1 > $OUT/${T}_bench_ramp1_sub256.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sweep --no-full --option sub_rows=384 2>/dev/null | tail -1 > $OUT/${T}_bench_ramp1_sub384.json
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/r02l_bench_*.json')):
    try:
        d = json.load(open(f))
        print(f, 'value %.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], 'e2e ms %.1f' % d['e2e']['ms_per_step'],
              'trace ms %.2f' % d['roofline']['launch_ms'], 'frac %.3f' % d['roofline']['frac'], d.get('parity_check', {}).get('ok'), d['e2e'].get('step_ms_rank0'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
echo "== timeline"
python tools/diag_timeline.py 2>&1 | tail -9 | tee $OUT/${T}_timeline.log
for f in 5 6 7; do mv $OUT/timeline_rep$f.csv $OUT/${T}_timeline_rep$f.csv; done
echo "== fill kernels"
ncu --set full --clock-control none --import-source on -k regex:'emit_kernel|unpermute' -s 2 -c 2 -o $OUT/${T}_fill \
    python tools/prof_one.py 4096 317 > $OUT/${T}_prof_fill.log 2>&1
ls -la $OUT | tail -5

#!/bin/bash
# Third GPU call of round 2: second-generation trace kernel (trace2.cuh: lean front end + shared-memory stacks)
# and the new emit kernel against the first generation.
#   gpurun --timeout 2400 -- 'bash tools/gpu_call_r02c.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
T=${TAG:-r02c}
echo "== sanitizer first (small case; queues, shared-memory atomics)"
timeout 600 compute-sanitizer --tool memcheck  python tools/sanitize_case.py horizon_zone=32 2>&1 | tail -3 | tee $OUT/${T}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 40 python tools/sanitize_case.py horizon_zone=32 > $OUT/${T}_sanitizer_racecheck_full.log 2>&1; tail -3 $OUT/${T}_sanitizer_racecheck_full.log | tee $OUT/${T}_sanitizer_racecheck.log; grep -A6 "hazard detected" $OUT/${T}_sanitizer_racecheck_full.log | head -60
echo "== gpu tier, default (trace_variant 2, horizon skip on)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/${T}_tests_default.log
echo "== parity files with the first-generation kernel"
FLUXB200_TEST_VARIANT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${T}_tests_variant1.log
echo "== bench A/B (short arms)"
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sweep --no-full 2>$OUT/${T}_bench_v2.err | tail -1 > $OUT/${T}_bench_v2.json
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sweep --no-full --option trace_variant=1 2>/dev/null | tail -1 > $OUT/${T}_bench_v1.json
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/r02c_bench_v*.json')):
    try:
        d = json.load(open(f))
        print(f, 'value %.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'],
              'trace ms %.2f' % d['roofline']['launch_ms'], 'frac %.3f' % d['roofline']['frac'], d['config'].get('trace_counters'), d.get('parity_check'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
echo "== the full default bench line (sweep, full matrix, cpu baseline)"
( time python bench.py ) > $OUT/${T}_bench_full.json 2> $OUT/${T}_bench_full.err
tail -c 600 $OUT/${T}_bench_full.err
echo "== launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep --no-full > $OUT/${T}_bench_under_ncu.log 2>&1
echo "== full capture of the trace kernel (second repetition)"
ncu --set full --clock-control none --import-source on -k regex:trace2_kernel -s 1 -c 1 -o $OUT/${T}_trace2 \
    python tools/prof_one.py 4096 317 > $OUT/${T}_prof_one.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:emit_kernel -s 1 -c 1 -o $OUT/${T}_emit \
    python tools/prof_one.py 4096 317 > /dev/null 2>&1
ls -la $OUT | tail -20

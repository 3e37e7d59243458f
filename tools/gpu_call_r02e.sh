#!/bin/bash
# Fifth GPU call of round 2 (1 GPU): the build that ships -- tests, bench line, final ncu captures, block assembly.
set -u
OUT=gpurun_out
mkdir -p $OUT
T=${TAG:-r02e}
echo "== gpu tier"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${T}_tests_default.log
echo "== bench A/B"
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sweep --no-full 2>$OUT/${T}_bench.err | tail -1 > $OUT/${T}_bench_v2.json
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sweep --no-full --option trace_variant=1 2>/dev/null | tail -1 > $OUT/${T}_bench_v1.json
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/r02e_bench_v*.json')):
    try:
        d = json.load(open(f))
        print(f, 'value %.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'],
              'trace ms %.2f' % d['roofline']['launch_ms'], 'frac %.3f' % d['roofline']['frac'], d['config']['trace_counters'], d.get('parity_check', {}).get('ok'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
echo "== the full default bench line"
( time python bench.py --steps 20 --warmup 5 ) > $OUT/${T}_bench_full.json 2> $OUT/${T}_bench_full.err
tail -c 300 $OUT/${T}_bench_full.err
echo "== reference arm"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${T}_bench_reference.json 2>/dev/null
echo "== block assembly"
timeout 600 python tools/bench_blocks.py > $OUT/${T}_blocks.json 2> $OUT/${T}_blocks.err; cat $OUT/${T}_blocks.json
echo "== launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep --no-full > $OUT/${T}_bench_under_ncu.log 2>&1
echo "== full captures (second repetition): trace, emit, unpermute"
ncu --set full --clock-control none --import-source on -k regex:trace2_kernel -s 1 -c 1 -o $OUT/${T}_trace2 \
    python tools/prof_one.py 4096 317 > $OUT/${T}_prof_one.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"emit_kernel|unpermute" -s 2 -c 2 -o $OUT/${T}_fill \
    python tools/prof_one.py 4096 317 > /dev/null 2>&1
ls -la $OUT | tail -14

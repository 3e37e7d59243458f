#!/bin/bash
# First GPU call of a round: everything needed to decide whether the trace kernel's horizon skip becomes the
# default, in ONE gpurun (about 12 minutes of box time).  Run from the repo root:
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
#
# Writes into gpurun_out/ (merged back): test logs, bench lines off / on (+ zone sizes), the launch list and
# one full ncu capture of the horizon variant of K4, the horizon kernel's own duration.
set -u
OUT=gpurun_out
mkdir -p $OUT
T=${TAG:-r02}
echo "== gpu tier, default path";            python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/${T}_tests_default.log
echo "== gpu tier, horizon skip on (Z=1023)"; FLUXB200_TEST_HORIZON=1023 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/${T}_tests_horizon.log
echo "== bench, default";     python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $OUT/${T}_bench_off.json
echo "== bench, horizon on";  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --option horizon_skip=1 2>/dev/null | tail -1 > $OUT/${T}_bench_on_z1023.json
for Z in 256 512; do
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --option horizon_zone=$Z --option horizon_skip=1 2>/dev/null | tail -1 > $OUT/${T}_bench_on_z$Z.json
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/*_bench_o*.json')):
    try:
        d = json.load(open(f))
        print(f, 'value %.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'],
              'trace ms %.2f' % d['roofline']['launch_ms'], d['config'].get('trace_counters'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
echo "== launch list (durations) of one 4096-row slab, horizon on: K4, horizon_kernel, zone_kernel, col_horizon_kernel"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${T}_launches_horizon.csv \
    python tools/prof_one.py 4096 317 horizon_skip=1 > $OUT/${T}_prof_one.log 2>&1
grep -E "horizon_kernel|zone_kernel|col_horizon|trace_kernel" $OUT/${T}_launches_horizon.csv | awk -F'","' '{print $5, $NF}' | head -12
echo "== compute-sanitizer on the horizon variant (small case, zone 32 so that the skip is taken)"
compute-sanitizer --tool memcheck  python tools/sanitize_case.py horizon_zone=32 horizon_skip=1 2>&1 | tail -3 | tee $OUT/${T}_sanitizer_memcheck_horizon.log
compute-sanitizer --tool racecheck python tools/sanitize_case.py horizon_zone=32 horizon_skip=1 2>&1 | tail -3 | tee $OUT/${T}_sanitizer_racecheck_horizon.log
echo "== full capture of the horizon variant of K4 (second repetition)"
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o $OUT/${T}_trace_horizon \
    python tools/prof_one.py 4096 317 horizon_skip=1 > /dev/null 2>&1
ls -la $OUT | tail -12

#!/bin/bash
# Round 2, 1 GPU (final sources): what the driver will run at round end, on the build that ships, plus the ncu evidence for it.
set -u
OUT=gpurun_out
mkdir -p $OUT
T=${TAG:-r02m}
echo "== smoke"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/${T}_smoke.log
echo "== gpu tier"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${T}_tests_default.log
echo "== sanitizer"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py horizon_zone=32 2>&1 | tail -2 | tee $OUT/${T}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py horizon_zone=32 2>&1 | tail -2 | tee $OUT/${T}_sanitizer_racecheck.log
echo "== bench (driver arguments)"
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > $OUT/${T}_bench_full.json 2> $OUT/${T}_bench_full.err
tail -c 200 $OUT/${T}_bench_full.err
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $OUT/${T}_bench_reference.json 2> $OUT/${T}_bench_reference.err
tail -c 120 $OUT/${T}_bench_reference.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02m_bench_full.json'))
print('value %.4e ms/step %.2f e2e %.4e (%.1f ms, median %.1f) frac %.3f launch %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['step_ms_median_rank0'], d['roofline']['frac'], d['roofline']['launch_ms']))
print('e2e steps', d['e2e']['step_ms_rank0'], 'retries', d['e2e']['output_buffer_retries'])
for s in d['sweep']:
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in s.items() if k in ('faces', 'dtype', 'pairs_per_s', 'ms_per_step', 'trace_ms_per_launch', 'fill_ms_per_step', 'roofline_frac')})
print('full', {k: d['full_matrix'][k] for k in ('t_build_s', 't_assemble_s', 't_gather_s', 't_total_s')}, 'parity', d['parity_check']['ok'], 'traffic', d['roofline']['traffic'])
r = json.load(open('gpurun_out/r02m_bench_reference.json'))
print('reference arm %.4e pairs/s, %d threads; ratio %.0f, e2e ratio %.0f' % (r['value'], r['cpu_baseline']['omp_threads_used'], d['value']/r['value'], d['e2e']['value']/r['value']))
PY
echo "== block assembly"
timeout 600 python tools/bench_blocks.py > $OUT/${T}_blocks.json 2> $OUT/${T}_blocks.err; cat $OUT/${T}_blocks.json
echo "== launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep --no-full > $OUT/${T}_bench_under_ncu.log 2>&1
echo "== full captures (second repetition)"
ncu --set full --clock-control none --import-source on -k regex:trace2_kernel -s 1 -c 1 -o $OUT/${T}_trace2 \
    python tools/prof_one.py 4096 317 > $OUT/${T}_prof_one.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"emit_kernel|unpermute" -s 2 -c 2 -o $OUT/${T}_fill \
    python tools/prof_one.py 4096 317 > /dev/null 2>&1
ls -la $OUT | tail -12

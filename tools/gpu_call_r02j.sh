#!/bin/bash
# Round 2, 1 GPU: the packed-FP32 child test (pair-aligned record layout).
set -u
OUT=gpurun_out
mkdir -p $OUT
T=${TAG:-r02j}
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sweep --no-full 2>$OUT/${T}_bench.err | tail -1 > $OUT/${T}_bench_v2.json
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sweep --no-full --option trace_variant=1 2>/dev/null | tail -1 > $OUT/${T}_bench_v1.json
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/r02j_bench_v*.json')):
    try:
        d = json.load(open(f))
        print(f, 'value %.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'],
              'trace ms %.2f' % d['roofline']['launch_ms'], 'frac %.3f' % d['roofline']['frac'], d.get('parity_check', {}).get('ok'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
echo "== gpu tier"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${T}_tests_default.log
FLUXB200_TEST_VARIANT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | tee $OUT/${T}_tests_variant1.log
echo "== full capture"
ncu --set full --clock-control none --import-source on -k regex:trace2_kernel -s 1 -c 1 -o $OUT/${T}_trace2 \
    python tools/prof_one.py 4096 317 > $OUT/${T}_prof_one.log 2>&1
ls -la $OUT | tail -5

#!/bin/bash
# Round 2, 1 GPU: scalar / packed box+slab test per phase (A/B of four builds), e2e timeline, sub-slab sizes.
set -u
OUT=gpurun_out
mkdir -p $OUT
T=${TAG:-r02k}
for v in p0 p1 p2 p3; do
  so=fluxpy_b200/libfluxb200_$v.so
  [ $v = p0 ] && so=fluxpy_b200/libfluxb200.so
  FLUXB200_SO=$PWD/$so python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sweep --no-full 2>$OUT/${T}_bench_$v.err | tail -1 > $OUT/${T}_bench_$v.json
done
for s in 256 1024; do
  python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-sweep --no-full --option sub_rows=$s 2>/dev/null | tail -1 > $OUT/${T}_bench_sub$s.json
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/r02k_bench_*.json')):
    try:
        d = json.load(open(f))
        print(f, 'value %.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], 'e2e ms %.1f' % d['e2e']['ms_per_step'],
              'trace ms %.2f' % d['roofline']['launch_ms'], 'frac %.3f' % d['roofline']['frac'], d.get('parity_check', {}).get('ok'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
echo "== timeline"
python tools/diag_timeline.py 2>&1 | tail -9 | tee $OUT/${T}_timeline.log
ls -la $OUT | tail -5

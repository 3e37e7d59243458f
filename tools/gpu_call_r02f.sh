#!/bin/bash
# Sixth GPU call of round 2 (1 GPU): phase-C loop form variant, sweep with per-stage times, pageable full host matrix.
set -u
OUT=gpurun_out
mkdir -p $OUT
T=${TAG:-r02f}
for V in "" _c1; do
  FLUXB200_SO=$PWD/fluxpy_b200/libfluxb200$V.so python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sweep --no-full 2>$OUT/${T}_bench$V.err | tail -1 > $OUT/${T}_bench_var$V.json
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/r02f_bench_var*.json')):
    try:
        d = json.load(open(f))
        print(f, 'value %.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'],
              'trace ms %.2f' % d['roofline']['launch_ms'], 'frac %.3f' % d['roofline']['frac'], d.get('parity_check', {}).get('ok'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
echo "== gpu tier (parity + meshes files)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${T}_tests_parity.log
echo "== full default bench line with the host CSR of G(159)"
( time python bench.py --steps 20 --warmup 5 --full-host-grid 159 ) > $OUT/${T}_bench_full.json 2> $OUT/${T}_bench_full.err
tail -c 300 $OUT/${T}_bench_full.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02f_bench_full.json'))
print('value %.4e ms/step %.2f e2e %.4e (%.1f ms) frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))
print('e2e steps', d['e2e']['step_ms_rank0'])
for s in d['sweep']:
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in s.items() if k in ('faces', 'dtype', 'pairs_per_s', 'ms_per_step', 'trace_ms_per_launch', 'fill_ms_per_step', 'prepare_ms_per_step', 'roofline_frac')})
print('full', d['full_matrix'])
print('host', d['full_matrix_host_csr'])
PY
echo "== block assembly"
timeout 600 python tools/bench_blocks.py > $OUT/${T}_blocks.json 2> $OUT/${T}_blocks.err; cat $OUT/${T}_blocks.json
echo "== ncu of the c1 variant"
FLUXB200_SO=$PWD/fluxpy_b200/libfluxb200_c1.so ncu --set full --clock-control none --import-source on -k regex:trace2_kernel -s 1 -c 1 -o $OUT/${T}_trace2_c1 \
    python tools/prof_one.py 4096 317 > $OUT/${T}_prof_one.log 2>&1
ls -la $OUT | tail -8

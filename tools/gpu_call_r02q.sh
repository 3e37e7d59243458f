#!/bin/bash
# 8-GPU call on the final sources of round 2:  gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_call_r02q.sh'
# The bench line at N = 8 with the driver's arguments, the pipeline ramp on / off end to end, the 2-rank tests.
set -u
OUT=gpurun_out
mkdir -p $OUT
T=${TAG:-r02q}
run() { # name, extra args
  name=$1; shift
  ( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus 8 "$@" ) > $OUT/${T}_bench_$name.json 2> $OUT/${T}_bench_$name.err
  python - "$OUT/${T}_bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d['e2e']
    print(sys.argv[1], 'value %.4e ms/step %.2f | e2e %.4e %.1f ms floor %.1f ms (x%.2f) host-memory floor %.1f | parity %s'
          % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['pcie_floor_ms'], e['ms_per_step_over_floor'],
             e.get('host_memory_floor_ms', -1), d['parity_check'].get('ok')), e.get('step_ms_rank0'))
    if d.get('full_matrix'):
        print('   full matrix', {k: d['full_matrix'][k] for k in ('t_build_s', 't_assemble_s', 't_gather_s', 't_total_s', 'nnz')})
    for s in d.get('sweep') or []:
        print('   sweep', s['faces'], s['dtype'], '%.3e pairs/s, %.2f ms/step' % (s['pairs_per_s'], s['ms_per_step']))
except Exception as ex:
    print(sys.argv[1], 'unreadable', ex)
PY
}
run default --steps 20 --warmup 5
run ramp0 --steps 8 --warmup 4 --no-sweep --no-full --no-cpu-baseline --option pipeline_ramp=0
echo "== 2-rank tests"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${T}_tests_multi.log
echo "== reference arm under torchrun"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --impl reference --gpus 8 --steps 20 --warmup 5 2>/dev/null | tail -1 | tee $OUT/${T}_bench_reference_n8.json | cut -c1-600
ls -la $OUT | tail -8

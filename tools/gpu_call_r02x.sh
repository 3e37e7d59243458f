#!/bin/bash
# 8 GPUs, the end-to-end arm with the final block sizing (arena log on)
set -u
OUT=gpurun_out
mkdir -p $OUT
free -g | head -2 > $OUT/r02x_free.log
FLUXB200_ARENA_LOG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 8 --steps 10 --warmup 5 --no-sweep --no-full --no-cpu-baseline 2>$OUT/r02x_n8.err | tail -1 > $OUT/r02x_bench_n8.json
grep "fluxb200 arena" $OUT/r02x_n8.err | sort | uniq -c | sort -rn | head -12
free -g | head -2
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02x_bench_n8.json').read().strip().splitlines()[-1]); e = d['e2e']
print('value %.4e | e2e %.4e %.1f ms, floors pcie %.1f host %.1f, parity %s' % (d['value'], e['value'], e['ms_per_step'], e['pcie_floor_ms'], e['host_memory_floor_ms'], d['parity_check']['ok']), e['step_ms_rank0'], e['output_buffer_retries'])
PY

#!/bin/bash
# 2 GPUs, end-to-end arm with the arena log: what locks memory in a timed step?
set -u
OUT=gpurun_out
mkdir -p $OUT
FLUXB200_ARENA_LOG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 2 --steps 6 --warmup 5 --no-sweep --no-full --no-cpu-baseline 2>$OUT/r02t_n2.err | tail -1 > $OUT/r02t_bench_n2.json
grep "fluxb200" $OUT/r02t_n2.err | head -80
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02t_bench_n2.json').read().strip().splitlines()[-1])
print(d['e2e']['step_ms_rank0'], d['e2e']['output_buffer_retries'])
PY

#!/bin/bash
# 8 GPUs: column indices written by host threads from visibility words (default) against indices copied from the device
set -u
OUT=gpurun_out
mkdir -p $OUT
for he in 1 0; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus 8 --steps 8 --warmup 4 --no-sweep --no-full --no-cpu-baseline --option host_expand=$he 2>$OUT/r02r_he$he.err | tail -1 > $OUT/r02r_bench_n8_host_expand$he.json
  python - $OUT/r02r_bench_n8_host_expand$he.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e = d['e2e']
print(sys.argv[1], 'e2e %.1f ms/step, d2h %.2f GB/step/rank, pcie floor %.1f ms, host-memory floor %.1f ms, parity %s' % (e['ms_per_step'], e['d2h_bytes_per_step']/1e9, e['pcie_floor_ms'], e['host_memory_floor_ms'], d['parity_check']['ok']), e['step_ms_rank0'])
PY
done

#!/bin/bash
# N GPUs (N = $1) on the final sources, the driver's arguments
set -u
N=$1
OUT=gpurun_out
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 5 2>$OUT/r02s_bench_n$N.err | tail -1 > $OUT/r02s_bench_n$N.json
python - $OUT/r02s_bench_n$N.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e = d['e2e']
print(sys.argv[1], 'value %.4e ms/step %.2f | e2e %.4e %.1f ms/step, pcie floor %.1f ms, host-memory floor %.1f ms, parity %s' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['pcie_floor_ms'], e['host_memory_floor_ms'], d['parity_check']['ok']))
print('   full matrix', {k: d['full_matrix'][k] for k in ('t_build_s', 't_assemble_s', 't_gather_s', 't_total_s')}, 'cpu', d['cpu_baseline'])
PY

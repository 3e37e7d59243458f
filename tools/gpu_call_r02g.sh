#!/bin/bash
# 2-GPU call of round 2:  gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_call_r02g.sh'
# The N > 1 parity test the driver's 1-GPU box skips, and the bench line at N = 2 with its in-bench parity check.
set -u
OUT=gpurun_out
mkdir -p $OUT
T=${TAG:-r02g}
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee $OUT/${T}_gpus.log
echo "== tests/test_gpu_multi.py on 2 GPUs"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${T}_tests_multi.log
echo "== bench N=2"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 ) > $OUT/${T}_bench_n2.json 2> $OUT/${T}_bench_n2.err
tail -c 400 $OUT/${T}_bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02g_bench_n2.json').read().strip().splitlines()[-1])
print('value %.4e ms/step %.2f e2e %.4e (%.1f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
print('parity', d['parity_check'])
print('full', d['full_matrix'])
print('e2e', d['e2e'])
PY
echo "== reference arm under torchrun (rank 0 only; threads must equal the cores)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | tee $OUT/${T}_bench_reference_n2.json | cut -c1-700

#!/bin/bash
# 8-GPU call of round 2:  gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_call_r02h.sh'
# The bench line at N = 8 (device-resident full 200k matrix on 8 GPUs measured, in-bench parity check, e2e with
# the PCIe floor) and the A/B of the host-side knobs of the end-to-end path.
set -u
OUT=gpurun_out
mkdir -p $OUT
T=${TAG:-r02h}
nvidia-smi topo -m > $OUT/${T}_topo.log 2>&1; lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" > $OUT/${T}_lscpu.log
run() { # name, extra env, extra args
  name=$1; shift; envs=$1; shift
  ( time env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus 8 "$@" ) > $OUT/${T}_bench_$name.json 2> $OUT/${T}_bench_$name.err
  python - "$OUT/${T}_bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d['e2e']
    print(sys.argv[1], 'value %.4e ms/step %.2f | e2e %.4e %.1f ms floor %.1f ms (x%.2f) probe %.1f GB/s slowest, %.0f all | parity %s | bound cpus %s'
          % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['pcie_floor_ms'], e['ms_per_step_over_floor'],
             e['d2h_probe_gbs_slowest_rank'], e['d2h_probe_gbs_all_ranks'], d['parity_check'].get('ok'), d['config'].get('cpus_bound_rank0')))
    if d.get('full_matrix'):
        print('   full matrix', {k: d['full_matrix'][k] for k in ('t_build_s', 't_assemble_s', 't_gather_s', 't_total_s', 'nnz', 'mode')})
except Exception as ex:
    print(sys.argv[1], 'unreadable', ex)
PY
}
run default "X=1" --steps 10 --warmup 3
run nobind "FLUXB200_NO_BIND=1" --steps 6 --warmup 3 --no-sweep --no-full
run threads2 "X=1" --steps 6 --warmup 3 --no-sweep --no-full --option host_threads=2
run threads6 "X=1" --steps 6 --warmup 3 --no-sweep --no-full --option host_threads=6
run sub1024 "X=1" --steps 6 --warmup 3 --no-sweep --no-full --option sub_rows=1024
echo "== reference arm under torchrun"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | tail -1 | tee $OUT/${T}_bench_reference_n8.json | cut -c1-600
ls -la $OUT | tail -12

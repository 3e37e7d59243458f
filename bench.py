#!/usr/bin/env python
"""bench.py -- form-factor assembly throughput on B200 (contract: see DESIGN.md section 7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2]/[4], the config the metric's target is quoted
on): the synthetic Gaussian crater G(317, seed 0), 199 712 faces, float32.  The
full 4.0e10-pair CSR (about 160 GB) does not fit one GPU, so a STEP is one
contiguous slab of ROWS source faces x ALL 199 712 columns -- exactly the unit
of work a rank owns in the row-sharded 8-GPU assembly and exactly one
``get_form_factor_matrix(shape_model, I_slab)`` call of the reference API.
At N GPUs every rank assembles its own slab each step (weak scaling: per-GPU
work fixed) and fills its CSR slab; the row counts of all its slabs are
exchanged with one all-gather at the end of the timed region (the path's only
collective).

value  = visibility-tested pairs / s, whole job, CSR left resident in HBM
         (device timing, CUDA events on the library's stream, max over ranks)
e2e    = the same through ``fluxpy_b200.get_form_factor_matrix`` with host
         (NumPy) inputs and outputs: H2D of the index set and face arrays and
         D2H of the CSR arrays inside the timed region.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID_N = 317            # G(317, 0): 199 712 faces
EPS = 1e-5
METRIC = 'visibility-tested form-factor pairs/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=6)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--rows', type=int, default=4096, help='rows per slab (per rank, per step)')
    ap.add_argument('--grid', type=int, default=GRID_N, help='Gaussian-crater grid size n')
    ap.add_argument('--cpu-rows', type=int, default=0, help='rows of the CPU sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--option', action='append', default=[], metavar='NAME=VALUE',
                    help='library option (fluxb200_set_option), e.g. horizon_skip=1; recorded in config.options')
    return ap.parse_args()


def workload(args):
    from fluxpy_b200 import meshes
    V, F = meshes.gaussian_crater(args.grid, 0, dtype=np.float32)
    N = meshes.upward_normals(V, F)
    return V, F, N


def slab_rows(step, rank, world, rows, nf):
    """Rows of slab number step*world + rank (slabs tile the matrix top to bottom, wrapping)."""
    nslabs = max(1, nf//rows)          # equal slabs only (the remainder rows are not benched)
    # stride 7 (coprime to the 48 slabs of the 200k mesh): every run samples rim, wall and floor
    # rows alike, whatever the number of ranks -- slab cost varies by +-15 % across the crater
    s = ((step*world + rank)*7) % nslabs
    lo = s*rows
    return np.arange(lo, min(nf, lo + rows), dtype=np.int64)


def alg_flops(pairs_all, tested, nnz, nf):
    """22 flop per candidate pair, 50*ceil(log2 Nf)+50 per traced
    ray (one root-to-leaf descent + one triangle test), 6 per stored entry."""
    return 22.0*pairs_all + tested*(50.0*math.ceil(math.log2(nf)) + 50.0) + 6.0*nnz


def alg_bytes(nnz, m, nf, w=8):
    return nnz*w + 8.0*(m + 1) + 176.0*nf


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, indices):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = ','.join(str(i) for i in indices), [], False

    def run(self):
        try:
            p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                  '-lms', '100', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
        except OSError:
            return
        self.proc = p
        for line in p.stdout:
            self.rows.append([c.strip() for c in line.split(',')])
            if self.stop_flag:
                break
        p.terminate()

    def summary(self):
        self.stop_flag = True
        time.sleep(0.15)
        if getattr(self, 'proc', None):
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith('active'):
                        reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def measured_traffic():
    """DRAM bytes of one trace-kernel launch from the committed ncu capture
    (profiles/r01_trace_kernel_traffic.json), or None."""
    p = os.path.join(ROOT, 'profiles', 'r01_trace_kernel_traffic.json')
    if os.path.exists(p):
        return json.load(open(p))
    return None


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), float(d.get('sm_max_mhz', 1965.0)), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1965.0, 'fallback (B200_PROFILING.md)'


def cpu_port_sample(V, F, N, rows, nthreads=0):
    """The oracle port of the reference path (oracle/ff_oracle.c) on the host
    cores: `rows` x all columns.  Returns (tested pairs, seconds, threads)."""
    from oracle import oracle
    om = oracle.OracleShapeModel(V, F, N=N.copy(), nthreads=nthreads)
    t0 = time.perf_counter()
    _, st = oracle.get_form_factor_matrix(om, rows, None, EPS, return_stats=True)
    dt = time.perf_counter() - t0
    return st['pairs_tested'], st['pairs_all'], dt


def run_reference(args, rank, world):
    """--impl reference: the reference path's CPU implementation on the host
    cores.  /root/reference (Python + Embree) cannot travel to the GPU box and
    Embree is not installable offline, so this is the oracle PORT of that path
    (kind 'port'), all host threads, a bounded row sample per step."""
    if rank != 0:
        return
    V, F, N = workload(args)
    nf = F.shape[0]
    cores = len(os.sched_getaffinity(0))
    nrows = args.cpu_rows or 4*cores  # a multiple of the thread count: 24 rows on 16 threads left a quarter of them idle
    tested = pairs = 0
    times = []
    for s in range(args.warmup + args.steps):
        full = slab_rows(s, 0, 1, args.rows, nf)
        rows = full[np.linspace(0, len(full) - 1, min(nrows, len(full))).astype(int)]
        t, p, dt = cpu_port_sample(V, F, N, rows)
        if s >= args.warmup:
            tested += t
            pairs += p
            times.append(dt)
    total = sum(times)
    val = tested/total
    sample = f'{nrows} evenly spaced rows of each {args.rows}-row slab x all {nf} columns per step'
    out = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'pairs/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3*total/max(1, args.steps),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': f'G({args.grid},0) Gaussian crater, {nf} faces, float32; step = row sample x all columns',
                   'rows_per_step': nrows, 'eps': EPS},
        'pairs_all_per_s': pairs/total,
        'cpu_baseline': {'value': val, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(out), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import fluxpy_b200
    from fluxpy_b200 import form_factors, sharded

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'      # NCCL prints its version banner on stdout: one JSON line only
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus or world == 1, 'launch one rank per GPU (torchrun)'

    V, F, N = workload(args)
    nf = F.shape[0]
    fluxpy_b200.CudaTrimeshShapeModel.device = local_rank
    sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, N)
    options = {}
    for item in args.option:
        name, _, value = item.partition('=')
        sm.set_option(name, int(value))
        options[name] = int(value)
    stream = torch.cuda.ExternalStream(sm.cuda_stream(), device=dev)
    flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    info = sm.bvh_info()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident arm ---------------------------------------------------------
    def step_device(s):
        rows = slab_rows(s, rank, world, args.rows, nf)
        with torch.cuda.stream(stream):
            flush.zero_()                       # L2 flush between steps (in-stream, ~0.1 ms)
        m, n, counts, st = sm._ff_assemble_device(rows, None, EPS, 4, want_row_counts=world > 1)
        if world > 1:
            pending_counts.append(counts)       # exchanged once, at the end of the assembly (C1)
        return st, st

    pending_counts = []

    def exchange_pending():
        """C1, the path's only collective: every rank's row counts -> global indptr on every
        rank.  One all-gather per assembly (here: per timed region), as a rank that owns a
        contiguous set of rows and works through it slab by slab would do."""
        if world > 1 and pending_counts:
            mine = np.concatenate(pending_counts)
            starts = np.arange(world + 1, dtype=np.int64)*len(mine)
            sharded.exchange_row_counts(mine, starts, None, dev)
            pending_counts.clear()

    for s in range(args.warmup):
        step_device(s)
    exchange_pending()
    # one sampler for the whole box (rank 0 watches every GPU of the job): a sampler per rank
    # means N nvidia-smi processes taking the driver lock every 100 ms inside the timed region
    sampler = ClockSampler(range(world) if world > 1 else [local_rank]) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    acc = {'tested': 0, 'pairs': 0, 'nnz': 0, 'trace_ms': 0.0, 'fill_ms': 0.0, 'launches': 0, 'rows': 0,
           'trace_launches': 0}
    for s in range(args.warmup, args.warmup + args.steps):
        st, st2 = step_device(s)
        acc['tested'] += st.pairs_tested
        acc['pairs'] += st.pairs_all
        acc['nnz'] += st.nnz
        acc['trace_ms'] += st.ms_trace
        acc['fill_ms'] += st2.ms_fill
        acc['trace_launches'] += st.trace_launches
        acc['launches'] += st2.kernel_launches + 1      # + the L2-flush memset
        acc['rows'] += len(slab_rows(s, rank, world, args.rows, nf))
    exchange_pending()
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    clocks = sampler.summary() if sampler else None

    # ---- end-to-end arm: public API, host buffers in and out -----------------------------
    def step_e2e(s):
        rows = slab_rows(s, rank, world, args.rows, nf)
        if world > 1:
            res = sharded.get_form_factor_matrix_sharded(sm, np.concatenate(
                [slab_rows(s, r, world, args.rows, nf) for r in range(world)]), None, EPS)
            FF = res.local_csr
        else:
            FF = fluxpy_b200.get_form_factor_matrix(sm, rows, None, EPS)
        st = form_factors.last_stats if world == 1 else res.stats
        # bytes the library actually moved (its own count): the CSR values + the visibility words
        # the column indices are expanded from on the host + row counts; index set + face arrays in
        d2h = st['d2h_bytes']
        h2d = st['h2d_bytes'] + sm.P.nbytes + sm.N.nbytes + sm.A.nbytes
        chk = float(FF.data[:16].sum())                  # touch the result on the host
        return st, h2d, d2h, chk

    for s in range(args.warmup):
        step_e2e(s)
    barrier()
    t0 = time.perf_counter()
    e2e = {'tested': 0, 'h2d': 0, 'd2h': 0, 'step_ms': []}
    for s in range(args.warmup, args.warmup + args.steps):
        ts = time.perf_counter()
        st, h2d, d2h, _ = step_e2e(s)
        e2e['step_ms'].append(round(1e3*(time.perf_counter() - ts), 2))
        e2e['tested'] += st['pairs_tested']
        e2e['h2d'] += h2d
        e2e['d2h'] += d2h
    barrier()
    t_e2e = time.perf_counter() - t0

    # ---- reduce over ranks: max time, summed work --------------------------------------
    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_dev_max = allmax(ms_dev)
    t_e2e_max = allmax(t_e2e)
    tested_all, pairs_all, nnz_all = allsum(acc['tested']), allsum(acc['pairs']), allsum(acc['nnz'])
    e2e_tested_all = allsum(e2e['tested'])
    launches_all = allsum(acc['launches'])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        full = slab_rows(args.warmup, 0, 1, args.rows, nf)
        probe = full[np.linspace(0, len(full) - 1, 2*cores).astype(int)]
        t, p, dt = cpu_port_sample(V, F, N, probe)
        nrows = args.cpu_rows or int(min(len(full), max(2*cores, round(15.0/(dt/len(probe))))))
        rows = full[np.linspace(0, len(full) - 1, nrows).astype(int)]
        t, p, dt = cpu_port_sample(V, F, N, rows)
        cpu = {'value': t/dt, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
               'sample': f'{nrows} evenly spaced rows of one {len(full)}-row slab x all {nf} columns '
                         f'({t} rays, {dt:.1f} s, OpenMP over rows)',
               'pairs_all_per_s': p/dt}

    if rank == 0:
        hbm_peak, sm_max_mhz, peak_src = measured_peaks()
        props = torch.cuda.get_device_properties(dev)
        fp32_peak = props.multi_processor_count*128*2*sm_max_mhz*1e6/1e12       # TFLOP/s
        # dominant kernel = trace_kernel (one launch per step per rank); per-launch figures of rank 0
        steps = args.steps
        nl = max(1, acc['trace_launches'])
        fl = alg_flops(acc['pairs'], acc['tested'], 0, nf)/nl                      # per trace launch
        trace_s = acc['trace_ms']/nl/1e3
        by = alg_bytes(acc['nnz']/steps, acc['rows']/steps, nf)
        assemble_s = ms_dev/steps/1e3
        roof = {
            'kernel': 'trace_kernel<float> (fused cull + occlusion traversal)',
            'bound': 'fp32', 'achieved': fl/trace_s/1e12, 'peak': fp32_peak, 'unit': 'TFLOP/s',
            'frac': fl/trace_s/1e12/fp32_peak, 'traffic': (measured_traffic() or {}).get('dram_bytes_per_launch'),
            'traffic_note': (measured_traffic() or {}).get('note'), 'peak_source': 'SMs*128*2*sm_max_mhz',
            'alg_flop_per_launch': fl, 'launch_ms': 1e3*trace_s, 'launches_per_step': nl/steps,
            'hbm': {'bound': 'hbm', 'achieved': by/assemble_s/1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': by/assemble_s/1e9/hbm_peak, 'alg_bytes_per_step': by, 'peak_source': peak_src},
            'trace_share_of_step': acc['trace_ms']/ms_dev,
        }
        full_est = nf/args.rows*(ms_dev_max/steps)/1e3/world
        out = {
            'metric': METRIC, 'value': tested_all/(ms_dev_max/1e3), 'unit': 'pairs/s',
            'n_gpus': world, 'steps': steps, 'warmup': args.warmup, 'ms_per_step': ms_dev_max/steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': f'G({args.grid},0) Gaussian crater, {nf} faces, float32, eps=1e-5; '
                                   f'step = {args.rows}-row slab x all {nf} columns per GPU '
                                   f'(the 8-GPU row-sharded unit of the 200k-face config)',
                       'rows_per_step_per_gpu': args.rows, 'parallelism': f'row-slabs x{world}',
                       'l2': 'explicit 256 MB flush between steps + each step streams >3 GB of CSR output',
                       'bvh': {'nodes': info.num_nodes, 'top_nodes_smem': info.num_top_nodes,
                               'depth': info.max_depth, 'build_ms': info.ms_build},
                       'options': options, 'trace_counters': sm.trace_counters()},
            'pairs_all_per_s': pairs_all/(ms_dev_max/1e3),
            'nnz_per_step': nnz_all/steps,
            'csr_assembly_s_full_matrix_est': full_est,
            'clocks': clocks,
            'e2e': {'value': e2e_tested_all/t_e2e_max, 'unit': 'pairs/s',
                    'h2d_bytes_per_step': e2e['h2d']/steps, 'd2h_bytes_per_step': e2e['d2h']/steps,
                    'ms_per_step': 1e3*t_e2e_max/steps, 'step_ms_rank0': e2e['step_ms'],
                    'output_buffer_retries': fluxpy_b200.CudaTrimeshShapeModel.overflow_retries},
            'gpu_launches': int(launches_all),
            'roofline': roof,
            'cpu_baseline': cpu,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

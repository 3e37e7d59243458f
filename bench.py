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
    ap.add_argument('--no-sweep', action='store_true', help='skip the mesh-size / fp64 arms (BASELINE config 5)')
    ap.add_argument('--no-full', action='store_true', help='skip the measured full-matrix assembly')
    ap.add_argument('--full-host-grid', type=int, default=0,
                    help='also assemble the FULL matrix of G(n,0) into a host SciPy CSR (n = 159: 10 GB)')
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


def source_sha16():
    """Hash of the CUDA sources the loaded library was built from (the build is in-tree and
    __graft_entry__.build() rebuilds whenever a source is newer than the .so)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'fluxpy_b200', 'csrc')
    for f in sorted(os.listdir(d)):
        if f.endswith(('.cu', '.cuh', '.cpp', '.h')):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), 'rb').read())
    return h.hexdigest()[:16]


def measured_traffic():
    """DRAM bytes of one trace-kernel launch from the newest committed `ncu --set full` capture
    (profiles/*_trace_kernel_traffic.json, written by tools/ncu_traffic.py with the hash of the
    sources it profiled).  A capture of other sources is not this build's traffic: None."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_trace_kernel_traffic.json')))
    if not files:
        return None
    d = json.load(open(files[-1]))
    if d.get('source_sha16') != source_sha16():
        return {'dram_bytes_per_launch': None,
                'note': f'{os.path.basename(files[-1])} is a capture of sources {d.get("source_sha16")}, '
                        f'this build is {source_sha16()}: no traffic figure for this build'}
    return d


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), float(d.get('sm_max_mhz', 1965.0)), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1965.0, 'fallback (B200_PROFILING.md)'


def host_threads():
    """Threads the CPU arm uses = the cores this process may run on.  Passed to OpenMP
    explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, which made a default-sized
    team one thread in round 1 while the line said 32."""
    return len(os.sched_getaffinity(0))


def cpu_port_sample(V, F, N, rows, nthreads):
    """The oracle port of the reference path (oracle/ff_oracle.c) on the host
    cores: `rows` x all columns.  Returns (tested pairs, all pairs, seconds)."""
    from oracle import oracle
    got = oracle.team_size(nthreads)
    assert got == nthreads, f'OpenMP gave {got} threads, asked for {nthreads}'
    om = oracle.OracleShapeModel(V, F, N=N.copy(), nthreads=nthreads)
    t0 = time.perf_counter()
    _, st = oracle.get_form_factor_matrix(om, rows, None, EPS, return_stats=True)
    dt = time.perf_counter() - t0
    return st['pairs_tested'], st['pairs_all'], dt


def csr_digest(FF):
    """sha256 over indptr / indices / data of a SciPy CSR (index arrays as int64): equal digests = equal arrays."""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(FF.indptr, np.int64).tobytes())
    h.update(np.ascontiguousarray(FF.indices, np.int64).tobytes())
    h.update(np.ascontiguousarray(FF.data).tobytes())
    return np.frombuffer(h.digest(), np.uint8).copy()


PARITY_ROWS = 64


def parity_check(ctx, last_rows_of, last_slab):
    """Driver-run proof that the N-GPU path computes what one GPU computes (rows are independent,
    reference src/flux/form_factors.py:45-70).  Every rank re-assembles PARITY_ROWS sampled rows of
    the last slab it timed; rank 0 assembles the same rows OF EVERY RANK on its own GPU and the
    digests of indptr / indices / data must agree.  On every rank the sampled block must also equal
    the same rows sliced out of the slab it produced in the end-to-end arm, and the global indptr
    of the all-gather must equal the concatenated local row counts."""
    import fluxpy_b200
    torch, dist, sm, rank, world, dev = ctx['torch'], ctx['dist'], ctx['sm'], ctx['rank'], ctx['world'], ctx['dev']

    def sample(r):
        rows = last_rows_of(r)
        pick = np.linspace(0, len(rows) - 1, min(PARITY_ROWS, len(rows))).astype(int)
        return rows[pick], pick

    mine, pick = sample(rank)
    FFm = fluxpy_b200.get_form_factor_matrix(sm, mine, None, EPS)
    slab_ok = True
    if last_slab is not None:
        sl = last_slab['csr'][pick]
        sl.sort_indices()
        slab_ok = bool(np.array_equal(sl.indptr, FFm.indptr) and np.array_equal(sl.indices, FFm.indices)
                       and np.array_equal(sl.data, FFm.data))
    indptr_ok = True
    if last_slab is not None and last_slab.get('global_indptr') is not None:
        g, lo, hi = last_slab['global_indptr'], last_slab['row_start'], last_slab['row_stop']
        local = np.diff(np.asarray(last_slab['csr'].indptr, np.int64))
        indptr_ok = bool(np.array_equal(np.diff(g)[lo:hi], local))
        tot = torch.tensor([float(last_slab['csr'].nnz)], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        indptr_ok = indptr_ok and int(tot.item()) == int(g[-1])
    dig = torch.as_tensor(csr_digest(FFm), device=dev)
    flags = torch.tensor([1.0 if slab_ok else 0.0, 1.0 if indptr_ok else 0.0], dtype=torch.float64, device=dev)
    cross_ok = True
    if world > 1:
        allg = torch.empty(world*32, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allg, dig)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            allg = allg.cpu().numpy().reshape(world, 32)
            for r in range(world):
                FFr = fluxpy_b200.get_form_factor_matrix(sm, sample(r)[0], None, EPS)
                cross_ok = cross_ok and bool(np.array_equal(csr_digest(FFr), allg[r]))
    slab_ok, indptr_ok = bool(flags[0].item() > 0.5), bool(flags[1].item() > 0.5)
    return {'rows_per_rank': int(len(mine)), 'ranks': world,
            'every_rank_equals_rank0_single_gpu': cross_ok, 'block_equals_slab_slice': slab_ok,
            'global_indptr_equals_local_counts': indptr_ok, 'ok': bool(cross_ok and slab_ok and indptr_ok)}


def pcie_probe(ctx, nbytes=1 << 29, reps=4):
    """Page-locked device-to-host copy bandwidth of every rank WITH ALL RANKS COPYING AT ONCE: what the box
    gives the end-to-end arm's copy-out (one GPU alone gets 50+ GB/s; eight at once shared 91-146 GB/s of
    host-side bandwidth in round 1).  e2e.pcie_floor_ms = d2h bytes per step / this."""
    torch, dev = ctx['torch'], ctx['dev']
    x = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h.copy_(x, non_blocking=True)
    best = 0.0
    for _ in range(2):
        ctx['barrier']()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            h.copy_(x, non_blocking=True)
        e1.record()
        torch.cuda.synchronize(dev)
        best = max(best, reps*nbytes/(e0.elapsed_time(e1)/1e3)/1e9)
    ctx['barrier']()
    del x, h
    return best


def sweep_arm(ctx, grid, dtype, rows_want, steps=3, warmup=2):
    """One short device-resident arm on another mesh size / dtype (BASELINE config 5; the reference's own
    methodology sweeps the mesh size: examples/spherical_crater/run_example.sh:3-4,19)."""
    import fluxpy_b200
    from fluxpy_b200 import meshes
    torch, rank, world, dev = ctx['torch'], ctx['rank'], ctx['world'], ctx['dev']
    V, F = meshes.gaussian_crater(grid, 0, dtype=dtype)
    N = meshes.upward_normals(V, F)
    sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, N)
    for name, value in ctx['options'].items():
        sm.set_option(name, value)
    nf = F.shape[0]
    rows = min(rows_want, nf)
    stream = torch.cuda.ExternalStream(sm.cuda_stream(), device=dev)
    acc = {'tested': 0, 'pairs': 0, 'trace_ms': 0.0, 'fill_ms': 0.0, 'prepare_ms': 0.0, 'launches': 0, 'trace_launches': 0}

    def step(s, timed):
        I = slab_rows(s, rank, world, rows, nf)
        with torch.cuda.stream(stream):
            ctx['flush'].zero_()
        m, n, _, st = sm._ff_assemble_device(I, None, EPS, 4)
        if timed:
            acc['tested'] += st.pairs_tested
            acc['pairs'] += st.pairs_all
            acc['trace_ms'] += st.ms_trace
            acc['fill_ms'] += st.ms_fill
            acc['prepare_ms'] += st.ms_prepare
            acc['trace_launches'] += st.trace_launches
            acc['launches'] += st.kernel_launches + 1

    for s in range(warmup):
        step(s, False)
    ctx['barrier']()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    wall = []
    for s in range(warmup, warmup + steps):
        t = time.perf_counter()
        step(s, True)
        wall.append(round(1e3*(time.perf_counter() - t), 2))
    e1.record(stream)
    ctx['barrier']()
    ms = ctx['allmax'](e0.elapsed_time(e1))
    tested, pairs = ctx['allsum'](acc['tested']), ctx['allsum'](acc['pairs'])
    peak = ctx['fp32_peak']*(0.5 if dtype == np.float64 else 1.0)
    fl = alg_flops(acc['pairs'], acc['tested'], 0, nf)/max(1, acc['trace_launches'])
    trace_s = acc['trace_ms']/max(1, acc['trace_launches'])/1e3
    ctx['launches_extra'] += ctx['allsum'](acc['launches'])
    del sm
    return {'faces': int(nf), 'grid': grid, 'dtype': 'f64' if dtype == np.float64 else 'f32', 'rows_per_step_per_gpu': rows,
            'steps': steps, 'pairs_per_s': tested/(ms/1e3), 'pairs_all_per_s': pairs/(ms/1e3), 'ms_per_step': ms/steps,
            'trace_ms_per_launch': 1e3*trace_s, 'fill_ms_per_step': acc['fill_ms']/steps,
            'prepare_ms_per_step': acc['prepare_ms']/steps, 'step_ms_wall_rank0': wall, 'roofline_frac': fl/trace_s/1e12/peak,
            'roofline_peak_tflops': peak}


def full_matrix(ctx, V, F, N, nf):
    """The metric's second half, measured not extrapolated: the FULL nf x nf CSR assembled by the job's N
    GPUs (reference methodology: examples/gerlache/make_true_form_factor_matrix.py:27-32 times the whole
    matrix).  Rows are cut into N contiguous slabs (fluxb200_slab_plan); every rank builds its own replica of
    the mesh + LBVH + horizons and assembles its slab device-resident; the row counts are all-gathered into
    the global indptr.  One GPU cannot hold the 200k-face CSR (about 148 GB): at N = 1 the slab is streamed
    through the same device buffers 4096 rows at a time (all kernels run, the entries are overwritten)."""
    import ctypes
    import fluxpy_b200
    from fluxpy_b200 import _lib, sharded
    from fluxpy_b200.device_csr import DeviceCsrSlab
    torch, dist, rank, world, dev = ctx['torch'], ctx['dist'], ctx['rank'], ctx['world'], ctx['dev']
    ctx['barrier']()
    t0 = time.perf_counter()
    sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, N)
    for name, value in ctx['options'].items():
        sm.set_option(name, value)
    torch.cuda.synchronize(dev)
    t_build = time.perf_counter() - t0
    starts = sharded.slab_bounds(nf, world)
    lo, hi = int(starts[rank]), int(starts[rank + 1])
    stream = torch.cuda.ExternalStream(sm.cuda_stream(), device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx['barrier']()
    t1 = time.perf_counter()
    e0.record(stream)
    nnz = tested = launches = 0
    trace_ms = 0.0
    slab = None
    if world == 1:
        counts = []
        for r0 in range(lo, hi, 4096):
            m, n, cnt, st = sm._ff_assemble_device(np.arange(r0, min(hi, r0 + 4096), dtype=np.int64), None, EPS, 4,
                                                   want_row_counts=True)
            counts.append(cnt)
            nnz += st.nnz
            tested += st.pairs_tested
            trace_ms += st.ms_trace
            launches += st.kernel_launches
        counts = np.concatenate(counts)
        mode = 'streamed: 4096-row slabs through one set of device buffers (the CSR does not fit one GPU)'
    else:
        m, n, counts, st = sm._ff_assemble_device(np.arange(lo, hi, dtype=np.int64), None, EPS, 4, want_row_counts=True)
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().fluxb200_ff_detach_csr(sm._handle, ctypes.byref(h)))
        slab = DeviceCsrSlab(h, sm.device, lo, nf)
        nnz, tested, trace_ms, launches = st.nnz, st.pairs_tested, st.ms_trace, st.kernel_launches
        mode = 'device-resident: every rank keeps its row slab of the CSR in HBM'
    e1.record(stream)
    torch.cuda.synchronize(dev)
    t_assemble_dev = e0.elapsed_time(e1)/1e3
    t2 = time.perf_counter()
    if world > 1:
        indptr = sharded.exchange_row_counts(counts, starts, None, dev)
    else:
        indptr = np.zeros(nf + 1, np.int64)
        np.cumsum(counts, out=indptr[1:])
    torch.cuda.synchronize(dev)
    t3 = time.perf_counter()
    nnz_all = ctx['allsum'](nnz)
    ok = int(indptr[-1]) == int(nnz_all)
    out = {'faces': int(nf), 'rows': int(nf), 'n_gpus': world, 'mode': mode, 'nnz': int(nnz_all),
           'pairs_tested': int(ctx['allsum'](tested)), 'csr_bytes': int(nnz_all)*8 + 8*(nf + 1),
           't_build_s': ctx['allmax'](t_build), 't_assemble_s': ctx['allmax'](t_assemble_dev),
           't_assemble_wall_s': ctx['allmax'](t2 - t1), 't_gather_s': ctx['allmax'](t3 - t2),
           't_total_s': ctx['allmax'](t3 - t0), 'trace_s_rank0': trace_ms/1e3,
           'indptr_last_equals_nnz': ok,
           'note': 't_build = mesh upload + face geometry + LBVH (+ zone table); the per-face horizons (once per mesh) '
                   'fall into t_assemble; t_gather = all-gather of the row counts + prefix sum -> global indptr'}
    ctx['launches_extra'] += ctx['allsum'](launches)
    del slab, sm
    return out


def full_host_csr(ctx, grid):
    """The full matrix of a mesh whose CSR fits host memory, through the public API into a SciPy CSR
    (H2D of the mesh, D2H of the CSR inside the time)."""
    import fluxpy_b200
    from fluxpy_b200 import meshes, form_factors
    V, F = meshes.gaussian_crater(grid, 0, dtype=np.float32)
    N = meshes.upward_normals(V, F)
    t0 = time.perf_counter()
    sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, N)
    for name, value in ctx['options'].items():
        sm.set_option(name, value)
    t1 = time.perf_counter()
    FF = fluxpy_b200.get_form_factor_matrix(sm, None, None, EPS)
    t2 = time.perf_counter()
    st = form_factors.last_stats
    out = {'faces': int(F.shape[0]), 'nnz': int(FF.nnz), 'csr_bytes': int(FF.data.nbytes + FF.indices.nbytes + FF.indptr.nbytes),
           't_build_s': t1 - t0, 't_assemble_to_host_s': t2 - t1, 'pairs_per_s': st['pairs_tested']/(t2 - t1),
           'index_dtype': str(FF.indices.dtype), 'first_call': 'includes page-locking the output buffers'}
    t3 = time.perf_counter()
    FF2 = fluxpy_b200.get_form_factor_matrix(sm, None, None, EPS)
    out['t_assemble_to_host_s_second_call'] = time.perf_counter() - t3
    out['same_result'] = bool(FF2.nnz == FF.nnz and np.array_equal(FF2.indptr, FF.indptr))
    del FF, FF2, sm
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference path's CPU implementation on the host
    cores.  /root/reference (Python + Embree) cannot travel to the GPU box and
    Embree is not installable offline, so this is the oracle PORT of that path
    (kind 'port'), all host threads, a bounded row sample per step."""
    if rank != 0:
        return
    V, F, N = workload(args)
    nf = F.shape[0]
    cores = host_threads()
    nrows = args.cpu_rows or 4*cores  # a multiple of the thread count: 24 rows on 16 threads left a quarter of them idle
    tested = pairs = 0
    times = []
    for s in range(args.warmup + args.steps):
        full = slab_rows(s, 0, 1, args.rows, nf)
        rows = full[np.linspace(0, len(full) - 1, min(nrows, len(full))).astype(int)]
        t, p, dt = cpu_port_sample(V, F, N, rows, cores)
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
        'cpu_baseline': {'value': val, 'unit': 'pairs/s', 'cores': cores, 'omp_threads_used': cores, 'kind': 'port',
                         'sample': sample},
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

    bound_cpus = None
    if world > 1 and os.environ.get('FLUXB200_NO_BIND') != '1':
        bound_cpus = sharded.bind_to_gpu_cpus(local_rank)     # NUMA-local threads and page-locked memory
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
    trace_counters = sm.trace_counters()        # of the last timed step's launch

    # ---- end-to-end arm: public API, host buffers in and out -----------------------------
    last_slab = {}

    def step_e2e(s):
        rows = slab_rows(s, rank, world, args.rows, nf)
        if world > 1:
            res = sharded.get_form_factor_matrix_sharded(sm, np.concatenate(
                [slab_rows(s, r, world, args.rows, nf) for r in range(world)]), None, EPS)
            FF = res.local_csr
            last_slab.update(csr=FF, global_indptr=res.global_indptr, row_start=res.row_start, row_stop=res.row_stop)
        else:
            FF = fluxpy_b200.get_form_factor_matrix(sm, rows, None, EPS)
            last_slab.update(csr=FF, global_indptr=None)
        st = form_factors.last_stats if world == 1 else res.stats
        # bytes the library actually moved (its own count): the CSR values + the visibility words
        # the column indices are expanded from on the host + row counts; index set + face arrays in
        d2h = st['d2h_bytes']
        # (the face arrays P, N, A are compared with what the device holds and re-sent only when they
        # changed -- they are the shape model's state, like the reference's Embree scene; the per-call
        # inputs are the index sets)
        h2d = st['h2d_bytes'] + sm.face_bytes_sent_last
        chk = float(FF.data[:16].sum())                  # touch the result on the host
        return st, h2d, d2h, chk

    for s in range(args.warmup):
        step_e2e(s)
    barrier()
    t0 = time.perf_counter()
    e2e = {'tested': 0, 'h2d': 0, 'd2h': 0, 'idx_bytes': 0, 'step_ms': []}
    for s in range(args.warmup, args.warmup + args.steps):
        ts = time.perf_counter()
        st, h2d, d2h, _ = step_e2e(s)
        e2e['step_ms'].append(round(1e3*(time.perf_counter() - ts), 2))
        e2e['tested'] += st['pairs_tested']
        e2e['h2d'] += h2d
        e2e['d2h'] += d2h
        e2e['idx_bytes'] += 4*st['nnz']                 # int32 column indices written by the host threads
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
    props = torch.cuda.get_device_properties(dev)
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    fp32_peak = props.multi_processor_count*128*2*sm_max_mhz*1e6/1e12       # TFLOP/s
    ctx = {'torch': torch, 'dist': dist, 'sm': sm, 'rank': rank, 'world': world, 'dev': dev, 'flush': flush,
           'barrier': barrier, 'allmax': allmax, 'allsum': allsum, 'options': options, 'fp32_peak': fp32_peak,
           'launches_extra': 0}
    d2h_gbs = pcie_probe(ctx)
    d2h_gbs_min = -allmax(-d2h_gbs)
    d2h_gbs_sum = allsum(d2h_gbs)
    host_bytes_all = allsum(e2e['d2h'] + e2e['idx_bytes'])   # everything the step writes into host DRAM, all ranks
    # ---- after the timed regions: N-GPU result == one-GPU result, on the box the driver runs --------------
    s_last = args.warmup + args.steps - 1
    parity = parity_check(ctx, lambda r: slab_rows(s_last, r, world, args.rows, nf), last_slab or None)
    last_slab.clear()
    # ---- other mesh sizes and the fp64 mode (BASELINE config 5), short arms ----------------------------------
    sweep = []
    if not args.no_sweep:
        for grid, dt in ((72, np.float32), (159, np.float32), (501, np.float32), (159, np.float64)):
            if grid == args.grid and dt == np.float32:
                continue
            sweep.append(sweep_arm(ctx, grid, dt, args.rows))
    # ---- the full matrix, measured ---------------------------------------------------------------------------
    full = None if args.no_full else full_matrix(ctx, V, F, N, nf)
    full_host = full_host_csr(ctx, args.full_host_grid) if (args.full_host_grid and world == 1) else None
    tested_all, pairs_all, nnz_all = allsum(acc['tested']), allsum(acc['pairs']), allsum(acc['nnz'])
    e2e_tested_all = allsum(e2e['tested'])
    launches_all = allsum(acc['launches'])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        slab0 = slab_rows(args.warmup, 0, 1, args.rows, nf)
        probe = slab0[np.linspace(0, len(slab0) - 1, 2*cores).astype(int)]
        t, p, dt = cpu_port_sample(V, F, N, probe, cores)
        nrows = args.cpu_rows or int(min(len(slab0), max(2*cores, round(15.0/(dt/len(probe))))))
        rows = slab0[np.linspace(0, len(slab0) - 1, nrows).astype(int)]
        t, p, dt = cpu_port_sample(V, F, N, rows, cores)
        cpu = {'value': t/dt, 'unit': 'pairs/s', 'cores': cores, 'omp_threads_used': cores, 'kind': 'port',
               'sample': f'{nrows} evenly spaced rows of one {len(slab0)}-row slab x all {nf} columns '
                         f'({t} rays, {dt:.1f} s, OpenMP over rows)',
               'pairs_all_per_s': p/dt}

    if rank == 0:
        # dominant kernel = trace_kernel (one launch per step per rank); per-launch figures of rank 0
        steps = args.steps
        nl = max(1, acc['trace_launches'])
        fl = alg_flops(acc['pairs'], acc['tested'], 0, nf)/nl                      # per trace launch
        trace_s = acc['trace_ms']/nl/1e3
        by = alg_bytes(acc['nnz']/steps, acc['rows']/steps, nf)
        assemble_s = ms_dev/steps/1e3
        roof = {
            'kernel': ('trace_kernel' if options.get('trace_variant') == 1 else 'trace2_kernel') + '<float> (fused cull + occlusion traversal)',
            'bound': 'fp32', 'achieved': fl/trace_s/1e12, 'peak': fp32_peak, 'unit': 'TFLOP/s',
            'frac': fl/trace_s/1e12/fp32_peak, 'traffic': (measured_traffic() or {}).get('dram_bytes_per_launch'),
            'traffic_note': (measured_traffic() or {}).get('note'), 'peak_source': 'SMs*128*2*sm_max_mhz',
            'alg_flop_per_launch': fl, 'launch_ms': 1e3*trace_s, 'launches_per_step': nl/steps,
            'hbm': {'bound': 'hbm', 'achieved': by/assemble_s/1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': by/assemble_s/1e9/hbm_peak, 'alg_bytes_per_step': by, 'peak_source': peak_src},
            'trace_share_of_step': acc['trace_ms']/ms_dev,
        }
        full_est = nf/args.rows*(ms_dev_max/steps)/1e3/world
        sweep.insert(min(2, len(sweep)), {
            'faces': int(nf), 'grid': args.grid, 'dtype': 'f32', 'rows_per_step_per_gpu': args.rows, 'steps': steps,
            'pairs_per_s': tested_all/(ms_dev_max/1e3), 'pairs_all_per_s': pairs_all/(ms_dev_max/1e3),
            'ms_per_step': ms_dev_max/steps, 'trace_ms_per_launch': 1e3*trace_s, 'roofline_frac': roof['frac'],
            'roofline_peak_tflops': fp32_peak, 'note': 'the headline arm'})
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
                       'options': options, 'trace_counters': trace_counters,
                       'cpus_bound_rank0': (len(bound_cpus) if bound_cpus else None)},
            'pairs_all_per_s': pairs_all/(ms_dev_max/1e3),
            'nnz_per_step': nnz_all/steps,
            'csr_assembly_s_full_matrix': (full or {}).get('t_assemble_s'),
            'csr_assembly_s_full_matrix_extrapolated_from_steps': full_est,
            'full_matrix': full,
            'full_matrix_host_csr': full_host,
            'sweep': sweep,
            'parity_check': parity,
            'source_sha16': source_sha16(),
            'clocks': clocks,
            'e2e': {'value': e2e_tested_all/t_e2e_max, 'unit': 'pairs/s',
                    'h2d_bytes_per_step': e2e['h2d']/steps, 'd2h_bytes_per_step': e2e['d2h']/steps,
                    'ms_per_step': 1e3*t_e2e_max/steps, 'step_ms_rank0': e2e['step_ms'],
                    'd2h_probe_gbs_slowest_rank': d2h_gbs_min, 'd2h_probe_gbs_all_ranks': d2h_gbs_sum,
                    'pcie_floor_ms': e2e['d2h']/steps/(d2h_gbs_min*1e9)*1e3,
                    'ms_per_step_over_floor': (1e3*t_e2e_max/steps)/(e2e['d2h']/steps/(d2h_gbs_min*1e9)*1e3),
                    # the box's host side moves a fixed total (the probe's sum over ranks) whoever writes: the DMA of
                    # values + visibility words AND the host threads' index stores share it
                    'host_write_bytes_per_step_all_ranks': host_bytes_all/steps,
                    'host_memory_floor_ms': host_bytes_all/steps/(d2h_gbs_sum*1e9)*1e3,
                    'ms_per_step_over_host_memory_floor': (1e3*t_e2e_max/steps)/(host_bytes_all/steps/(d2h_gbs_sum*1e9)*1e3),
                    'step_ms_median_rank0': float(np.median(e2e['step_ms'])),
                    'bound': 'host side of the copy-out (PCIe DMA + index expansion both write host DRAM)'
                             if (1e3*t_e2e_max/steps) > 1.5*ms_dev_max/steps else 'trace kernel (copy-out hidden under it)',
                    'output_buffer_retries': fluxpy_b200.CudaTrimeshShapeModel.overflow_retries},
            'gpu_launches': int(launches_all),
            'gpu_launches_outside_timed_region': int(ctx['launches_extra']),
            'roofline': roof,
            'cpu_baseline': cpu,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

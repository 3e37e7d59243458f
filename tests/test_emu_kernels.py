"""CPU tier: the library's own CUDA sources executed on the SIMT emulator (tools/simt).

`tools/simt/build_emu.py` compiles fluxpy_b200/csrc/*.cu(h) for the host -- every CUDA thread a fiber,
warp collectives and __syncthreads with CUDA's semantics, a synchronous stand-in for the runtime API --
into `libfluxb200_emu.so` with the same C ABI.  `tests/conftest.py` swaps it in when
FLUXB200_TEST_EMU=1, so the gpu-marked parity tests (bit-for-bit against the oracle and the golden
vectors) run here unchanged, on the kernels' real source.  This is test infrastructure: the package has
no switch for it and never loads it; what it proves is the kernels' logic and warp synchronisation, not
their speed, and not anything that depends on the hardware's memory model.

The default path is the second-generation trace kernel (csrc/trace2.cuh: warp-shared traversal queue) with the
horizon skip on.  The second run repeats the subset with a small zone size (FLUXB200_TEST_HORIZON=16: both ends
of the skip exercised on these small meshes), the third with the first-generation kernel and the skip off.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the quick part of the gpu tier (seconds each on the emulator); the rest runs there with
#   FLUXB200_TEST_EMU=1 python -m pytest tests -m gpu      (conftest skips what needs a real device or is too large)
SUBSET = ('test_sphere_fixtures or test_sphere_shape_api or (test_crater_vs_oracle_and_golden and not case2) '
          'or test_steady_state_temperature or test_edge_cases or test_fill_kernel_variants_agree '
          'or test_coincident_faces_tie_rule or test_ingersoll_bowl or test_random_triangle_soup '
          # independent ground truths: exact convex clipping, height-field clearance, analytic Ingersoll crater
          'or test_is_occluded_equals_exact_convex_answer '
          'or (test_visibility_and_csr_pattern_equal_heightfield_geometry and case0) '
          'or test_ingersoll_analytic_flux_and_temperature')


@pytest.fixture(scope='module')
def emu_lib():
    sys.path.insert(0, os.path.join(ROOT, 'tools', 'simt'))
    import build_emu
    return build_emu.build()


def run_gpu_subset(extra_env):
    env = dict(os.environ, FLUXB200_TEST_EMU='1', **extra_env)
    out = subprocess.run([sys.executable, '-m', 'pytest', 'tests/test_gpu_parity.py', 'tests/test_gpu_meshes.py',
                          '-m', 'gpu', '-q', '-x', '-k', SUBSET, '-p', 'no:cacheprovider'],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = '\n'.join(out.stdout.splitlines()[-15:]) + out.stderr[-2000:]
    assert out.returncode == 0, tail
    assert ' passed' in out.stdout and 'failed' not in out.stdout, tail
    return out.stdout


def test_emulated_library_exports_the_abi(emu_lib):
    import ctypes
    from fluxpy_b200 import _lib
    L = ctypes.CDLL(emu_lib)
    missing = [s for s in _lib.EXPORTS if not hasattr(L, s)]
    assert not missing
    assert L.fluxb200_abi_version() == _lib.ABI_VERSION


def test_gpu_parity_subset_on_the_emulator(emu_lib):
    run_gpu_subset({})


def test_gpu_parity_subset_on_the_emulator_with_horizon_skip(emu_lib):
    run_gpu_subset({'FLUXB200_TEST_HORIZON': '16'})


def test_gpu_parity_subset_on_the_emulator_first_generation_kernel(emu_lib):
    """The quick crater / sphere cases with the first-generation trace kernel (per-lane stacks) and the skip
    off: the A/B reference of trace2.cuh stays bit-identical to the oracle too."""
    env = dict(os.environ, FLUXB200_TEST_EMU='1', FLUXB200_TEST_VARIANT='1', FLUXB200_TEST_HORIZON='0')
    out = subprocess.run([sys.executable, '-m', 'pytest', 'tests/test_gpu_parity.py', '-m', 'gpu', '-q', '-x', '-k',
                          'test_sphere_fixtures or (test_crater_vs_oracle_and_golden and not case2) or test_edge_cases',
                          '-p', 'no:cacheprovider'], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0 and ' passed' in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]


def test_host_pipeline_on_the_emulator(emu_lib):
    """The host side of fluxb200_ff_assemble on the emulated library: sub-slab pipeline with and without the ramp,
    overflow retry, ordinary-memory output, recycled page-locked blocks, host / device index expansion, the
    handle's cache of prepared column sets, device-resident slab -> disk."""
    env = dict(os.environ, FLUXB200_TEST_EMU='1')
    out = subprocess.run([sys.executable, '-m', 'pytest', 'tests/test_gpu_parity.py', 'tests/test_gpu_meshes.py', '-m', 'gpu',
                          '-q', '-x', '-k', 'test_streaming_and_two_phase_paths_agree or '
                          'test_block_assembly_reuses_prepared_column_sets or test_device_resident_slab_to_disk_roundtrip',
                          '-p', 'no:cacheprovider'], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0 and '3 passed' in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]


HORIZON_SCRIPT = r'''
import sys, json, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + '/tools/simt')
import build_emu
from fluxpy_b200 import _lib
_lib.SO_PATH = build_emu.build(); _lib._lib = None
import fluxpy_b200
from fluxpy_b200 import meshes
from oracle import oracle
res = []
for (n, seed, scale, zone, nrows, flip) in [(40, 2, 1.0, 64, 0, False), (72, 0, 1.0, 256, 160, False),
                                             (72, 0, 25.0, 128, 96, False), (56, 3, 1000.0, 64, 96, False),
                                             (40, 1, 1.0, 32, 0, True)]:
    V, F = meshes.gaussian_crater(n, seed, dtype=np.float32)
    V = (V * np.float32(scale)).astype(np.float32)
    N = meshes.upward_normals(V, F)
    if flip:  # user-mutated normals (reference tests/test_form_factors.py:33-34): every third face flipped
        N[::3] *= -1
    nf = len(F)
    I = None if nrows == 0 else np.linspace(0, nf - 1, nrows).astype(np.int64)
    sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, N.copy())
    sm.set_option('horizon_skip', 0)
    F0 = fluxpy_b200.get_form_factor_matrix(sm, I)
    sm.set_option('horizon_zone', zone)
    sm.set_option('horizon_skip', 1)
    F1 = fluxpy_b200.get_form_factor_matrix(sm, I)
    c = sm.trace_counters()
    same = bool(F0.nnz == F1.nnz and np.array_equal(F0.indptr, F1.indptr) and np.array_equal(F0.indices, F1.indices)
                and np.array_equal(F0.data, F1.data))
    oracle_same = None
    if nrows == 0:
        om = oracle.OracleShapeModel(V, F, N=N.copy())
        FO = oracle.get_form_factor_matrix(om); FO.sort_indices()
        oracle_same = bool(FO.nnz == F1.nnz and np.array_equal(FO.indices, F1.indices) and np.array_equal(FO.data, F1.data))
    res.append(dict(case=[n, seed, scale, zone, nrows, flip], same=same, oracle_same=oracle_same, nnz=int(F1.nnz), **c))
print('RESULT ' + json.dumps(res))
'''


def test_horizon_skip_is_exact_and_exercised(emu_lib):
    """Horizon skip on == off bit for bit (and == the oracle on the full matrices), at three length scales,
    with user-flipped normals, and the counters show that both ends of the skip are actually taken."""
    import json
    out = subprocess.run([sys.executable, '-c', HORIZON_SCRIPT, ROOT], cwd=ROOT, capture_output=True, text=True,
                         timeout=1500)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith('RESULT ')][-1]
    res = json.loads(line[len('RESULT '):])
    for r in res:
        assert r['same'], r
        assert r['oracle_same'] in (None, True), r
        assert r['batches'] > 0, r
    # unit-scale craters: most batches leave above the source horizon, many rays arrive above the target's
    assert res[1]['batches_source_skip'] > 0.5 * res[1]['batches'], res[1]
    assert res[1]['rays_target_skip'] > 0.3 * res[1]['rays'], res[1]
    assert res[0]['batches_source_skip'] > 0 and res[0]['rays_target_skip'] > 0, res[0]


def test_fuzz_kernels_against_the_oracle(emu_lib):
    """A short run of tools/simt/fuzz.py: random small meshes (soups, degenerate and coincident triangles,
    sheets, fans, closed bodies, translated and rescaled coordinates, user-supplied normals), random index
    sets, both dtypes, horizon skip off and on -- every CSR bit-identical to the oracle's."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'simt', 'fuzz.py'), '25', '500000'],
                         cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and 'fuzz ok' in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


_SHARDED_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
sys.path.insert(0, os.path.join(sys.argv[1], 'tools', 'simt'))
import build_emu
from fluxpy_b200 import _lib
_lib.SO_PATH = build_emu.build()            # test infrastructure: the kernels' own source on the SIMT emulator
import torch.distributed as dist
rank, world = int(sys.argv[3]), int(sys.argv[4])
dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{sys.argv[2]}', rank=rank, world_size=world)
import fluxpy_b200
from fluxpy_b200 import meshes, sharded, io as ffio
V, F = meshes.gaussian_crater(14, 3, dtype=np.float32)
sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, meshes.upward_normals(V, F))
nf = sm.num_faces
full = fluxpy_b200.get_form_factor_matrix(sm)                       # the one-process answer, on every rank
for weights in (None, np.diff(full.indptr)):                        # equal slabs / slabs balanced by row counts
    res = sharded.get_form_factor_matrix_sharded(sm, weights=weights)
    assert np.array_equal(res.global_indptr, full.indptr.astype(np.int64))
    mine = full[res.row_start:res.row_stop]
    assert np.array_equal(res.local_csr.indptr, mine.indptr) and np.array_equal(res.local_csr.indices, mine.indices)
    assert np.array_equal(res.local_csr.data, mine.data) and res.nnz_offset == full.indptr[res.row_start]
# an index subset in caller order, sharded
I = np.random.default_rng(5).permutation(nf)[:nf//2]
resI = sharded.get_form_factor_matrix_sharded(sm, I)
sub = full[I][resI.row_start:resI.row_stop]
sub.sort_indices()
assert np.array_equal(resI.local_csr.data, sub.data) and np.array_equal(resI.local_csr.indices, sub.indices)
# device-resident slabs -> one file per rank + manifest -> the reference's matrix
resd = sharded.get_form_factor_matrix_sharded(sm, to_host=False)
prefix = sys.argv[5]
ffio.save_sharded_result(prefix, resd, full.shape, world, rank)
dist.barrier()
if rank == 0:
    back = ffio.load_sharded(prefix)
    back.sort_indices()
    assert np.array_equal(back.indptr, full.indptr) and np.array_equal(back.indices, full.indices)
    assert np.array_equal(back.data, full.data)
dist.barrier()
dist.destroy_process_group()
print('ok', rank)
'''


def test_row_sharded_assembly_on_two_ranks(emu_lib, tmp_path):
    """SURVEY section 8e on the CPU tier: two gloo ranks, each with its own (emulated) device library, assemble
    their row slabs; slabs, global indptr, index subsets and the slab files equal the one-process matrix.  The
    NCCL version of this runs on two B200s (tests/test_gpu_multi.py), which a one-GPU test box skips."""
    import socket
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / 'sharded_worker.py'
    script.write_text(_SHARDED_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), '2', str(tmp_path / 'ff')],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]

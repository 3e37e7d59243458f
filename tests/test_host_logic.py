"""CPU tier: C-ABI surface, host-side logic, multi-rank row-count exchange (gloo)."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shared_library_exports_every_declared_symbol():
    from fluxpy_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'fluxb200.h')).read()
    declared = set(re.findall(r'\b(fluxb200_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().fluxb200_abi_version() == _lib.ABI_VERSION


def test_no_oracle_in_product_path():
    """The shipped package never imports, links or executes the oracle."""
    pkg = os.path.join(ROOT, 'fluxpy_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')) or f == 'Makefile':
                txt = open(os.path.join(dirpath, f)).read()
                for needle in ('import oracle', 'from oracle', 'ff_oracle', 'oracle/', 'oracle.'):
                    assert needle not in txt, (f, needle)


def test_missing_device_fails_loudly():
    """No CUDA device in the build container: constructing a shape model must
    raise, not fall back to a CPU path."""
    from fluxpy_b200 import _lib, meshes, shape
    try:
        n = _lib.device_count()
    except RuntimeError:
        n = 0
    if n:
        pytest.skip('a GPU is present')
    V, F = meshes.gaussian_crater(8, 0)
    with pytest.raises(RuntimeError):
        shape.CudaTrimeshShapeModel(V, F)
    from fluxpy_b200.form_factors import get_form_factor_matrix

    class Fake:
        dtype = np.dtype(np.float32)
    with pytest.raises(RuntimeError):
        get_form_factor_matrix(Fake())


def _expand(words, width):
    from fluxpy_b200 import _lib
    words = np.ascontiguousarray(words, np.uint32)
    out = np.full(int(np.unpackbits(words.view(np.uint8)).sum()) + 4, -7, np.int32 if width == 4 else np.int64)
    cnt = ctypes.c_int64(-1)
    _lib.check(_lib.lib().fluxb200_expand_words(_lib.ptr(words), len(words), width, _lib.ptr(out),
                                                ctypes.byref(cnt)))
    assert (out[cnt.value:] == -7).all()          # nothing written past the row
    return out[:cnt.value]


def test_expand_words_matches_numpy():
    """Host half of the copy-out: set-bit positions of the visibility words ==
    np.flatnonzero of the unpacked bits (ascending), int32 and int64, ragged
    lengths, empty / full / single-bit words."""
    rng = np.random.default_rng(5)
    cases = [np.zeros(0, np.uint32), np.zeros(7, np.uint32), np.full(5, 0xffffffff, np.uint32),
             np.array([1, 0x80000000, 0x00010000, 0x0000ffff, 0xffff0000], np.uint32)]
    for nw in (1, 2, 31, 33, 1000, 6241):
        cases.append(rng.integers(0, 2**32, nw, dtype=np.uint64).astype(np.uint32))
        cases.append((rng.integers(0, 2**32, nw, dtype=np.uint64) & rng.integers(0, 2**32, nw, dtype=np.uint64)
                      & rng.integers(0, 2**32, nw, dtype=np.uint64)).astype(np.uint32))
    for w in cases:
        want = np.flatnonzero(np.unpackbits(w.view(np.uint8), bitorder='little'))
        for width in (4, 8):
            assert np.array_equal(_expand(w, width), want)


def test_expand_words_any_alignment_and_no_stray_writes():
    """The AVX-512 path writes whole 64-byte lines with streaming stores between a masked head and a
    masked tail: every start alignment, rows shorter than a line, guards on both sides."""
    from fluxpy_b200 import _lib
    rng = np.random.default_rng(9)
    for nw in (1, 2, 5, 40, 700):
        for dens in (0.02, 0.5, 1.0):
            bits = rng.random(nw*32) < dens
            w = np.packbits(bits, bitorder='little').view(np.uint32)
            want = np.flatnonzero(bits)
            for width, dt in ((4, np.int32), (8, np.int64)):
                for shift in range(17):
                    buf = np.full(len(want) + 64, -7, dt)
                    out = buf[16 + shift:]
                    cnt = ctypes.c_int64(-1)
                    _lib.check(_lib.lib().fluxb200_expand_words(_lib.ptr(w), len(w), width,
                                                                out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cnt)))
                    assert cnt.value == len(want) and np.array_equal(out[:len(want)], want)
                    assert (buf[:16 + shift] == -7).all() and (out[len(want):] == -7).all()


def test_expand_rows_thread_pool():
    """The worker pool of the copy-out: ragged rows packed back to back (so neighbouring rows share
    cache lines across threads), 1 to 5 threads, int32 and int64; a wrong row length is an error."""
    from fluxpy_b200 import _lib
    rng = np.random.default_rng(12)
    nw, mr = 97, 301
    dens = rng.random(mr)**2
    bits = rng.random((mr, nw*32)) < dens[:, None]
    bits[5] = False
    bits[7] = True
    words = np.ascontiguousarray(np.packbits(bits, axis=1, bitorder='little')).view(np.uint32).reshape(mr, nw)
    counts = bits.sum(1)
    offs = np.zeros(mr + 1, np.int64)
    np.cumsum(counts, out=offs[1:])
    offs += 3                                        # unaligned start
    want = np.concatenate([np.flatnonzero(b) for b in bits])
    for width, dt in ((4, np.int32), (8, np.int64)):
        for nthreads in (1, 2, 5, 0):
            out = np.full(int(offs[-1]) + 40, -7, dt)
            _lib.check(_lib.lib().fluxb200_expand_rows(_lib.ptr(words), nw, mr, _lib.ptr(offs), width, _lib.ptr(out),
                                                       nthreads))
            assert np.array_equal(out[3:int(offs[-1])], want)
            assert (out[:3] == -7).all() and (out[int(offs[-1]):] == -7).all()
    bad = offs.copy()
    bad[100:] += 1                                   # row 99 one entry too long
    out = np.full(int(bad[-1]) + 40, -7, np.int32)
    with pytest.raises(RuntimeError, match='do not match'):
        _lib.check(_lib.lib().fluxb200_expand_rows(_lib.ptr(words), nw, mr, _lib.ptr(bad), 4, _lib.ptr(out), 3))


def test_expand_portable_path():
    """The same checks with the AVX-512 path switched off (fresh process: the choice is cached)."""
    env = dict(os.environ, FLUXB200_NO_AVX512='1')
    out = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-x', os.path.abspath(__file__), '-k',
                          'matches_numpy or any_alignment or thread_pool'], env=env, cwd=ROOT,
                         capture_output=True, text=True)
    assert out.returncode == 0 and '3 passed' in out.stdout, out.stdout + out.stderr


def test_slab_plan():
    from fluxpy_b200 import _lib
    assert _lib.slab_plan(10, 3).tolist() == [0, 3, 6, 10]
    assert _lib.slab_plan(0, 4).tolist() == [0, 0, 0, 0, 0]
    assert _lib.slab_plan(3, 8).tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 3]
    s = _lib.slab_plan(1000, 8, np.r_[np.zeros(900), np.full(100, 1000)])
    assert s[0] == 0 and s[-1] == 1000 and (np.diff(s) >= 0).all() and s[1] >= 900
    w = np.random.default_rng(0).integers(0, 100, 5000)
    s = _lib.slab_plan(5000, 8, w)
    loads = [w[s[k]:s[k + 1]].sum() for k in range(8)]
    assert max(loads) < 1.1*np.mean(loads)


def test_geometry_helpers_match_reference_formulas():
    from fluxpy_b200 import meshes, shape
    V, F = meshes.gaussian_crater(12, 3, dtype=np.float64)
    N, A = shape.get_surface_normals_and_face_areas(V, F)
    assert np.allclose(np.linalg.norm(N, axis=1), 1) and (N[:, 2] > 0).all()
    assert np.isclose(A.sum(), shape.get_face_areas(V, F).sum())
    assert np.allclose(shape.get_centroids(V, F), V[F].mean(1))
    assert meshes.gaussian_crater(72)[1].shape[0] == 10082
    assert meshes.grid_faces(159).shape[0] == 49928 and meshes.grid_faces(317).shape[0] == 199712
    Vs, Fs = meshes.icosphere(2)
    assert Fs.shape[0] == 320 and np.allclose(np.linalg.norm(Vs, axis=1), 1)
    Ns = shape.get_surface_normals(Vs, Fs)
    assert ((Ns*shape.get_centroids(Vs, Fs)).sum(1) > 0).all()


def test_block_index_sets_match_reference_golden():
    """get_quadrant_order / get_octant_order (quadtree.py:5-18, octree.py:5-18)."""
    from fluxpy_b200 import blocks
    from tests import helpers
    g = helpers.load('block_inds')
    P = g['P']
    for k, I in enumerate(blocks.get_quadrant_order(P[:, :2])):
        assert np.array_equal(I, g[f'quad{k}'])
    for k, I in enumerate(blocks.get_octant_order(P)):
        assert np.array_equal(I, g[f'oct{k}'])


_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from fluxpy_b200 import sharded
dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{sys.argv[2]}',
                        rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, world = dist.get_rank(), dist.get_world_size()
m = 37
rng = np.random.default_rng(5)
counts = rng.integers(0, 1000, m).astype(np.int64)      # what a 1-GPU pass would count
starts = sharded.slab_bounds(m, world)
mine = counts[starts[rank]:starts[rank + 1]]
indptr = sharded.exchange_row_counts(mine, starts)
expect = np.r_[0, np.cumsum(counts)]
assert np.array_equal(indptr, expect), (indptr, expect)
# weighted plan agrees on every rank and covers all rows
s2 = sharded.slab_bounds(m, world, counts)
assert s2[0] == 0 and s2[-1] == m
indptr2 = sharded.exchange_row_counts(counts[s2[rank]:s2[rank + 1]], s2)
assert np.array_equal(indptr2, expect)
dist.barrier()
dist.destroy_process_group()
print('ok', rank)
'''


@pytest.mark.parametrize('world', [2, 3])
def test_row_count_exchange_gloo(world, tmp_path):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), str(world)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_install_into_reference_package():
    """INTEGRATION.md section 1 against the real reference package, when it is
    present (build container only; the GPU box has no /root/reference)."""
    ref = '/root/reference/src'
    if not os.path.isdir(ref):
        pytest.skip('reference tree not present')
    code = r'''
import sys, types, functools
cp = types.ModuleType('cached_property'); cp.cached_property = functools.cached_property
sys.modules['cached_property'] = cp
sys.path.insert(0, %r); sys.path.insert(0, %r)
import flux.shape, flux.form_factors
from fluxpy_b200 import integration, CudaTrimeshShapeModel
orig = flux.form_factors.get_form_factor_matrix
f = integration.install()
assert flux.shape.trimesh_shape_models[-1] is CudaTrimeshShapeModel
assert flux.form_factors.get_form_factor_matrix is f and f is not orig
integration.install()                                   # idempotent
assert flux.shape.trimesh_shape_models.count(CudaTrimeshShapeModel) == 1
# a non-CUDA model still goes to the reference implementation
class Dummy: pass
try:
    f(Dummy())
except AttributeError:
    pass
else:
    raise SystemExit('reference path not taken')
# same public surface as the reference's shape model
import inspect
for name in ('get_visibility', 'get_visibility_1_to_N', 'get_visibility_matrix', 'is_occluded',
             'intersect1', 'get_direct_irradiance', 'num_faces', 'num_verts'):
    assert hasattr(CudaTrimeshShapeModel, name), name
a = inspect.signature(flux.shape.TrimeshShapeModel.__init__)
b = inspect.signature(CudaTrimeshShapeModel.__init__)
assert list(a.parameters) == list(b.parameters)
print('ok')
''' % (ROOT, ref)
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and 'ok' in out.stdout, out.stdout + out.stderr


def test_sharded_npz_roundtrip(tmp_path):
    """Next row N4: per-slab scipy.sparse.save_npz files + manifest == the
    reference's single save_npz file (collect_data.py:141)."""
    import scipy.sparse
    from fluxpy_b200 import io as ffio, sharded
    rng = np.random.default_rng(0)
    FF = scipy.sparse.random(101, 57, density=0.3, format='csr', random_state=rng, dtype=np.float32)
    world = 4
    starts = sharded.slab_bounds(FF.shape[0], world)
    prefix = str(tmp_path / 'FF')
    for r in range(world):
        ffio.save_slab(prefix, r, world, FF[starts[r]:starts[r + 1]], starts[r], FF.shape, FF.indptr.astype(np.int64))
    back = ffio.load_sharded(prefix)
    assert back.shape == FF.shape and back.dtype == FF.dtype and (back != FF).nnz == 0
    assert np.array_equal(back.indptr, FF.indptr) and np.array_equal(back.indices, FF.indices)
    one = scipy.sparse.load_npz(ffio.slab_path(prefix, 2))        # each slab is a plain save_npz file
    assert (one != FF[starts[2]:starts[3]]).nnz == 0
    part = ffio.load_sharded(prefix, rows=(30, 70))
    assert (part != FF[30:70]).nnz == 0
    man = ffio.load_manifest(prefix)
    assert man['nnz'] == FF.nnz and man['shape'] == [101, 57] and len(man['slabs']) == world
    # the reference's own file for the same matrix holds the same arrays
    scipy.sparse.save_npz(str(tmp_path / 'ref.npz'), FF)
    ref = scipy.sparse.load_npz(str(tmp_path / 'ref.npz'))
    assert np.array_equal(ref.data, back.data)


def test_reference_own_unit_tests_on_the_cuda_backend():
    """The reference's UNMODIFIED tests/test_shape.py, test_form_factors.py and
    test_compressed_form_factors.py, run from /root/reference against CudaTrimeshShapeModel as the only
    entry of flux.shape.trimesh_shape_models -- once through the reference's own row loop
    (form_factors.py:11-72 calling get_visibility_1_to_N per row) and once through the fused assembly
    installed by fluxpy_b200.integration -- plus the row loop vs fused cross-check on an occluded crater
    (tools/run_reference_tests.py; kernels on the SIMT emulator here).  Build container only."""
    if not os.path.isdir('/root/reference/src'):
        pytest.skip('reference tree not present')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'run_reference_tests.py'), 'both'],
                         capture_output=True, text=True, timeout=900)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-4000:]
    assert 'FAILED on the backend' not in text
    for mode in ('loop', 'fused'):
        assert f'[{mode}] reference tests run 7' in text
        assert text.count(f'[{mode}] is_occluded adjudication') == 2
    assert text.count('backend != exact convex answer on 0 faces') == 4
    assert text.count('pattern identical: True') == 4


def test_bench_host_logic(tmp_path, monkeypatch):
    """bench.py's bookkeeping: the slab sequence covers the mesh for every rank count, the flop model is SURVEY
    section 8d's, and `roofline.traffic` is only reported for a capture of exactly the sources that are built."""
    import json
    import sys
    sys.path.insert(0, ROOT)
    import bench
    nf, rows = 199712, 4096
    nslabs = nf//rows
    for world in (1, 2, 4, 8):
        seen = set()
        for step in range(nslabs):
            for rank in range(world):
                r = bench.slab_rows(step, rank, world, rows, nf)
                assert len(r) == rows and r[0] % rows == 0 and r[-1] < nf
                seen.add(int(r[0])//rows)
        assert seen == set(range(nslabs))                 # stride 7 is coprime to the slab count: all slabs sampled
        assert len({int(bench.slab_rows(0, rank, world, rows, nf)[0]) for rank in range(world)}) == world
    # 22 flop per candidate pair + (50 ceil(log2 Nf) + 50) per traced ray + 6 per stored entry
    assert bench.alg_flops(10, 3, 2, 199712) == 22*10 + 3*(50*18 + 50) + 6*2
    assert bench.alg_flops(1, 1, 0, 1 << 10) == 22 + 50*10 + 50
    # the committed capture counts only if it belongs to the sources in the tree (sources edited since: no figure) ...
    sha = bench.source_sha16()
    t = bench.measured_traffic()
    if t is not None:
        if t.get('source_sha16') == sha:
            assert t['dram_bytes_per_launch'] > 0
        else:
            assert t['dram_bytes_per_launch'] is None and 'no traffic figure' in t['note']
    # ... and a capture of other sources is refused
    fake = tmp_path / 'profiles'
    fake.mkdir()
    (fake / 'r99_trace_kernel_traffic.json').write_text(json.dumps({'source_sha16': 'f'*16, 'dram_bytes_per_launch': 1}))
    monkeypatch.setattr(bench, 'ROOT', str(tmp_path))
    monkeypatch.setattr(bench, 'source_sha16', lambda: sha)
    t2 = bench.measured_traffic()
    assert t2['dram_bytes_per_launch'] is None and 'no traffic figure' in t2['note']

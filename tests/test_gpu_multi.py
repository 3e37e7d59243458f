"""GPU tier, N > 1: the row-sharded assembly and the sharded Jacobi solve on two
ranks (NCCL) reproduce the single-GPU result.  Skipped when the box has one GPU."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank, world = int(sys.argv[3]), int(sys.argv[4])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', init_method=f'tcp://127.0.0.1:{sys.argv[2]}', rank=rank, world_size=world,
                        device_id=torch.device('cuda', rank))
import scipy.sparse
import fluxpy_b200
from fluxpy_b200 import meshes, sharded, solve
V, F = meshes.gaussian_crater(40, 2, dtype=np.float32)
N = meshes.upward_normals(V, F)
fluxpy_b200.CudaTrimeshShapeModel.device = rank
sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, N)
nf = sm.num_faces
full = fluxpy_b200.get_form_factor_matrix(sm)                      # every rank: the 1-GPU answer
res = sharded.get_form_factor_matrix_sharded(sm)                   # my slab, host CSR
assert np.array_equal(res.global_indptr, full.indptr.astype(np.int64))
mine = full[res.row_start:res.row_stop]
assert np.array_equal(res.local_csr.indices, mine.indices) and np.array_equal(res.local_csr.data, mine.data)
assert res.nnz_offset == full.indptr[res.row_start]
# weighted slabs (balanced by the row counts of the first pass) cover the same matrix
w = np.diff(full.indptr)
res2 = sharded.get_form_factor_matrix_sharded(sm, weights=w)
mine2 = full[res2.row_start:res2.row_stop]
assert np.array_equal(res2.local_csr.data, mine2.data)
# device-resident slabs + sharded Jacobi == single-GPU host solve
resd = sharded.get_form_factor_matrix_sharded(sm, to_host=False)
assert resd.device_csr.shape == (res.row_stop - res.row_start, nf) and resd.device_csr.m_global == nf
E = sm.get_direct_irradiance(1365.0, np.array([0.98, 0, 0.17], np.float32)).astype(np.float64)
B, nit = solve.solve_radiosity(resd.device_csr, E, 0.2)
from oracle import radiosity
Bref, nref = radiosity.solve_radiosity_jacobi_right(full, E, 0.2)
assert nit == nref and np.allclose(B, Bref, rtol=1e-13, atol=1e-10)
T = solve.compute_steady_state_temp(resd.device_csr, E, 0.2, 0.95)
Tref = radiosity.compute_steady_state_temp(full, E, 0.2, 0.95)
assert np.linalg.norm(T - Tref) <= 1e-9*np.linalg.norm(Tref)
# N4: the device-resident slabs go to disk, one save_npz file per rank + a manifest, and load back as `full`
from fluxpy_b200 import io as ffio
prefix = sys.argv[5]
ffio.save_sharded_result(prefix, resd, full.shape, world, rank)
dist.barrier()
if rank == 0:
    back = ffio.load_sharded(prefix)
    back.sort_indices()
    assert np.array_equal(back.indptr, full.indptr) and np.array_equal(back.indices, full.indices)
    assert np.array_equal(back.data, full.data) and ffio.load_manifest(prefix)['nnz'] == full.nnz
dist.barrier()
dist.destroy_process_group()
print('ok', rank)
'''


def test_two_rank_sharded_assembly_and_solve(tmp_path):
    from fluxpy_b200 import _lib
    if _lib.device_count() < 2:
        pytest.skip('needs two GPUs')
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), '2', str(tmp_path / 'ff')],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]

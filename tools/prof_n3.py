"""The kernels of the path's next row N3 (block extraction + thin products for the low-rank feed) on the
device-resident 50k-face matrix: times from the library's own events and the bandwidth they amount to.

    python tools/prof_n3.py            (under ncu: -k regex:"extract|csr_matmat|csr_rmatmat")
"""
import json
import sys
import numpy as np
sys.path.insert(0, '.')
import torch
import fluxpy_b200
from fluxpy_b200 import blocks, meshes, get_form_factor_matrix_device

V, F = meshes.gaussian_crater(159, 0, dtype=np.float32)
sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, meshes.upward_normals(V, F))
FFd = get_form_factor_matrix_device(sm)
m, n = FFd.shape
parts = [p for p in blocks.get_quadrant_order(sm.P[:, :2]) if len(p)]
out = {'faces': m, 'nnz': int(FFd.nnz), 'csr_bytes': int(FFd.nbytes)}
dev = torch.device('cuda', 0)
rng = np.random.default_rng(0)
for rep in range(2):
    B = FFd.extract(parts[0], parts[1])               # an off-diagonal quadrant block
    t_extract = FFd.last_ms()
    X = torch.as_tensor(rng.normal(size=(n, 32)), device=dev)
    Y = FFd.matmat(X)
    t_mm = FFd.last_ms()
    Z = FFd.rmatmat(Y)
    t_rmm = FFd.last_ms()
    Xb = torch.as_tensor(rng.normal(size=(B.shape[1], 32)), device=dev)
    Yb = B.matmat(Xb)
    t_mmb = B.last_ms()
    Zb = B.rmatmat(Yb)
    t_rmmb = B.last_ms()
ent = 8  # bytes per stored entry (float32 value + int32 column)
out.update({
    'extract_quadrant_ms': t_extract, 'extract_block_nnz': int(B.nnz),
    'extract_gbs_source_rows_read_twice': 2*ent*FFd.nnz*len(parts[0])/m/1e9/(t_extract/1e3),
    'matmat_k32_ms': t_mm, 'matmat_k32_csr_gbs': ent*FFd.nnz/1e9/(t_mm/1e3), 'matmat_k32_gflops': 2*32*FFd.nnz/1e9/(t_mm/1e3),
    'rmatmat_k32_ms': t_rmm, 'rmatmat_k32_csr_gbs': ent*FFd.nnz/1e9/(t_rmm/1e3), 'rmatmat_k32_gflops': 2*32*FFd.nnz/1e9/(t_rmm/1e3),
    'block_matmat_k32_ms': t_mmb, 'block_rmatmat_k32_ms': t_rmmb, 'block_shape': list(B.shape)})
print(json.dumps(out))

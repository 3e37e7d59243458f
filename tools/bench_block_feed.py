"""Timing of the N3 feed on the 50k-face crater: quadrant-block extraction from the resident
matrix, thin products, randomised SVD of one far-field block."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import fluxpy_b200
from fluxpy_b200 import meshes, blocks, lowrank, get_form_factor_matrix_device
V, F = meshes.gaussian_crater(159, 0, dtype=np.float32)
sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, meshes.upward_normals(V, F))
FF = get_form_factor_matrix_device(sm)
parts = blocks.get_quadrant_order(sm.P[:, :2])
out = {'faces': sm.num_faces, 'nnz': FF.nnz}
t = time.perf_counter()
B = [[FF.extract(I, J) for J in parts] for I in parts]
torch.cuda.synchronize(); out['extract_16_blocks_s'] = time.perf_counter() - t
out['extract_gbs_of_source'] = 4*FF.nnz*8/out['extract_16_blocks_s']/1e9     # every block row re-reads its source rows: 4 passes
blk = B[0][3]
out.update(block_shape=list(blk.shape), block_nnz=blk.nnz)
X = torch.randn(blk.shape[1], 32, dtype=torch.float64, device='cuda')
for name, fn, Xin in (('matmat', blk.matmat, X), ('rmatmat', blk.rmatmat, torch.randn(blk.shape[0], 32, dtype=torch.float64, device='cuda'))):
    ms = []
    for _ in range(5):
        fn(Xin); ms.append(blk.last_ms())
    out[name + '_k32_ms'] = float(np.median(ms)); out[name + '_entries_per_s'] = blk.nnz/np.median(ms)*1e3
t = time.perf_counter(); U, S, Vt = lowrank.sparse_svd(blk, 40); out['sparse_svd_k40_s'] = time.perf_counter() - t
out['sigma_0_39'] = [float(S[0]), float(S[39])]
print(json.dumps(out))

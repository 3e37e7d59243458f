"""BASELINE config 2 end to end on one GPU: G(159,0) (49 928 faces) full form-factor
matrix assembled device-resident, then the steady-state thermal solve on it
(Haworth stand-in parameters: F0=1365, e0=3 deg, rho=0.12, emiss=0.95,
examples/haworth_crater/haworth.py:50-54,129)."""
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
import fluxpy_b200
from fluxpy_b200 import meshes, solve, get_form_factor_matrix_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 159
V, F = meshes.gaussian_crater(n, 0, dtype=np.float32)
V *= 25.0                                            # km, as SURVEY 8d config 2
sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, meshes.upward_normals(V, F))
nf = sm.num_faces
out = {'faces': nf}
for rep in range(2):
    t = time.perf_counter(); FF = get_form_factor_matrix_device(sm); out['assemble_s'] = time.perf_counter() - t
    if rep == 0: del FF
out.update(nnz=FF.nnz, csr_gb=FF.nbytes/1e9, pairs=nf*nf)
e0 = np.deg2rad(3.0)
Dsun = np.array([0, -np.cos(e0), np.sin(e0)], np.float32)
t = time.perf_counter(); E = sm.get_direct_irradiance(1365.0, Dsun); out['irradiance_s'] = time.perf_counter() - t
out['lit_fraction'] = float((E > 0).mean())
x = np.random.default_rng(0).random(nf)
import torch
xt = torch.as_tensor(x, device='cuda')
ms = []
for rep in range(10):
    FF.step(xt); ms.append(FF.last_ms())
out['spmv_ms'] = float(np.median(ms)); out['spmv_gbs'] = FF.nnz*8/np.median(ms)/1e6
t = time.perf_counter(); T = solve.compute_steady_state_temp(FF, E.astype(np.float64), 0.12, 0.95); out['steady_state_s'] = time.perf_counter() - t
B, nit = solve.solve_radiosity(FF, E.astype(np.float64), 0.12)
out.update(jacobi_iters_visible=nit, T_min=float(T.min()), T_max=float(T.max()), T_mean=float(T.mean()))
print(json.dumps(out))

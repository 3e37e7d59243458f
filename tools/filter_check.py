import sys, numpy as np
sys.path.insert(0, '.')
import fluxpy_b200
from fluxpy_b200 import meshes
from oracle import oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 159
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 25.0
V, F = meshes.gaussian_crater(n, 0, dtype=np.float32); V *= scale
N = meshes.upward_normals(V, F)
sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, N.copy())
res = {}
for flt in (1, 0):
    sm.set_option('shaft_filter', flt)
    m, nn, counts, st = sm._ff_assemble_device(None, None, 1e-5, 4, want_row_counts=True)
    res[flt] = counts.copy(); print('filter', flt, 'nnz', st.nnz, 'tested', st.pairs_tested)
bad = np.where(res[0] != res[1])[0]
print('rows differing', bad, res[1][bad] - res[0][bad])
if len(bad):
    om = oracle.OracleShapeModel(V, F, N=N.copy())
    FO = oracle.get_form_factor_matrix(om, bad[:8])
    print('oracle counts', np.diff(FO.indptr), 'filter-off', res[0][bad[:8]], 'filter-on', res[1][bad[:8]])

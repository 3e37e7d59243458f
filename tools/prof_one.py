"""One small assembly for ncu: G(317,0), a slab of rows x all columns."""
import sys
import numpy as np
sys.path.insert(0, '.')
import fluxpy_b200
from fluxpy_b200 import meshes
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 317
V, F = meshes.gaussian_crater(n, 0, dtype=np.float32)
sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, meshes.upward_normals(V, F))
nf = sm.num_faces
I = np.arange(rows) + nf//2
for rep in range(2):
    m, ncol, _, st = sm._ff_count(I, None, 1e-5, want_row_counts=False)
    st2 = sm._ff_fill_device(4)
print(st.as_dict(), st2.ms_fill)

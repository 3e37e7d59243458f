"""One small assembly for ncu: G(n,0), a slab of rows x all columns.

    python tools/prof_one.py [rows] [grid_n] [option=value ...]      e.g.  4096 317 horizon_skip=1
"""
import sys
import numpy as np
sys.path.insert(0, '.')
import fluxpy_b200
from fluxpy_b200 import meshes
pos = [a for a in sys.argv[1:] if '=' not in a]
opts = [a.split('=', 1) for a in sys.argv[1:] if '=' in a]
rows = int(pos[0]) if len(pos) > 0 else 512
n = int(pos[1]) if len(pos) > 1 else 317
V, F = meshes.gaussian_crater(n, 0, dtype=np.float32)
sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, meshes.upward_normals(V, F))
for name, value in opts:
    sm.set_option(name, int(value))
nf = sm.num_faces
I = np.arange(rows) + nf//2
import os
for rep in range(int(os.environ.get('PROF_ONE_REPS', '2'))):
    m, ncol, _, st = sm._ff_count(I, None, 1e-5, want_row_counts=False)
    st2 = sm._ff_fill_device(4)
    print('rep', rep, 'trace ms %.3f fill ms %.3f' % (st.ms_trace, st2.ms_fill), flush=True)
print(st.as_dict(), st2.ms_fill, sm.trace_counters())

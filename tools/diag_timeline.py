"""Scratch: per-sub-slab timeline of one host-output assembly (FLUXB200_TIMELINE)."""
import os, sys, time
import numpy as np
sys.path.insert(0, '.')
import fluxpy_b200
from fluxpy_b200 import meshes, form_factors
V, F = meshes.gaussian_crater(317, 0, dtype=np.float32)
sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, meshes.upward_normals(V, F))
for o in sys.argv[1:]:
    k, v = o.split('=')
    sm.set_option(k, int(v))
I = np.arange(4096) + 4096*20
for rep in range(8):
    os.environ['FLUXB200_TIMELINE'] = f'gpurun_out/timeline_rep{rep}.csv' if rep >= 5 else ''
    if rep < 5:
        os.environ.pop('FLUXB200_TIMELINE')
    t = time.perf_counter()
    FF = fluxpy_b200.get_form_factor_matrix(sm, I)
    print(rep, round(1e3*(time.perf_counter() - t), 1), dict(form_factors.last_stats)['ms_trace'], flush=True)
    del FF

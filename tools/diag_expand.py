"""Scratch: end-to-end call time of get_form_factor_matrix with the column indices copied from
the device (host_expand 0) or expanded on the host from the visibility words (1)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import fluxpy_b200
from fluxpy_b200 import meshes, form_factors
V, F = meshes.gaussian_crater(317, 0, dtype=np.float32)
sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, meshes.upward_normals(V, F))
for expand, threads, sub in ((0, 0, 512), (1, 0, 512), (1, 4, 512), (1, 12, 512), (1, 8, 1024), (1, 8, 256), (0, 0, 512)):
    sm.set_option('host_expand', expand); sm.set_option('host_threads', threads); sm.set_option('sub_rows', sub)
    ts = []
    for rep in range(5):
        I = np.arange(4096) + 4096*(rep + 3)
        t = time.perf_counter()
        FF = fluxpy_b200.get_form_factor_matrix(sm, I)
        ts.append(1e3*(time.perf_counter() - t))
        st = dict(form_factors.last_stats)
        del FF
    print(f'expand={expand} threads={threads} sub={sub}: call ms {[round(x, 1) for x in ts]} trace={st["ms_trace"]:.1f} '
          f'copy_span={st["ms_fill"]:.1f} d2h={st["d2h_bytes"]/1e9:.2f} GB', flush=True)

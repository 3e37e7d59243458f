"""Scratch: per-step wall / device times of the device-resident assembly on one mesh size (the sweep arm of bench.py)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
import fluxpy_b200
from fluxpy_b200 import meshes
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 159
opts = [a.split('=') for a in sys.argv[2:]]
V, F = meshes.gaussian_crater(grid, 0, dtype=np.float32)
sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, meshes.upward_normals(V, F))
for k, v in opts:
    sm.set_option(k, int(v))
nf = F.shape[0]
rows = min(4096, nf)
stream = torch.cuda.ExternalStream(sm.cuda_stream(), device=0)
nslab = max(1, nf//rows)
for s in range(8):
    I = (np.arange(rows) + (s % nslab)*rows).astype(np.int64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    e0.record(stream)
    m, n, _, st = sm._ff_assemble_device(I, None, 1e-5, 4)
    e1.record(stream)
    torch.cuda.synchronize()
    w = 1e3*(time.perf_counter() - t)
    print(grid, s, 'wall %.2f' % w, 'events %.2f' % e0.elapsed_time(e1), 'trace %.2f fill %.2f prepare %.2f' % (st.ms_trace, st.ms_fill, st.ms_prepare),
          'nnz', st.nnz, 'launches', st.kernel_launches, flush=True)

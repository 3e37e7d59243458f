"""Scratch timing of the assembly stages at the BASELINE sizes (not a bench)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import os
import fluxpy_b200
from fluxpy_b200 import meshes, form_factors, _lib
if os.environ.get('FLUXB200_SO'):  # A/B against another build of the library
    _lib.SO_PATH = os.path.abspath(os.environ['FLUXB200_SO'])

def run(n, rows, dtype=np.float32, opts=()):
    V, F = meshes.gaussian_crater(n, 0, dtype=dtype)
    N = meshes.upward_normals(V, F)
    t = time.time(); sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, N); tb = time.time() - t
    for k, v in opts: sm.set_option(k, v)
    info = sm.bvh_info()
    nf = sm.num_faces
    I = None if rows is None else np.arange(min(rows, nf)) + (nf//2 if rows < nf else 0)
    for rep in range(2):
        t = time.time()
        m, ncol, _, st = sm._ff_count(I, None, 1e-5, want_row_counts=False)
        st2 = sm._ff_fill_device(4)
        dt = time.time() - t
    print(f'n={n} nf={nf} rows={m} dtype={np.dtype(dtype).name} opts={opts} build_ms={info.ms_build:.2f} (ctor {tb*1e3:.0f} ms) depth={info.max_depth} ntop={info.num_top_nodes} '
          f'pairs={st.pairs_all:.3e} tested={st.pairs_tested:.3e} nnz={st.nnz:.3e} prep={st.ms_prepare:.2f} trace={st.ms_trace:.2f} scan={st.ms_scan:.2f} fill={st2.ms_fill:.2f} '
          f'wall={dt*1e3:.1f} ms  -> {st.pairs_tested/st.ms_trace/1e6:.3f} Grays/s, {st.pairs_all/(st.ms_trace+st2.ms_fill+st.ms_prepare+st.ms_scan)/1e6:.3f} Gpairs/s', flush=True)

if __name__ == '__main__':
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == 'ab3':
        run(317, 4096, opts=(('blocks_per_sm', 3),))
        run(159, 4096, opts=(('blocks_per_sm', 3),))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'ab':
        run(317, 4096)
        run(159, 4096)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'quick':
        run(317, 4096)
        run(317, 4096, opts=(('blocks_per_sm', 4),))
        run(317, 4096, opts=(('top_nodes', 256),))
        run(317, 4096, opts=(('slab_limit', 4096),))
        run(159, 4096)
        run(501, 2048)
        sys.exit(0)
    run(72, None)
    run(159, 4096)
    run(317, 1024)
    run(317, 4096)
    run(317, 4096, np.float64)
    run(317, 4096, opts=(('top_nodes', 256),))
    run(317, 4096, opts=(('top_nodes', 1024),))
    run(317, 4096, opts=(('slab_limit', 0),))
    run(317, 4096, opts=(('slab_limit', 256),))
    run(317, 4096, opts=(('slab_limit', 4096),))
    run(317, 4096, opts=(('blocks_per_sm', 2),))
    run(317, 4096, opts=(('blocks_per_sm', 4),))
    run(501, 2048)
